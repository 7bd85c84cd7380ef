"""Recipe that makes the UNMODIFIED reference model code available to the checker on the GPU box — TEST INFRASTRUCTURE.

The reference's hot path is pure Python on torch (`src/models/{__init__,components,tts_model}.py`), so "building" it is
vendoring those three files, byte for byte, from where they lie under /root/reference into the git-ignored `oracle/_ref/`
(listed in .gitignore, NOT in .gpurunignore: it travels to the GPU box like a built .so; nothing under /root/reference is read
at run time there). `bench.py --impl reference` then times the reference's own `MelDecoder` / `SimpleVocoder` modules
(cpu_baseline.kind = "reference"); without `oracle/_ref` it falls back to the oracle port (kind = "port").
Nothing under m2-tts_b200/ imports this.
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.util
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference/src/models")
OUT = ROOT / "_ref" / "ref_models"
FILES = ("__init__.py", "components.py", "tts_model.py")


def build() -> bool:
    """Vendor the files when the reference is present; returns whether oracle/_ref is usable afterwards."""
    if REF_SRC.exists():
        OUT.mkdir(parents=True, exist_ok=True)
        digests = {}
        for name in FILES:
            src = REF_SRC / name
            shutil.copyfile(src, OUT / name)
            digests[name] = hashlib.sha256(src.read_bytes()).hexdigest()
        (OUT.parent / "MANIFEST.json").write_text(json.dumps({"source": str(REF_SRC), "sha256": digests}, indent=1))
    return available()


def available() -> bool:
    return all((OUT / name).exists() for name in FILES)


def load():
    """Import the vendored reference as package `ref_models` (does not collide with the mirror's `models`)."""
    if "ref_models" not in sys.modules:
        spec = importlib.util.spec_from_file_location("ref_models", OUT / "__init__.py", submodule_search_locations=[str(OUT)])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["ref_models"] = pkg
        spec.loader.exec_module(pkg)
    return importlib.import_module("ref_models.tts_model")


if __name__ == "__main__":
    print("oracle/_ref available:", build())
