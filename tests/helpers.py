"""Shared helpers for the test-suite: paths, seeded weights/inputs, oracle access.

Seeded construction is reproducible across machines with the same torch build: the CPU RNG
stream after ``torch.manual_seed`` is deterministic, and the mirror modules consume it in the
reference's order (checked against the live reference in ``test_oracle_golden.py`` and against
``tests/golden/state_digests.json`` everywhere else).
"""
from __future__ import annotations

import hashlib
import importlib
import importlib.util
import sys
from pathlib import Path
from typing import Dict, Optional

import numpy as np
import torch

REPO = Path(__file__).resolve().parents[1]
PKG_SRC = REPO / "m2-tts_b200" / "src"
GOLDEN = REPO / "tests" / "golden"
REFERENCE = Path("/root/reference")

for p in (str(PKG_SRC), str(REPO)):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import m2tts_oracle as oracle  # noqa: E402  (tests may import the oracle)

STAGE_KWARGS = oracle.STAGE_KWARGS
HELLO_WORLD_IDS = [39, 21, 6, 24, 11, 40, 35, 7, 24, 17, 39]  # TextProcessor("Hello world"), SURVEY §8d C1


def product_model(stage: str = "stage1", seed: int = 1234, perturb: Optional[int] = None, **override):
    """The B200 mirror model with seeded random-init weights (eval mode, CPU)."""
    from models.tts_model import M2TTSModel
    kw = dict(STAGE_KWARGS[stage]); kw.update(override)
    torch.manual_seed(seed)
    m = M2TTSModel(**kw)
    if perturb is not None:
        perturb_(m, perturb)
    return m.eval()


def perturb_(model: torch.nn.Module, seed: int) -> None:
    """Second weight set: random non-zero biases, LayerNorm affine, BatchNorm affine
    and running statistics — the zeros/ones of the default init hide padding and bias bugs."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, t in list(model.named_parameters()) + list(model.named_buffers()):
            if name.endswith("num_batches_tracked") or name.endswith("pos_encoding.pe"):
                continue
            if name.endswith("running_var"):
                t.copy_(torch.rand(t.shape, generator=g) + 0.5)
            elif name.endswith("running_mean") or name.endswith(".bias"):
                t.copy_(torch.randn(t.shape, generator=g) * 0.1)
            elif ("norm" in name) and name.endswith(".weight"):
                t.copy_(1.0 + torch.randn(t.shape, generator=g) * 0.1)


def state_digest(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def load_reference_module():
    """Import the UNMODIFIED reference model code under the package name ``ref_models`` (so it
    does not collide with the mirror's ``models``). Only possible where /root/reference exists."""
    if "ref_models" not in sys.modules:
        src = REFERENCE / "src" / "models"
        spec = importlib.util.spec_from_file_location("ref_models", src / "__init__.py",
                                                      submodule_search_locations=[str(src)])
        pkg = importlib.util.module_from_spec(spec)
        sys.modules["ref_models"] = pkg
        spec.loader.exec_module(pkg)
    return importlib.import_module("ref_models.tts_model")


def have_reference() -> bool:
    return (REFERENCE / "src" / "models" / "tts_model.py").exists()


# ---- seeded synthetic inputs ------------------------------------------------------
def c2_inputs():
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, 256, (16, 64), generator=g)
    lengths = torch.randint(32, 65, (16,), generator=g)
    dur = torch.randint(1, 9, (16, 64), generator=g).float()
    return ids, lengths, dur


def small_inputs(B: int, S: int, vocab: int, seed: int, lo: float = 1.0, hi: float = 3.0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, S), generator=g)
    lengths = torch.randint(max(1, S // 2), S + 1, (B,), generator=g)
    dur = torch.rand((B, S), generator=g) * (hi - lo) + lo
    return ids, lengths, dur


def lr_edge_durations(seed: int = 7, B: int = 6, S: int = 24) -> torch.Tensor:
    """Durations in [-1, 5) with the edge cases the reference's loop handles: negatives,
    fractions below one, an all-zero utterance, exact integers."""
    g = torch.Generator().manual_seed(seed)
    d = torch.rand((B, S), generator=g) * 6.0 - 1.0
    d[1] = torch.rand((S,), generator=g) * 0.99          # all truncate to 0 -> one zero row
    d[2, ::2] = d[2, ::2].round()                         # exact integers
    d[3] = -d[3].abs()                                    # all negative
    d[4, 0] = 37.0                                        # one long phoneme
    return d


def max_abs(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float32)
    b = torch.as_tensor(b, dtype=torch.float32)
    return float((a - b).abs().max())


def rel_l2(a, b) -> float:
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return float((a - b).norm() / b.norm().clamp_min(1e-30))
