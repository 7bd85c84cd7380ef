"""Checkpoint loading for synthesis, the `load_model` of the reference's scripts/synthesize.py:24-55.

Reference checkpoints are `torch.save({'model_state_dict', 'config', 'step', ...})` where `config` is an OmegaConf
DictConfig pickled by the trainer (training/train.py:359). OmegaConf is not a dependency of this build: the file is
first read with `weights_only=True` (tensors, plain containers); if that refuses the pickled config object, a second
pass unpickles with a stub that turns every `omegaconf.*` class into a plain attribute dict, which is all
`model_kwargs` needs. The state_dict keys/shapes are the reference's, loaded strictly.
"""
from __future__ import annotations

import pickle
from pathlib import Path
from typing import Any, Optional, Tuple, Union

import torch

from utils.config import AttrDict, model_kwargs


class _ConfigStub(dict):
    """Stands in for omegaconf container classes while unpickling; keeps whatever state they carried."""

    def __init__(self, *a, **k):
        super().__init__()

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.update(state)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module.split(".")[0] == "omegaconf":
            return _ConfigStub
        return super().find_class(module, name)


class _PickleShim:
    """`pickle_module` for torch.load: everything stock except omegaconf classes."""
    __name__ = "pickle"
    Unpickler = _Unpickler
    load = staticmethod(lambda f, **kw: _Unpickler(f, **kw).load())
    loads = staticmethod(pickle.loads)
    dumps = staticmethod(pickle.dumps)
    dump = staticmethod(pickle.dump)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    PickleError, UnpicklingError, PicklingError = pickle.PickleError, pickle.UnpicklingError, pickle.PicklingError


def _plain(obj: Any) -> Any:
    """OmegaConf stubs keep their payload under '_content' (nodes under '_val'); unwrap to plain containers."""
    if isinstance(obj, _ConfigStub):
        if "_content" in obj:
            return _plain(obj["_content"])
        if "_val" in obj:
            return _plain(obj["_val"])
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, dict):
        return {k: _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    return obj


def read_checkpoint(path: Union[str, Path], map_location="cpu") -> dict:
    path = Path(path)
    if not path.exists():
        raise FileNotFoundError(f"Checkpoint not found: {path}")          # scripts/synthesize.py:26-27
    try:
        return torch.load(path, map_location=map_location, weights_only=True)
    except Exception:
        ckpt = torch.load(path, map_location=map_location, weights_only=False, pickle_module=_PickleShim)
        if "config" in ckpt:
            ckpt["config"] = _plain(ckpt["config"])
        return ckpt


def load_model(checkpoint_path: Union[str, Path], device: torch.device, config: Optional[Any] = None) -> Tuple[Any, dict]:
    """-> (M2TTSModel in eval mode on `device`, checkpoint dict). `config` overrides the one in the file; with neither,
    the model's default constructor arguments are used, as the reference does (scripts/synthesize.py:32-35)."""
    from models.tts_model import M2TTSModel
    ckpt = read_checkpoint(checkpoint_path)
    cfg = config if config is not None else ckpt.get("config")
    if cfg is None:
        model = M2TTSModel()
    else:
        model = M2TTSModel(**model_kwargs(AttrDict.wrap(cfg) if isinstance(cfg, dict) else cfg))
    model.load_state_dict(ckpt["model_state_dict"])
    model.to(device).eval()
    return model, ckpt
