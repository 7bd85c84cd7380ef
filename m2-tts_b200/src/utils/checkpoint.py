"""Checkpoint loading for synthesis, the `load_model` of the reference's scripts/synthesize.py:24-55.

Reference checkpoints are `torch.save({'model_state_dict', 'config', 'step', ...})` where `config` is an OmegaConf
DictConfig pickled by the trainer (training/train.py:359). OmegaConf is not a dependency of this build: the file is
first read with `weights_only=True` (tensors, plain containers); if that refuses the pickled config object (and only
then), a second pass unpickles with an ALLOW-LIST unpickler: `omegaconf.*` classes become plain attribute dicts, tensor
rebuild helpers and plain containers resolve normally, every other global raises `UnpicklingError` — a checkpoint cannot
run code through this loader. The state_dict keys/shapes are the reference's, loaded strictly.
"""
from __future__ import annotations

import logging
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


# Globals the fallback unpickler will resolve besides the omegaconf stubs: what `torch.save` itself emits for tensors and
# plain containers. Anything else in the pickle stream (os.system, builtins.eval, ...) is refused: an untrusted checkpoint
# must not be able to run code just because it also contains an OmegaConf object.
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "set"), ("builtins", "frozenset"),
    ("builtins", "int"), ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"),
    ("builtins", "complex"), ("builtins", "slice"), ("builtins", "object"), ("builtins", "getattr"),
    ("typing", "Any"), ("enum", "Enum"), ("pathlib", "PosixPath"), ("pathlib", "Path"), ("pathlib", "PurePosixPath"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch", "Size"), ("torch", "device"), ("torch", "dtype"),
    ("torch._tensor", "_rebuild_from_type_v2"), ("torch.serialization", "_get_layout"),
    ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("numpy", "dtype"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy", "ndarray"),
}
_ALLOWED_TORCH_STORAGES = {"FloatStorage", "DoubleStorage", "HalfStorage", "BFloat16Storage", "LongStorage", "IntStorage",
                           "ShortStorage", "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage"}


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if module.split(".")[0] == "omegaconf":
            return _ConfigStub
        if (module, name) in _ALLOWED_GLOBALS or (module == "torch" and (name in _ALLOWED_TORCH_STORAGES or name.endswith("Tensor"))) \
                or (module == "torch" and name in ("float32", "float64", "float16", "bfloat16", "int64", "int32", "int16", "int8",
                                                    "uint8", "bool")):
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}, which the m2tts_b200 loader does not allow "
                                     "(only tensors, plain containers and OmegaConf config objects are read)")


class _PickleShim:
    """`pickle_module` for torch.load: an allow-list unpickler (tensors, containers, omegaconf stubs)."""
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
    except pickle.UnpicklingError as err:
        # weights_only refused a global: for reference checkpoints that is the pickled OmegaConf config. Second pass with the
        # allow-list unpickler (a corrupt or truncated file raises other exceptions and is NOT retried).
        if "omegaconf" not in str(err).lower():
            raise
        logging.getLogger(__name__).info("checkpoint %s holds a pickled OmegaConf config: reading it with the allow-list "
                                         "unpickler (omegaconf classes become plain dicts)", path)
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
