"""ctypes binding of the C-ABI library (include/m2tts_b200.h) for the B200 synthesis path.

The library is built in-tree by ``__graft_entry__.build()`` / ``m2-tts_b200/csrc/Makefile`` into
``m2-tts_b200/lib/libm2tts_b200.so``.  There is NO fallback: if the library is missing, or a
tensor is not a contiguous fp32 CUDA tensor, the eval-mode forward raises.

Besides the raw bindings this module holds the three pieces of host state the stage modules share:

* the **status word** (one int32 per device): the kernels OR ``M2TTS_ST_*`` bits into it when an operand of the
  16-bit split leaves the fp16 range or an embedding id is out of range (``include/m2tts_b200.h``, "Status word").
  ``guarded()`` runs a stage, reads the word, and on ``FP16_RANGE`` runs the stage again with the TF32 split (no range
  limit); ``BAD_ID`` raises ``IndexError`` like ``nn.Embedding``. Inside ``deferred_status()`` (throughput pipelines, CUDA
  graph capture) the read is postponed to ``check_status()`` and a range violation raises ``Fp16RangeError`` instead.
* the **weight-image cache**: ``packed_weights()`` keeps the output of ``m2tts_*_pack`` per module, keyed on every
  parameter's ``(data_ptr, _version)`` — ``load_state_dict``, ``.to()`` or an optimizer step invalidate it.
* grow-only **workspaces** per (device, stream, tag).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import os
import threading
import warnings
from pathlib import Path
from typing import Callable, Dict, Iterable, Optional, Tuple

import torch

_PKG_ROOT = Path(__file__).resolve().parents[2]          # .../m2-tts_b200
_LIB_PATH = Path(os.environ.get("M2TTS_B200_LIB", _PKG_ROOT / "lib" / "libm2tts_b200.so"))

NUM_STAGES = 19
STAGE_NAMES = ["embed", "pack", "ln_qkv", "attention", "out_proj", "ffn1", "ffn2", "ln_proj",
               "layernorm", "durpred", "lr_count", "lr_gather", "voc_in", "voc_up", "voc_res1",
               "voc_res2", "voc_out", "probe", "voc_fused"]

# include/m2tts_b200.h
ST_FP16_RANGE, ST_BAD_ID = 1, 2
# host-side extension of the word: the length regulator's own status bits (NaN, inf, int32 overflow) shifted by 2 when its
# host read is deferred
ST_LR_NAN, ST_LR_INF, ST_LR_OVERFLOW = 4, 8, 16
PREC_DEFAULT, PREC_SPLIT16, PREC_FFMA, PREC_TF32 = -1, 0, 1, 2
PRECISION_NAMES = {"default": PREC_DEFAULT, "split16": PREC_SPLIT16, "ffma": PREC_FFMA, "tf32": PREC_TF32}


class NativeLibraryError(RuntimeError):
    """The sm_100a extension is missing or a call into it failed."""


class Fp16RangeError(RuntimeError):
    """An operand of the 16-bit split left the fp16 range and the check was deferred: the results of the calls since
    the last check are invalid. Re-run them under ``precision("tf32")`` (outside ``deferred_status`` this happens
    automatically)."""


class LayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "norm1_w", "norm1_b", "qkv_w", "out_w", "out_b", "norm2_w", "norm2_b",
        "ffn1_w", "ffn1_b", "ffn2_w", "ffn2_b")]


class DurPredWeights(C.Structure):
    _fields_ = [("conv_w", C.c_void_p * 2), ("conv_b", C.c_void_p * 2),
                ("bn_w", C.c_void_p * 2), ("bn_b", C.c_void_p * 2),
                ("bn_mean", C.c_void_p * 2), ("bn_var", C.c_void_p * 2),
                ("proj_w", C.c_void_p), ("proj_b", C.c_void_p), ("bn_eps", C.c_float)]


class VocoderWeights(C.Structure):
    _fields_ = [("in_w", C.c_void_p), ("in_b", C.c_void_p),
                ("up_w", C.c_void_p * 4), ("up_b", C.c_void_p * 4),
                ("res1_w", C.c_void_p * 4), ("res1_b", C.c_void_p * 4),
                ("res2_w", C.c_void_p * 4), ("res2_b", C.c_void_p * 4),
                ("out_w", C.c_void_p), ("out_b", C.c_void_p),
                ("res_dilation", C.c_int * 4)]


_P, _I, _SZ = C.c_void_p, C.c_int, C.c_size_t
_SIGNATURES = {
    "m2tts_version": (_I, []),
    "m2tts_last_error_string": (C.c_char_p, []),
    "m2tts_launch_count": (C.c_uint64, []),
    "m2tts_stage_timing_enable": (_I, [_I]),
    "m2tts_stage_timing_read": (_I, [C.POINTER(C.c_float), C.POINTER(_I), _I]),
    "m2tts_debug_words": (_I, [C.POINTER(_I), _I]),
    "m2tts_ffma_probe": (_I, [_P, _I, C.POINTER(C.c_double), _P]),
    "m2tts_embed_posenc": (_I, [_P] * 6 + [_I] * 4 + [_P, _P]),
    "m2tts_transformer_workspace_bytes": (_SZ, [_I] * 4),
    "m2tts_transformer_pack_bytes": (_SZ, [_I] * 3),
    "m2tts_transformer_pack": (_I, [C.POINTER(LayerWeights), _I, _I, _I, _P, _SZ, _P, _P]),
    "m2tts_transformer_layer": (_I, [C.POINTER(LayerWeights), _P, _P, _P, _P, _I, _I, _I, _I, _I, C.c_float, _I, _P,
                                     _P, _SZ, _P]),
    "m2tts_layernorm": (_I, [_P] * 4 + [_I, _I, C.c_float, _P]),
    "m2tts_ln_proj_workspace_bytes": (_SZ, [_I, _I]),
    "m2tts_ln_proj_rows_workspace_bytes": (_SZ, [_I, _I, _I]),
    "m2tts_ln_proj_pack_bytes": (_SZ, [_I] * 3),
    "m2tts_ln_proj_pack": (_I, [_P, _I, _I, _I, _P, _SZ, _P, _P]),
    "m2tts_layernorm_proj": (_I, [_P] * 7 + [_I, _I, _I, C.c_float, _I, _P, _P, _SZ, _P]),
    "m2tts_duration_predictor": (_I, [C.POINTER(DurPredWeights), _P, _P, _I, _I, _I, _P]),
    "m2tts_length_regulate_count": (_I, [_P, _I, _I] + [_P] * 5),
    "m2tts_length_regulate_gather": (_I, [_P] * 5 + [_I] * 4 + [_P]),
    "m2tts_vocoder_workspace_bytes": (_SZ, [_I] * 4),
    "m2tts_vocoder_pack_bytes": (_SZ, [_I] * 3),
    "m2tts_vocoder_pack": (_I, [C.POINTER(VocoderWeights), _I, _I, _I, _P, _SZ, _P, _P]),
    "m2tts_vocoder_forward": (_I, [C.POINTER(VocoderWeights), _P, _P, C.c_int64, C.c_int64, C.c_int64, _P,
                                   _I, _I, _I, _I, _I, _P, _P, _SZ, _P]),
    "m2tts_conv_workspace_bytes": (_SZ, [_I] * 3),
    "m2tts_conv1d_k3": (_I, [_P, C.c_int64, C.c_int64, C.c_int64, _P, _P, _P, _P] + [_I] * 6 + [_P, _SZ, _P]),
    "m2tts_conv_transpose1d_lrelu": (_I, [_P] * 4 + [_I] * 5 + [_P]),
    "m2tts_conv_tc_workspace_bytes": (_SZ, [_I] * 5),
    "m2tts_conv1d_k3_tc": (_I, [_P] * 5 + [_I] * 6 + [_P, _SZ, _P]),
    "m2tts_conv_transpose1d_lrelu_tc": (_I, [_P] * 4 + [_I] * 5 + [_P, _SZ, _P]),
    "m2tts_vocoder_stage_fused_workspace_bytes": (_SZ, [_I]),
    "m2tts_vocoder_stage_fused": (_I, [_P] * 10 + [_I] * 3 + [_P, _SZ, _P]),
    "m2tts_vocoder_plan": (_I, [_I, _I, _I, _P, _P]),
    "m2tts_vocoder_stage_fused_h_workspace_bytes": (_SZ, [_I] * 3),
    "m2tts_vocoder_stage_fused_h": (_I, [_P] * 10 + [_I] * 3 + [_P, _P, _SZ, _P]),
    "m2tts_conv1d_k3_h_workspace_bytes": (_SZ, [_I] * 3),
    "m2tts_conv1d_k3_h": (_I, [_P] * 5 + [_I] * 5 + [_P, _P, _SZ, _P]),
    "m2tts_conv_transpose_x4_h_workspace_bytes": (_SZ, [_I] * 3),
    "m2tts_conv_transpose_x4_h": (_I, [_P] * 4 + [_I] * 3 + [_P, _P, _SZ, _P]),
    "m2tts_resblock_fused_h_workspace_bytes": (_SZ, [_I] * 3),
    "m2tts_resblock_fused_h": (_I, [_P] * 6 + [_I] * 3 + [_P, _P, _SZ, _P]),
    "m2tts_pcm16": (_I, [_P, _P, C.c_longlong, _P]),
}
# include/m2tts_b200_tools.h — only in libm2tts_b200_tools.so (M2TTS_B200_LIB=.../libm2tts_b200_tools.so)
_TOOL_SIGNATURES = {
    "m2tts_attention_set_prof": (_I, [_P]),
    "m2tts_vocoder_stage_fused_set_prof": (_I, [_P]),
    "m2tts_tapgemm_set_prof": (_I, [_P]),
    "m2tts_voc_up_h_set_debug": (_I, [_I]),
    "m2tts_attention_tc_planes": (_I, [_P, _P, _P] + [_I] * 5 + [_P, _P, _P]),
    "m2tts_rowshift_probe": (_I, [_P] * 3 + [_I] * 6 + [_P]),
    "m2tts_mma_bench": (_I, [_I] * 5 + [_P, _P]),
    "m2tts_umma_probe": (_I, [_P] * 3 + [_I, _I, _P, _P, _P]),
    "m2tts_umma_probe_f16": (_I, [_P] * 3 + [_I] * 3 + [_P]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
TOOL_SYMBOLS = tuple(_TOOL_SIGNATURES)

_lib: Optional[C.CDLL] = None
_has_tools = False


def library_path() -> Path:
    return _LIB_PATH


def tools_library_path() -> Path:
    return _PKG_ROOT / "lib" / "libm2tts_b200_tools.so"


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise NativeLibraryError if it is absent."""
    global _lib, _has_tools
    if _lib is None:
        if not _LIB_PATH.exists():
            raise NativeLibraryError(
                f"m2tts_b200: CUDA extension not built ({_LIB_PATH} missing). Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C m2-tts_b200/csrc`. "
                "There is no CPU/PyTorch fallback for the eval-mode synthesis path.")
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _has_tools = hasattr(handle, "m2tts_attention_set_prof")
        if _has_tools:
            for name, (res, args) in _TOOL_SIGNATURES.items():
                fn = getattr(handle, name)
                fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def tools_lib() -> C.CDLL:
    """The loaded library, checked to be the tools build (bring-up scripts under tools/)."""
    handle = lib()
    if not _has_tools:
        raise NativeLibraryError(
            f"{_LIB_PATH} is the product library; the bring-up hooks live in {tools_library_path()} "
            "(`make -C m2-tts_b200/csrc tools`, then M2TTS_B200_LIB=<that path>)")
    return handle


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    msg = lib().m2tts_last_error_string().decode("utf-8", "replace")
    if rc in (-1, -2):  # bad shape / unsupported dimension
        raise ValueError(f"m2tts_b200 {what}: {msg}")
    raise NativeLibraryError(f"m2tts_b200 {what} failed (code {rc}): {msg}")


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, name: str, dtype: torch.dtype = torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise NativeLibraryError(
            f"m2tts_b200: `{name}` lives on {t.device}; the eval-mode synthesis path runs only on "
            "CUDA (sm_100a) and has no CPU/MPS fallback — move the model and inputs to a B200.")
    if t.dtype != dtype:
        raise TypeError(f"m2tts_b200: `{name}` must be {dtype}, got {t.dtype}")
    return t


def weight(t: torch.Tensor, name: str) -> int:
    """Pointer to a parameter/buffer after checking it is contiguous fp32 CUDA memory."""
    require_cuda(t, name)
    if not t.is_contiguous():
        raise ValueError(f"m2tts_b200: parameter `{name}` must be contiguous")
    return t.data_ptr()


def stream_handle(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def on_device(device: torch.device):
    """Context manager making `device` the current CUDA device: the library launches on the current device, so a model
    on cuda:1 must not run with cuda:0 current (invalid stream handle, wrong-device function attributes)."""
    return torch.cuda.device(device)


_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}


def _dev_index(device: torch.device) -> int:
    return device.index if device.index is not None else torch.cuda.current_device()


def workspace(device: torch.device, nbytes: int, tag: str = "main") -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream, tag); the library never keeps pointers."""
    key = (_dev_index(device), stream_handle(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


# ---- per-call precision -------------------------------------------------------------------------
_tls = threading.local()


def current_precision() -> int:
    return getattr(_tls, "precision", PREC_DEFAULT)


@contextlib.contextmanager
def precision(p):
    """Run the enclosed eval-mode calls with an explicit `precision` argument ("split16" | "tf32" | "ffma" or the
    M2TTS_PREC_* value). Thread-local; the library itself has no mode state."""
    val = PRECISION_NAMES[p] if isinstance(p, str) else int(p)
    old = current_precision()
    _tls.precision = val
    try:
        yield
    finally:
        _tls.precision = old


# ---- status word --------------------------------------------------------------------------------
_status: Dict[int, torch.Tensor] = {}
_range_warned = False


def status_word(device: torch.device) -> torch.Tensor:
    idx = _dev_index(device)
    t = _status.get(idx)
    if t is None:
        t = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", idx))
        _status[idx] = t
    return t


def status_ptr(device: torch.device) -> int:
    return status_word(device).data_ptr()


def read_status(device: torch.device) -> int:
    """Read and clear the device status word (synchronises the current stream)."""
    t = status_word(device)
    v = int(t.item())
    if v:
        t.zero_()
    return v


def _deferred_depth() -> int:
    return getattr(_tls, "deferred", 0)


@contextlib.contextmanager
def deferred_status():
    """Postpone the status read of the enclosed eval-mode calls (no host synchronisation per stage: throughput pipelines,
    CUDA-graph capture). The caller must call ``check_status(device)`` once the work has been synchronised."""
    _tls.deferred = _deferred_depth() + 1
    try:
        yield
    finally:
        _tls.deferred -= 1


def status_deferred() -> bool:
    return _deferred_depth() > 0


def raise_for_status(flags: int, what: str) -> None:
    if flags & ST_LR_NAN:
        raise ValueError("cannot convert float NaN to integer")
    if flags & ST_LR_INF:
        raise OverflowError("cannot convert float infinity to integer")
    if flags & ST_LR_OVERFLOW:
        raise OverflowError("length regulator: frame count exceeds int32")
    if flags & ST_BAD_ID:
        raise IndexError(f"m2tts_b200 {what}: index out of range in self (a phoneme id is outside the embedding table)")
    if flags & ST_FP16_RANGE:
        raise Fp16RangeError(f"m2tts_b200 {what}: an operand of the 16-bit split left the fp16 range; the results are "
                             "invalid — re-run under models._native.precision('tf32')")


def check_status(device: torch.device, what: str = "deferred status check") -> None:
    """Read the status word after a ``deferred_status`` region; raises IndexError / Fp16RangeError."""
    raise_for_status(read_status(device), what)


def guarded(device: torch.device, run: Callable[[int], object], what: str, owner=None):
    """Run ``run(precision)`` and validate it against the status word: on FP16_RANGE with the default precision the stage
    runs again with the TF32 split (same result contract, no operand range limit) and `owner` (the stage module) remembers
    to start there next time — its cached fp16 weight images may hold infinities. ``drop_packed`` forgets that."""
    global _range_warned
    prec = current_precision()
    if prec == PREC_DEFAULT and owner is not None and owner.__dict__.get("_m2tts_tf32_only", False):
        prec = PREC_TF32
    out = run(prec)
    if _deferred_depth() > 0:
        return out
    flags = read_status(device)
    if flags & ST_FP16_RANGE and prec in (PREC_DEFAULT, PREC_SPLIT16):
        if not _range_warned:
            warnings.warn(f"m2tts_b200 {what}: activations or weights exceed the fp16 range of the 16-bit split; "
                          "re-running with the TF32 split (slower, no range limit)", RuntimeWarning, stacklevel=3)
            _range_warned = True
        out = run(PREC_TF32)
        flags = (flags & ~ST_FP16_RANGE) | read_status(device)
        if owner is not None:
            owner.__dict__["_m2tts_tf32_only"] = True
    raise_for_status(flags, what)
    return out


# ---- weight-image cache --------------------------------------------------------------------------
def params_key(tensors: Iterable[torch.Tensor]) -> Tuple:
    return tuple((t.data_ptr(), t._version) for t in tensors)


def packed_weights(owner, slot: str, key: Tuple, nbytes: int, pack: Callable[[torch.Tensor], None],
                   device: torch.device) -> torch.Tensor:
    """Cached output of an ``m2tts_*_pack`` call, stored on the owning module (not a registered buffer: it is derived
    state and must not enter the state_dict). `key` = parameter identities/versions + shapes + precision."""
    cache = owner.__dict__.setdefault("_m2tts_packed", {})
    hit = cache.get(slot)
    if hit is not None and hit[0] == key and hit[1].device == device:
        return hit[1]
    buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
    pack(buf)
    cache[slot] = (key, buf)
    return buf


def drop_packed(module: torch.nn.Module) -> None:
    for m in module.modules():
        m.__dict__.pop("_m2tts_packed", None)
        m.__dict__.pop("_m2tts_tf32_only", None)


# ---- stage timers / launch counter ------------------------------------------------------------
def launch_count() -> int:
    return int(lib().m2tts_launch_count())


def stage_timing_enable(on: bool) -> None:
    check(lib().m2tts_stage_timing_enable(1 if on else 0), "stage_timing_enable")


def stage_timing_read() -> Dict[str, Tuple[float, int]]:
    ms = (C.c_float * NUM_STAGES)()
    cnt = (C.c_int * NUM_STAGES)()
    check(lib().m2tts_stage_timing_read(ms, cnt, NUM_STAGES), "stage_timing_read")
    return {STAGE_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(NUM_STAGES) if cnt[i] > 0}
