"""ctypes binding of the C-ABI library (include/m2tts_b200.h) for the B200 synthesis path.

The library is built in-tree by ``__graft_entry__.build()`` / ``m2-tts_b200/csrc/Makefile`` into
``m2-tts_b200/lib/libm2tts_b200.so``.  There is NO fallback: if the library is missing, or a
tensor is not a contiguous fp32 CUDA tensor, the eval-mode forward raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Dict, Optional, Tuple

import torch

_PKG_ROOT = Path(__file__).resolve().parents[2]          # .../m2-tts_b200
_LIB_PATH = Path(os.environ.get("M2TTS_B200_LIB", _PKG_ROOT / "lib" / "libm2tts_b200.so"))

NUM_STAGES = 19
STAGE_NAMES = ["embed", "pack", "ln_qkv", "attention", "out_proj", "ffn1", "ffn2", "ln_proj",
               "layernorm", "durpred", "lr_count", "lr_gather", "voc_in", "voc_up", "voc_res1",
               "voc_res2", "voc_out", "probe", "voc_fused"]


class NativeLibraryError(RuntimeError):
    """The sm_100a extension is missing or a call into it failed."""


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


_SIGNATURES = {
    "m2tts_version": (C.c_int, []),
    "m2tts_last_error_string": (C.c_char_p, []),
    "m2tts_launch_count": (C.c_uint64, []),
    "m2tts_stage_timing_enable": (C.c_int, [C.c_int]),
    "m2tts_stage_timing_read": (C.c_int, [C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_int]),
    "m2tts_debug_words": (C.c_int, [C.POINTER(C.c_int), C.c_int]),
    "m2tts_set_attention_mode": (C.c_int, [C.c_int]),
    "m2tts_set_vocoder_mode": (C.c_int, [C.c_int]),
    "m2tts_ffma_probe": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "m2tts_embed_posenc": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p]),
    "m2tts_transformer_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "m2tts_transformer_layer": (C.c_int, [C.POINTER(LayerWeights), C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                          C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_layernorm": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "m2tts_ln_proj_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "m2tts_ln_proj_rows_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "m2tts_layernorm_proj": (C.c_int, [C.c_void_p] * 6 + [C.c_int, C.c_int, C.c_int, C.c_float,
                                                           C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_duration_predictor": (C.c_int, [C.POINTER(DurPredWeights), C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "m2tts_length_regulate_count": (C.c_int, [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5),
    "m2tts_length_regulate_gather": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]),
    "m2tts_vocoder_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "m2tts_vocoder_forward": (C.c_int, [C.POINTER(VocoderWeights), C.c_void_p, C.c_int64, C.c_int64,
                                        C.c_int64, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_conv_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "m2tts_conv1d_k3": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_size_t,
                                                                              C.c_void_p]),
    "m2tts_conv_transpose1d_lrelu": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p]),
    "m2tts_conv_tc_workspace_bytes": (C.c_size_t, [C.c_int] * 5),
    "m2tts_conv1d_k3_tc": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 6 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_conv_transpose1d_lrelu_tc": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 5 + [C.c_void_p, C.c_size_t,
                                                                                      C.c_void_p]),
    "m2tts_vocoder_stage_fused_workspace_bytes": (C.c_size_t, [C.c_int]),
    "m2tts_vocoder_stage_fused": (C.c_int, [C.c_void_p] * 10 + [C.c_int] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_vocoder_stage_fused_h_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "m2tts_vocoder_stage_fused_h": (C.c_int, [C.c_void_p] * 10 + [C.c_int] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_conv1d_k3_h_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "m2tts_conv1d_k3_h": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_conv_transpose_x4_h_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "m2tts_conv_transpose_x4_h": (C.c_int, [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_voc_up_h_set_debug": (C.c_int, [C.c_int]),
    "m2tts_resblock_fused_h_workspace_bytes": (C.c_size_t, [C.c_int] * 3),
    "m2tts_resblock_fused_h": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    "m2tts_tapgemm_set_prof": (C.c_int, [C.c_void_p]),
    "m2tts_rowshift_probe": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]),
    "m2tts_mma_bench": (C.c_int, [C.c_int] * 5 + [C.c_void_p, C.c_void_p]),
    "m2tts_umma_probe_f16": (C.c_int, [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p]),
    "m2tts_pcm16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "m2tts_vocoder_stage_fused_set_prof": (C.c_int, [C.c_void_p]),
    "m2tts_attention_set_prof": (C.c_int, [C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib: Optional[C.CDLL] = None


def library_path() -> Path:
    return _LIB_PATH


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise NativeLibraryError if it is absent."""
    global _lib
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
        _lib = handle
    return _lib


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


_workspaces: Dict[Tuple[int, int, str], torch.Tensor] = {}


def workspace(device: torch.device, nbytes: int, tag: str = "main") -> torch.Tensor:
    """Grow-only scratch buffer per (device, stream, tag); the library never keeps pointers."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           stream_handle(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    _workspaces.clear()


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
