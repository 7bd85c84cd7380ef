"""Device selection for the B200 build — replaces the reference's mps/cpu picker
(reference src/utils/device.py:13-36, which never returns a CUDA device)."""
from __future__ import annotations

import os

import torch


def setup_device() -> torch.device:
    """cuda:LOCAL_RANK (one process per GPU). Raises when no CUDA device is visible: the
    eval-mode synthesis path has no CPU/MPS fallback."""
    if not torch.cuda.is_available():
        raise RuntimeError("m2tts_b200 needs a CUDA device (B200, sm_100a); none is visible")
    index = int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count()
    torch.cuda.set_device(index)
    return torch.device("cuda", index)


def get_device_info() -> dict:
    info = {"device_type": "cuda" if torch.cuda.is_available() else "none",
            "cuda_available": torch.cuda.is_available()}
    if torch.cuda.is_available():
        p = torch.cuda.get_device_properties(torch.cuda.current_device())
        info.update(name=p.name, total_memory_gb=round(p.total_memory / 2 ** 30, 2),
                    sm_count=p.multi_processor_count, capability=f"{p.major}.{p.minor}")
    return info


def clear_cache() -> None:
    if torch.cuda.is_available():
        torch.cuda.empty_cache()
