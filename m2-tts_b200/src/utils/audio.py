"""PCM16 wav writer for synthesis output (the reference uses soundfile, src/utils/audio.py:154-180; stdlib `wave`
keeps the B200 path dependency-free). CUDA tensors are converted to int16 ON THE DEVICE (`m2tts_pcm16`) and copied
through pinned memory, so 2 bytes per sample cross PCIe instead of 4."""
from __future__ import annotations

import wave
from pathlib import Path
from typing import Union

import numpy as np
import torch


def to_pcm16(audio: Union[np.ndarray, torch.Tensor]) -> np.ndarray:
    """float waveform -> little-endian int16 samples: clip to [-1, 1], scale by 32767, round half to even."""
    if isinstance(audio, torch.Tensor) and audio.is_cuda:
        from models import _native as nat
        x = audio.detach().to(torch.float32).contiguous().reshape(-1)
        if x.numel() == 0:
            return np.zeros((0,), dtype="<i2")
        if x.data_ptr() % 16:
            x = x.clone()
        pcm = torch.empty(x.numel(), dtype=torch.int16, device=x.device)
        nat.check(nat.lib().m2tts_pcm16(x.data_ptr(), pcm.data_ptr(), x.numel(), nat.stream_handle(x.device)), "pcm16")
        host = torch.empty(x.numel(), dtype=torch.int16, pin_memory=True)
        host.copy_(pcm, non_blocking=True)
        torch.cuda.current_stream(x.device).synchronize()
        return host.numpy().astype("<i2", copy=False)
    if isinstance(audio, torch.Tensor):
        audio = audio.detach().cpu().numpy()
    audio = np.asarray(audio, dtype=np.float32).reshape(-1)
    return (np.clip(audio, -1.0, 1.0) * np.float32(32767.0)).round().astype("<i2")


def save_audio(audio: Union[np.ndarray, torch.Tensor], output_path: Union[str, Path],
               sample_rate: int = 22050) -> None:
    """Same call as the reference's `save_audio(audio, path, sample_rate)`; any shape that squeezes to 1-D."""
    pcm = to_pcm16(audio)
    with wave.open(str(output_path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.tobytes())
