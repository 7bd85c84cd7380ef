"""PCM16 wav writer for synthesis output (the reference uses soundfile,
src/utils/audio.py:154-180; stdlib `wave` keeps the B200 path dependency-free)."""
from __future__ import annotations

import wave
from pathlib import Path
from typing import Union

import numpy as np
import torch


def save_audio(audio: Union[np.ndarray, torch.Tensor], output_path: Union[str, Path],
               sample_rate: int = 22050) -> None:
    if isinstance(audio, torch.Tensor):
        audio = audio.detach().cpu().numpy()
    audio = np.asarray(audio, dtype=np.float32)
    if audio.ndim > 1:
        audio = audio.squeeze()
    pcm = (np.clip(audio, -1.0, 1.0) * 32767.0).round().astype("<i2")
    with wave.open(str(output_path), "wb") as f:
        f.setnchannels(1)
        f.setsampwidth(2)
        f.setframerate(int(sample_rate))
        f.writeframes(pcm.tobytes())
