"""Synthetic training batches for the B200 build — the `DummyDataset` / `collate_fn` / `create_dataloader` surface of the
reference's ``src/data/dataset.py:231-354`` that ``training/train.py:25,218-246`` imports.

The reference module cannot even be imported as shipped (`from ..utils.audio import` beyond the top-level package,
SURVEY.md §0.1) and its `TTSDataset` is librosa/soundfile audio DSP, which is outside the synthesis hot path (SURVEY.md §8
"out of scope"). What the trainer needs to drive `M2TTSModel.forward` in train mode is the batch dictionary, and that is
restated here with the same keys, dtypes and padding rules:
``phoneme_ids [B,S] long, text_lengths [B] long, mel_specs [B,M,T] float, mel_lengths [B] long, durations [B,S] float, texts``.
"""
from __future__ import annotations

from typing import Any, Dict, List, Optional

import torch
from torch.utils.data import DataLoader, Dataset


class DummyDataset(Dataset):
    """Random samples with the reference's shapes (src/data/dataset.py:300-354): text length in [10, max_text_length), mel
    length in [50, max_mel_length), durations that sum to the mel length. `seed` (not in the reference, which never seeds
    anything) makes item `idx` reproducible."""

    def __init__(self, size: int = 100, max_text_length: int = 50, max_mel_length: int = 200, mel_dim: int = 64,
                 vocab_size: int = 256, seed: Optional[int] = None):
        self.size = size
        self.max_text_length = max_text_length
        self.max_mel_length = max_mel_length
        self.mel_dim = mel_dim
        self.vocab_size = vocab_size
        self.seed = seed

    def __len__(self) -> int:
        return self.size

    def __getitem__(self, idx: int) -> Dict[str, Any]:
        g = None if self.seed is None else torch.Generator().manual_seed(self.seed * 1000003 + idx)
        text_len = int(torch.randint(10, self.max_text_length, (1,), generator=g).item())
        mel_len = int(torch.randint(50, self.max_mel_length, (1,), generator=g).item())
        phoneme_ids = torch.randint(0, self.vocab_size, (text_len,), generator=g)
        mel_spec = torch.randn(self.mel_dim, mel_len, generator=g)
        durations = torch.rand(text_len, generator=g)
        durations = durations / durations.sum() * mel_len
        return {"phoneme_ids": phoneme_ids, "text_length": torch.LongTensor([text_len]), "mel_spec": mel_spec,
                "mel_length": torch.LongTensor([mel_len]), "durations": durations, "text": f"dummy_text_{idx}"}


def collate_fn(batch: List[Dict[str, Any]]) -> Dict[str, Any]:
    """Zero-pad to the longest text / mel of the batch (src/data/dataset.py:231-279)."""
    n = len(batch)
    s_max = max(item["phoneme_ids"].size(0) for item in batch)
    t_max = max(item["mel_spec"].size(1) for item in batch)
    mel_dim = batch[0]["mel_spec"].size(0)
    out = {"phoneme_ids": torch.zeros(n, s_max, dtype=torch.long), "text_lengths": torch.zeros(n, dtype=torch.long),
           "mel_specs": torch.zeros(n, mel_dim, t_max), "mel_lengths": torch.zeros(n, dtype=torch.long),
           "durations": torch.zeros(n, s_max), "texts": []}
    for i, item in enumerate(batch):
        s, t = item["phoneme_ids"].size(0), item["mel_spec"].size(1)
        out["phoneme_ids"][i, :s] = item["phoneme_ids"]
        out["text_lengths"][i] = item["text_length"]
        out["mel_specs"][i, :, :t] = item["mel_spec"]
        out["mel_lengths"][i] = item["mel_length"]
        out["durations"][i, :s] = item["durations"]
        out["texts"].append(item["text"])
    return out


def create_dataloader(dataset: Dataset, batch_size: int = 2, shuffle: bool = True, num_workers: int = 0,
                      pin_memory: bool = True, drop_last: bool = True) -> DataLoader:
    """src/data/dataset.py:282-297 with defaults for a CUDA box (pinned batches, in-process loading for the tiny items)."""
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=pin_memory,
                      drop_last=drop_last, collate_fn=collate_fn, persistent_workers=num_workers > 0)
