#!/usr/bin/env python
"""Train-mode drop-in check: the loss and the training step of the reference's ``training/train.py`` (TTSLoss :48-105,
train_step :290-342) driven against the B200 mirror model on a CUDA device.

In train mode the mirror runs the reference's formulation in differentiable torch ops on the same parameters (dropout,
activation checkpointing, no vocoder in forward; `models/tts_model.py`), so this is plain PyTorch autograd on the GPU; the
hand-written kernels serve `model.eval()`. The trainer's surroundings (wandb, thermal monitor, checkpoint rotation,
validation audio) are outside the hot path and not rebuilt.

    python m2-tts_b200/training/train.py --stage stage1 --steps 20 --batch-size 8
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path
from typing import Any, Dict, Optional

import torch
import torch.nn as nn

SRC = Path(__file__).resolve().parents[1] / "src"
if str(SRC) not in sys.path:
    sys.path.insert(0, str(SRC))


class TTSLoss(nn.Module):
    """L1 on the valid mel frames of every utterance (mean per utterance, then over the batch) + MSE on the durations
    (training/train.py:48-105). The reference loops over the batch with `.item()`; the masked form below is the same sum
    without B host synchronisations."""

    def __init__(self, mel_loss_weight: float = 1.0, duration_loss_weight: float = 0.1):
        super().__init__()
        self.mel_loss_weight = mel_loss_weight
        self.duration_loss_weight = duration_loss_weight

    def forward(self, mel_pred: torch.Tensor, mel_target: torch.Tensor, duration_pred: torch.Tensor,
                duration_target: torch.Tensor, mel_lengths: torch.Tensor) -> Dict[str, torch.Tensor]:
        mel_target = mel_target.transpose(1, 2)                                   # [B, T, M] like the prediction
        B, T, M = mel_pred.shape
        lens = mel_lengths.reshape(B).clamp(max=T)
        mask = (torch.arange(T, device=mel_pred.device)[None, :] < lens[:, None]).to(mel_pred.dtype)
        per_utt = ((mel_pred - mel_target).abs() * mask[:, :, None]).sum(dim=(1, 2)) / (lens.to(mel_pred.dtype) * M)
        mel_loss = per_utt.mean()
        duration_loss = nn.functional.mse_loss(duration_pred, duration_target)
        total = self.mel_loss_weight * mel_loss + self.duration_loss_weight * duration_loss
        return {"total_loss": total, "mel_loss": mel_loss, "duration_loss": duration_loss}


def train_step(model: nn.Module, batch: Dict[str, Any], criterion: TTSLoss, optimizer: torch.optim.Optimizer,
               device: torch.device, gradient_clip_norm: Optional[float] = 1.0) -> Dict[str, float]:
    """training/train.py:290-342: train-mode forward with teacher-forced durations, loss, backward, clip, step."""
    model.train()
    batch = {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}
    out = model(phoneme_ids=batch["phoneme_ids"], phoneme_lengths=batch["text_lengths"],
                target_durations=batch["durations"], max_target_length=batch["mel_specs"].size(2))
    losses = criterion(mel_pred=out["mel_output"], mel_target=batch["mel_specs"], duration_pred=out["duration_pred"],
                       duration_target=batch["durations"], mel_lengths=batch["mel_lengths"])
    optimizer.zero_grad()
    losses["total_loss"].backward()
    if gradient_clip_norm:
        torch.nn.utils.clip_grad_norm_(model.parameters(), gradient_clip_norm)
    optimizer.step()
    return {k: float(v.item()) for k, v in losses.items()}


def main(argv=None) -> int:
    from data.dataset import DummyDataset, create_dataloader
    from models.stage_configs import STAGE_KWARGS
    from models.tts_model import M2TTSModel
    from utils.device import setup_device

    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", default="stage1", choices=sorted(STAGE_KWARGS))
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch-size", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--seed", type=int, default=1234)
    args = ap.parse_args(argv)
    device = setup_device()
    torch.manual_seed(args.seed)
    kw = STAGE_KWARGS[args.stage]
    model = M2TTSModel(**kw).to(device)
    data = DummyDataset(size=args.steps * args.batch_size, mel_dim=kw["mel_channels"], vocab_size=kw["vocab_size"], seed=args.seed)
    loader = create_dataloader(data, batch_size=args.batch_size, shuffle=False)
    criterion, optimizer = TTSLoss(), torch.optim.AdamW(model.parameters(), lr=args.lr)
    for step, batch in enumerate(loader):
        m = train_step(model, batch, criterion, optimizer, device)
        print(f"step {step:4d}  total {m['total_loss']:.4f}  mel {m['mel_loss']:.4f}  duration {m['duration_loss']:.4f}", flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
