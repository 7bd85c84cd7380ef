"""Building blocks of the m2-tts model — B200 host-side mirror of the reference's
``src/models/components.py`` (same class names, constructor signatures, attribute names and
therefore the same ``state_dict`` keys).

These classes are parameter containers plus the TRAIN-mode (autograd) formulation in plain
torch ops.  The eval-mode synthesis path never runs them: the stage modules in
``tts_model.py`` hand the parameters to the sm_100a kernels behind ``include/m2tts_b200.h``.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint


class PositionalEncoding(nn.Module):
    """Sinusoidal table kept as the persistent buffer ``pe`` [1, max_length, H]
    (reference components.py:15-39; the buffer is part of the state_dict)."""

    def __init__(self, hidden_dim: int, max_length: int = 5000):
        super().__init__()
        pos = torch.arange(0, max_length).unsqueeze(1).float()
        freq = torch.exp(torch.arange(0, hidden_dim, 2).float() * -(math.log(10000.0) / hidden_dim))
        table = torch.zeros(max_length, hidden_dim)
        table[:, 0::2] = torch.sin(pos * freq)
        table[:, 1::2] = torch.cos(pos * freq)
        self.register_buffer("pe", table.unsqueeze(0))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.pe[:, : x.size(1)]


class MultiHeadAttention(nn.Module):
    """Fused-QKV self attention (reference components.py:42-90). qkv has no bias and its output
    rows are ordered [3, heads, head_dim]; only KEYS are masked, with the finite value -1e9."""

    def __init__(self, hidden_dim: int, num_heads: int, dropout: float = 0.1):
        super().__init__()
        assert hidden_dim % num_heads == 0
        self.hidden_dim = hidden_dim
        self.num_heads = num_heads
        self.head_dim = hidden_dim // num_heads
        self.scale = 1.0 / math.sqrt(self.head_dim)
        self.qkv = nn.Linear(hidden_dim, hidden_dim * 3, bias=False)
        self.out_proj = nn.Linear(hidden_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        B, L, _ = x.shape
        q, k, v = self.qkv(x).view(B, L, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        scores = (q @ k.transpose(-2, -1)) * self.scale
        if mask is not None:
            scores = scores.masked_fill(~mask.bool()[:, None, None, :], -1e9)
        attn = self.dropout(F.softmax(scores, dim=-1))
        ctx = (attn @ v).transpose(1, 2).reshape(B, L, self.hidden_dim)
        return self.out_proj(ctx)


class FeedForward(nn.Module):
    """linear2(dropout(relu(linear1(x)))) (reference components.py:93-103)."""

    def __init__(self, hidden_dim: int, ffn_dim: int, dropout: float = 0.1):
        super().__init__()
        self.linear1 = nn.Linear(hidden_dim, ffn_dim)
        self.linear2 = nn.Linear(ffn_dim, hidden_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.linear2(self.dropout(F.relu(self.linear1(x))))


class TransformerEncoderLayer(nn.Module):
    """Pre-LN layer (reference components.py:106-140). In eval mode on CUDA the whole layer is
    one C-ABI call (``m2tts_transformer_layer``); the torch formulation below serves training
    (dropout, activation checkpointing, autograd)."""

    def __init__(self, hidden_dim: int, num_heads: int, ffn_dim: int, dropout: float = 0.1,
                 use_checkpointing: bool = True):
        super().__init__()
        self.self_attn = MultiHeadAttention(hidden_dim, num_heads, dropout)
        self.ffn = FeedForward(hidden_dim, ffn_dim, dropout)
        self.norm1 = nn.LayerNorm(hidden_dim)
        self.norm2 = nn.LayerNorm(hidden_dim)
        self.dropout = nn.Dropout(dropout)
        self.use_checkpointing = use_checkpointing

    def forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if not self.training:
            from . import _native as nat
            from .tts_model import _native_layer_stack  # local import: avoids a cycle
            lengths = None
            if mask is not None:
                # the kernel masks keys >= length; a prefix mask is the only kind the model builds
                lengths = mask.to(torch.int64).sum(dim=1)
            nat.require_cuda(x, "x")
            out = torch.empty_like(x, memory_format=torch.contiguous_format)
            with nat.on_device(x.device):
                nat.guarded(x.device, lambda prec: _native_layer_stack([self], x, lengths, out=out, prec=prec),
                            "transformer_layer", owner=self)
            return out
        if self.use_checkpointing:
            return checkpoint(self._forward, x, mask, use_reentrant=False)
        return self._forward(x, mask)

    def _forward(self, x: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = x + self.dropout(self.self_attn(self.norm1(x), mask))
        return x + self.dropout(self.ffn(self.norm2(x)))


class ConvBlock(nn.Module):
    """Conv1d(k, pad=k//2) -> BatchNorm1d -> ReLU -> Dropout on [B, C, L]
    (reference components.py:143-174)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, dropout: float = 0.1):
        super().__init__()
        self.conv = nn.Conv1d(in_channels, out_channels, kernel_size, padding=kernel_size // 2)
        self.norm = nn.BatchNorm1d(out_channels)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.dropout(F.relu(self.norm(self.conv(x))))


class LightweightResBlock(nn.Module):
    """x + conv2(leaky_relu(conv1(x), 0.1)); conv1 is dilated (reference components.py:177-200)."""

    def __init__(self, channels: int, kernel_size: int = 3, dilation: int = 1):
        super().__init__()
        self.conv1 = nn.Conv1d(channels, channels, kernel_size,
                               padding=self._get_padding(kernel_size, dilation), dilation=dilation)
        self.conv2 = nn.Conv1d(channels, channels, kernel_size,
                               padding=self._get_padding(kernel_size, 1), dilation=1)

    @staticmethod
    def _get_padding(kernel_size: int, dilation: int) -> int:
        return (kernel_size - 1) * dilation // 2

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return x + self.conv2(F.leaky_relu(self.conv1(x), 0.1))


class VariancePredictor(nn.Module):
    """Two ConvBlocks and a 1x1 projection to one channel (reference components.py:203-223)."""

    def __init__(self, hidden_dim: int, kernel_size: int = 3, dropout: float = 0.1):
        super().__init__()
        self.conv_layers = nn.ModuleList([
            ConvBlock(hidden_dim, hidden_dim, kernel_size, dropout),
            ConvBlock(hidden_dim, hidden_dim, kernel_size, dropout),
        ])
        self.projection = nn.Conv1d(hidden_dim, 1, 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for block in self.conv_layers:
            x = block(x)
        return self.projection(x)


def create_padding_mask(lengths: torch.Tensor, max_length: int) -> torch.Tensor:
    """mask[b, s] = s < lengths[b] (reference components.py:226-241)."""
    steps = torch.arange(max_length, device=lengths.device)
    return steps.unsqueeze(0).expand(lengths.size(0), max_length) < lengths.unsqueeze(1)


def count_parameters(model: nn.Module) -> Tuple[int, int]:
    sizes = [(p.numel(), p.requires_grad) for p in model.parameters()]
    return sum(n for n, _ in sizes), sum(n for n, g in sizes if g)


def initialize_weights(module: nn.Module) -> None:
    """xavier-uniform Linear, kaiming-normal Conv1d, zero biases, unit LayerNorm
    (reference components.py:274-286). ConvTranspose1d / Embedding / BatchNorm keep torch defaults."""
    if isinstance(module, nn.Linear):
        nn.init.xavier_uniform_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.Conv1d):
        nn.init.kaiming_normal_(module.weight)
        if module.bias is not None:
            nn.init.constant_(module.bias, 0)
    elif isinstance(module, nn.LayerNorm):
        nn.init.constant_(module.weight, 1)
        nn.init.constant_(module.bias, 0)
