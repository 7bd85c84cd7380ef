"""CPU oracle for the m2-tts synthesis hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A functional restatement (torch CPU fp32 ops for the floating-point stages, numpy integer
arithmetic for the length regulator) of the reference's eval-mode forward, driven by a plain
``state_dict``.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module; the product path (``m2-tts_b200/``) never
does and fails loudly without its CUDA library.

Parity status: PINNED.  The reference ships no golden vectors, so the pin is
the live reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference/src`` in the build container, runs it on seeded weights/inputs and commits the
outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those files (bit-exact for the length regulator, <= 2e-6 max-abs for fp32 stages)
and, when ``/root/reference`` is present, against the reference module directly.

Every function cites the reference lines it restates (paths relative to the reference repo).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

SAMPLES_PER_FRAME = 64  # prod([4, 4, 2, 2]), src/models/tts_model.py:244
SAMPLE_RATE = 22050     # configs/stage2_quality.yaml:69
UPSAMPLE_RATES = (4, 4, 2, 2)

STAGE_KWARGS = {
    # configs/stage1_poc.yaml:6-27, configs/stage2_quality.yaml:6-28 as consumed by
    # scripts/synthesize.py:37-46
    "stage1": dict(vocab_size=256, hidden_dim=64, mel_channels=64, text_encoder_layers=2,
                   decoder_layers=2, num_heads=2, dropout=0.1, vocoder_channels=128),
    "stage2": dict(vocab_size=256, hidden_dim=96, mel_channels=80, text_encoder_layers=3,
                   decoder_layers=3, num_heads=2, dropout=0.1, vocoder_channels=256),
    # scripts/test_pipeline.py:72-80
    "tiny": dict(vocab_size=256, hidden_dim=32, mel_channels=32, text_encoder_layers=1,
                 decoder_layers=1, num_heads=2, dropout=0.1, vocoder_channels=64),
}


def _count_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.layers.{n}.norm1.weight" in sd:
        n += 1
    return n


# --------------------------------------------------------------------------- building blocks
def padding_mask(lengths: torch.Tensor, max_length: int) -> torch.Tensor:
    """src/models/components.py:226-241 — mask[b, s] = s < lengths[b]."""
    return torch.arange(max_length)[None, :] < lengths[:, None]


def attention(sd: SD, p: str, x: torch.Tensor, num_heads: int,
              mask: Optional[torch.Tensor]) -> torch.Tensor:
    """src/models/components.py:59-90 — fused-qkv multi-head attention, scores materialised,
    key-padding mask filled with the finite value -1e9, softmax over keys, output projection."""
    B, L, H = x.shape
    hd = H // num_heads
    qkv = F.linear(x, sd[f"{p}.qkv.weight"]).reshape(B, L, 3, num_heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scores = torch.matmul(q, k.transpose(-2, -1)) * (1.0 / math.sqrt(hd))
    if mask is not None:
        m = mask[:, None, None, :].expand(B, num_heads, L, L)
        scores = scores.masked_fill(m == 0, -1e9)
    attn = F.softmax(scores, dim=-1)
    out = torch.matmul(attn, v).transpose(1, 2).reshape(B, L, H)
    return F.linear(out, sd[f"{p}.out_proj.weight"], sd[f"{p}.out_proj.bias"])


def transformer_layer(sd: SD, p: str, x: torch.Tensor, num_heads: int,
                      mask: Optional[torch.Tensor]) -> torch.Tensor:
    """src/models/components.py:131-140 — pre-LN attention and pre-LN FFN, both residual;
    FFN is linear2(relu(linear1(.))) (:103). Dropout is the identity in eval mode."""
    H = x.shape[-1]
    h = F.layer_norm(x, (H,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"], 1e-5)
    x = x + attention(sd, f"{p}.self_attn", h, num_heads, mask)
    h = F.layer_norm(x, (H,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"], 1e-5)
    h = F.linear(F.relu(F.linear(h, sd[f"{p}.ffn.linear1.weight"], sd[f"{p}.ffn.linear1.bias"])),
                 sd[f"{p}.ffn.linear2.weight"], sd[f"{p}.ffn.linear2.bias"])
    return x + h


# --------------------------------------------------------------------------- stages
def text_encoder(sd: SD, ids: torch.Tensor, lengths: Optional[torch.Tensor],
                 num_heads: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """src/models/tts_model.py:57-89 — embedding * sqrt(H) + stored `pe` buffer, N layers with a
    key-padding mask, final LayerNorm."""
    B, S = ids.shape
    emb = sd["text_encoder.embedding.weight"]
    H = emb.shape[1]
    mask = padding_mask(lengths, S) if lengths is not None else None
    x = F.embedding(ids, emb) * (H ** 0.5)
    x = x + sd["text_encoder.pos_encoding.pe"][:, :S]
    for i in range(_count_layers(sd, "text_encoder")):
        x = transformer_layer(sd, f"text_encoder.layers.{i}", x, num_heads, mask)
    x = F.layer_norm(x, (H,), sd["text_encoder.norm.weight"], sd["text_encoder.norm.bias"], 1e-5)
    return x, mask


def duration_predictor(sd: SD, enc: torch.Tensor) -> torch.Tensor:
    """src/models/tts_model.py:99-117 + components.py:154-174,214-223 — two
    [Conv1d(k=3,pad=1) -> BatchNorm1d(eval) -> ReLU] blocks, 1x1 conv to one channel, softplus."""
    p = "duration_predictor.predictor"
    x = enc.transpose(1, 2)
    for i in range(2):
        q = f"{p}.conv_layers.{i}"
        x = F.conv1d(x, sd[f"{q}.conv.weight"], sd[f"{q}.conv.bias"], padding=1)
        x = F.batch_norm(x, sd[f"{q}.norm.running_mean"], sd[f"{q}.norm.running_var"],
                         sd[f"{q}.norm.weight"], sd[f"{q}.norm.bias"], False, 0.0, 1e-5)
        x = F.relu(x)
    x = F.conv1d(x, sd[f"{p}.projection.weight"], sd[f"{p}.projection.bias"])
    return F.softplus(x.squeeze(1))


def length_regulator_indices(durations: np.ndarray, max_length: Optional[int] = None):
    """Integer restatement of src/models/tts_model.py:146-178.

    Returns (index [B,T] int32 with -1 for zero rows, frames [B] int32 = number of expanded rows
    per utterance BEFORE padding/truncation, T).  `int(float)` truncates toward zero (:150), only
    counts > 0 expand (:151), an all-zero utterance is one zero row (:158-160), T = max_length or
    the longest sequence (:165-166), longer sequences are truncated (:173-174)."""
    d = np.asarray(durations, dtype=np.float32)
    if np.isnan(d).any():
        raise ValueError("cannot convert float NaN to integer")
    if np.isinf(d).any():
        raise OverflowError("cannot convert float infinity to integer")
    n = np.trunc(d.astype(np.float64)).astype(np.int64)
    n[n < 0] = 0
    frames = n.sum(axis=1)
    seq_len = np.maximum(frames, 1)
    T = int(max_length) if max_length is not None else int(seq_len.max())
    B, S = n.shape
    index = np.full((B, T), -1, dtype=np.int32)
    for b in range(B):
        src = np.repeat(np.arange(S, dtype=np.int32), n[b])[:T]
        index[b, : src.shape[0]] = src
    return index, frames.astype(np.int32), T


def length_regulator(enc: torch.Tensor, durations: torch.Tensor,
                     max_length: Optional[int] = None) -> torch.Tensor:
    """src/models/tts_model.py:126-178 via the integer index map above (pure copies, bit-exact)."""
    index, _, T = length_regulator_indices(durations.detach().cpu().numpy(), max_length)
    idx = torch.from_numpy(index.astype(np.int64))
    B, S, H = enc.shape
    gathered = torch.gather(enc, 1, idx.clamp(min=0)[:, :, None].expand(B, T, H))
    return torch.where((idx >= 0)[:, :, None], gathered, torch.zeros((), dtype=enc.dtype))


def length_regulator_loop(enc: torch.Tensor, durations: torch.Tensor,
                          max_length: Optional[int] = None) -> torch.Tensor:
    """The same algorithm as literal per-phoneme loops (small cases only; cross-checks the
    index formulation). src/models/tts_model.py:143-178."""
    B, S, H = enc.shape
    seqs = []
    for b in range(B):
        rows = []
        for s in range(S):
            n = int(durations[b, s].item())
            if n > 0:
                rows.extend([enc[b, s]] * n)
        seqs.append(torch.stack(rows) if rows else torch.zeros(1, H))
    T = max_length if max_length is not None else max(q.shape[0] for q in seqs)
    out = torch.zeros(B, T, H)
    for b, q in enumerate(seqs):
        m = min(T, q.shape[0])
        out[b, :m] = q[:m]
    return out


def mel_decoder(sd: SD, x: torch.Tensor, num_heads: int) -> torch.Tensor:
    """src/models/tts_model.py:211-228 — N unmasked layers, LayerNorm, Linear(H -> mel)."""
    H = x.shape[-1]
    for i in range(_count_layers(sd, "decoder")):
        x = transformer_layer(sd, f"decoder.layers.{i}", x, num_heads, None)
    x = F.layer_norm(x, (H,), sd["decoder.norm.weight"], sd["decoder.norm.bias"], 1e-5)
    return F.linear(x, sd["decoder.mel_projection.weight"], sd["decoder.mel_projection.bias"])


def vocoder(sd: SD, mel: torch.Tensor, res_dilation: int = 1) -> torch.Tensor:
    """src/models/tts_model.py:279-297 — input conv, 4 x [ConvTranspose1d(k=2r, s=r, p=r//2) ->
    leaky_relu(0.1) -> x + conv2(leaky_relu(conv1(x), 0.1))] (components.py:196-200), output
    conv, tanh. mel is [B, M, T]; returns [B, 1, 64 T]."""
    x = F.conv1d(mel, sd["vocoder.input_conv.weight"], sd["vocoder.input_conv.bias"], padding=1)
    for j, r in enumerate(UPSAMPLE_RATES):
        x = F.conv_transpose1d(x, sd[f"vocoder.upsamples.{j}.weight"], sd[f"vocoder.upsamples.{j}.bias"],
                               stride=r, padding=r // 2)
        x = F.leaky_relu(x, 0.1)
        p = f"vocoder.resblocks.{j}"
        h = F.conv1d(x, sd[f"{p}.conv1.weight"], sd[f"{p}.conv1.bias"], padding=res_dilation,
                     dilation=res_dilation)
        h = F.conv1d(F.leaky_relu(h, 0.1), sd[f"{p}.conv2.weight"], sd[f"{p}.conv2.bias"], padding=1)
        x = h + x
    x = F.conv1d(x, sd["vocoder.output_conv.weight"], sd["vocoder.output_conv.bias"], padding=1)
    return torch.tanh(x)


# --------------------------------------------------------------------------- model level
@torch.no_grad()
def forward(sd: SD, ids: torch.Tensor, lengths: Optional[torch.Tensor] = None,
            target_durations: Optional[torch.Tensor] = None, max_target_length: Optional[int] = None,
            num_heads: int = 2) -> Dict[str, Optional[torch.Tensor]]:
    """src/models/tts_model.py:350-400 in eval mode (the vocoder runs, :388-391)."""
    enc, mask = text_encoder(sd, ids, lengths, num_heads)
    dur = duration_predictor(sd, enc)
    use = target_durations if target_durations is not None else dur
    reg = length_regulator(enc, use, max_target_length)
    mel = mel_decoder(sd, reg, num_heads)
    audio = vocoder(sd, mel.transpose(1, 2))
    return {"encoder_output": enc, "duration_pred": dur, "regulated_output": reg,
            "mel_output": mel, "audio_output": audio, "padding_mask": mask}


@torch.no_grad()
def inference(sd: SD, ids: torch.Tensor, lengths: Optional[torch.Tensor] = None,
              duration_scale: float = 1.0, num_heads: int = 2):
    """src/models/tts_model.py:402-438 — forward, optional re-regulation with scaled predicted
    durations (:426-432), vocoder on the final mel (:435-436)."""
    out = forward(sd, ids, lengths, num_heads=num_heads)
    mel = out["mel_output"]
    if duration_scale != 1.0:
        reg = length_regulator(out["encoder_output"], out["duration_pred"] * duration_scale)
        mel = mel_decoder(sd, reg, num_heads)
    return mel, vocoder(sd, mel.transpose(1, 2))


def audio_seconds(frames: int) -> float:
    """The metric's numerator: 64 samples per mel frame at 22 050 Hz."""
    return frames * SAMPLES_PER_FRAME / SAMPLE_RATE


# --------------------------------------------------------------------------- FLOP model (roofline numerators)
def vocoder_flops_per_frame(M: int, C: int) -> int:
    """2*[3MC + sum_j(c_j*(c_j/2)*2r_j*R_j + 2*3*(c_j/2)^2*R_j*r_j) + 3*(C/16)*64]."""
    tot = 3 * M * C
    R, c = 1, C
    for r in UPSAMPLE_RATES:
        tot += c * (c // 2) * 2 * r * R + 2 * 3 * (c // 2) ** 2 * R * r
        R *= r
        c //= 2
    tot += 3 * c * R
    return 2 * tot


def decoder_flops_per_frame(H: int, M: int, layers: int, T: int) -> int:
    """L*(16 H^2 + 4 T H) + 2 H M."""
    return layers * (16 * H * H + 4 * T * H) + 2 * H * M


def attention_flops(B: int, L: int, H: int) -> int:
    """QK^T and PV: 4*L*H per query row (all heads together)."""
    return B * L * 4 * L * H
