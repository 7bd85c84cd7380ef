"""m2-tts model assembly for B200 — drop-in for the reference's ``src/models/tts_model.py``.

Same public surface: ``TextEncoder``, ``DurationPredictor``, ``LengthRegulator``, ``MelDecoder``,
``SimpleVocoder``, ``M2TTSModel`` with the reference's constructor arguments, forward/inference
signatures, return values and ``state_dict`` layout (reference tts_model.py:19-459), importable as
``models.tts_model`` with ``m2-tts_b200/src`` on ``sys.path`` exactly like the reference's ``src``.

Execution model
  * ``module.training == False`` (the synthesis path): every stage is one or a few C-ABI calls
    into ``libm2tts_b200.so`` (hand-written sm_100a CUDA, ``include/m2tts_b200.h``) on the
    caller's current CUDA stream. Outputs carry no autograd graph. There is NO CPU / MPS /
    PyTorch fallback: CPU tensors or a missing library raise ``NativeLibraryError``.
    Each stage module validates its result against the device status word (``_native.guarded``):
    operands outside the fp16 range of the default 16-bit split make the stage run again with the
    TF32 split, an out-of-range phoneme id raises ``IndexError`` like ``nn.Embedding``. Weight
    images are packed once per module and cached until a parameter changes.
  * ``module.training == True``: the reference's formulation in plain differentiable torch ops
    on the same parameters (dropout, activation checkpointing, no vocoder in ``forward``).
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nat
from .components import (LightweightResBlock, PositionalEncoding, TransformerEncoderLayer,
                         VariancePredictor, count_parameters, create_padding_mask,
                         initialize_weights)

UPSAMPLE_RATES = (4, 4, 2, 2)   # 64 samples per mel frame (reference tts_model.py:244)
LN_EPS = 1e-5


# =============================================================================================
# native helpers
# =============================================================================================
def _layer_struct(layer: TransformerEncoderLayer) -> nat.LayerWeights:
    a, f = layer.self_attn, layer.ffn
    w = nat.weight
    return nat.LayerWeights(
        w(layer.norm1.weight, "norm1.weight"), w(layer.norm1.bias, "norm1.bias"),
        w(a.qkv.weight, "self_attn.qkv.weight"),
        w(a.out_proj.weight, "self_attn.out_proj.weight"), w(a.out_proj.bias, "self_attn.out_proj.bias"),
        w(layer.norm2.weight, "norm2.weight"), w(layer.norm2.bias, "norm2.bias"),
        w(f.linear1.weight, "ffn.linear1.weight"), w(f.linear1.bias, "ffn.linear1.bias"),
        w(f.linear2.weight, "ffn.linear2.weight"), w(f.linear2.bias, "ffn.linear2.bias"))


def _layer_packed(layer: TransformerEncoderLayer, st: nat.LayerWeights, H: int, F_dim: int, prec: int,
                  dev: torch.device) -> torch.Tensor:
    """Weight images of one layer (m2tts_transformer_pack), cached on the layer until a weight changes."""
    a, f = layer.self_attn, layer.ffn
    key = nat.params_key((a.qkv.weight, a.out_proj.weight, f.linear1.weight, f.linear2.weight)) + (H, F_dim, prec)
    lib = nat.lib()

    def pack(buf: torch.Tensor) -> None:
        nat.check(lib.m2tts_transformer_pack(C.byref(st), H, F_dim, prec, buf.data_ptr(), buf.numel(),
                                             nat.status_ptr(dev), nat.stream_handle(dev)), "transformer_pack")

    return nat.packed_weights(layer, f"layer{prec}", key, lib.m2tts_transformer_pack_bytes(H, F_dim, prec), pack, dev)


def _native_layer_stack(layers: Sequence[TransformerEncoderLayer], x: torch.Tensor,
                        lengths: Optional[torch.Tensor], out: Optional[torch.Tensor] = None,
                        prec: int = nat.PREC_DEFAULT) -> torch.Tensor:
    """Run pre-LN transformer layers through ``m2tts_transformer_layer`` (status word: the caller's ``guarded``).
    ``x`` is not modified unless ``out is x``; the result is written to ``out`` (allocated when None)."""
    nat.require_cuda(x, "x")
    x = x.contiguous()
    B, L, H = x.shape
    dev = x.device
    if out is None:
        out = torch.empty_like(x)
    if lengths is not None:
        lengths = nat.require_cuda(lengths.to(device=dev, dtype=torch.int64), "lengths", torch.int64).contiguous()
    lib = nat.lib()
    src = x
    for layer in layers:
        F_dim = layer.ffn.linear1.out_features
        nbytes = lib.m2tts_transformer_workspace_bytes(B, L, H, F_dim)
        ws = nat.workspace(dev, nbytes)
        st = _layer_struct(layer)
        packed = _layer_packed(layer, st, H, F_dim, prec, dev)
        rc = lib.m2tts_transformer_layer(C.byref(st), packed.data_ptr(), src.data_ptr(), out.data_ptr(), nat.ptr(lengths),
                                         B, L, H, layer.self_attn.num_heads, F_dim, LN_EPS, prec, nat.status_ptr(dev),
                                         ws.data_ptr(), ws.numel(), nat.stream_handle(dev))
        nat.check(rc, "transformer_layer")
        src = out
    if not layers:
        out.copy_(x)
    return out


# =============================================================================================
# stages
# =============================================================================================
class TextEncoder(nn.Module):
    """Embedding * sqrt(H) + positional table, N masked pre-LN layers, LayerNorm
    (reference tts_model.py:19-89)."""

    def __init__(self, vocab_size: int = 256, hidden_dim: int = 64, num_layers: int = 2,
                 num_heads: int = 2, dropout: float = 0.1, max_seq_len: int = 1000):
        super().__init__()
        self.hidden_dim = hidden_dim
        self.embedding = nn.Embedding(vocab_size, hidden_dim)
        self.pos_encoding = PositionalEncoding(hidden_dim, max_seq_len)
        self.layers = nn.ModuleList(
            TransformerEncoderLayer(hidden_dim=hidden_dim, num_heads=num_heads,
                                    ffn_dim=hidden_dim * 2, dropout=dropout)
            for _ in range(num_layers))
        self.norm = nn.LayerNorm(hidden_dim)
        self.dropout = nn.Dropout(dropout)
        self.apply(initialize_weights)

    def forward(self, phoneme_ids: torch.Tensor,
                lengths: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        if self.training:
            return self._forward_train(phoneme_ids, lengths)
        return self._forward_native(phoneme_ids, lengths)

    def _forward_train(self, phoneme_ids, lengths):
        mask = create_padding_mask(lengths, phoneme_ids.size(1)) if lengths is not None else None
        x = self.embedding(phoneme_ids) * (self.hidden_dim ** 0.5)
        x = self.dropout(self.pos_encoding(x))
        for layer in self.layers:
            x = layer(x, mask)
        return self.norm(x), mask

    @torch.no_grad()
    def _forward_native(self, phoneme_ids, lengths):
        ids = nat.require_cuda(phoneme_ids, "phoneme_ids", torch.int64).contiguous()
        dev = ids.device
        B, S = ids.shape
        H = self.hidden_dim
        pe = self.pos_encoding.pe
        if S > pe.size(1):
            raise RuntimeError(f"sequence length {S} exceeds the positional table ({pe.size(1)})")
        lens = None
        mask = None
        if lengths is not None:
            lens = nat.require_cuda(lengths.to(device=dev, dtype=torch.int64), "lengths", torch.int64).contiguous()
            mask = torch.empty((B, S), dtype=torch.bool, device=dev)
        lib = nat.lib()
        y = torch.empty((B, S, H), dtype=torch.float32, device=dev)

        def run(prec: int):
            x = torch.empty((B, S, H), dtype=torch.float32, device=dev)
            rc = lib.m2tts_embed_posenc(ids.data_ptr(), nat.weight(self.embedding.weight, "embedding.weight"),
                                        nat.weight(pe, "pos_encoding.pe"), nat.ptr(lens), x.data_ptr(),
                                        nat.ptr(mask), B, S, H, self.embedding.num_embeddings, nat.status_ptr(dev),
                                        nat.stream_handle(dev))
            nat.check(rc, "embed_posenc")
            _native_layer_stack(list(self.layers), x, lens, out=x, prec=prec)
            rc = lib.m2tts_layernorm(x.data_ptr(), nat.weight(self.norm.weight, "norm.weight"),
                                     nat.weight(self.norm.bias, "norm.bias"), y.data_ptr(), B * S, H, LN_EPS,
                                     nat.stream_handle(dev))
            nat.check(rc, "layernorm")

        with nat.on_device(dev):
            nat.guarded(dev, run, "text_encoder", owner=self)
        return y, mask


class DurationPredictor(nn.Module):
    """softplus(VariancePredictor(enc^T)) -> [B, S] (reference tts_model.py:92-117)."""

    def __init__(self, hidden_dim: int = 64, kernel_size: int = 3, dropout: float = 0.1):
        super().__init__()
        self.predictor = VariancePredictor(hidden_dim, kernel_size, dropout)

    def forward(self, encoder_output: torch.Tensor) -> torch.Tensor:
        if self.training:
            return F.softplus(self.predictor(encoder_output.transpose(1, 2)).squeeze(1))
        return self._forward_native(encoder_output)

    @torch.no_grad()
    def _forward_native(self, enc: torch.Tensor) -> torch.Tensor:
        enc = nat.require_cuda(enc, "encoder_output").contiguous()
        B, S, H = enc.shape
        blocks = self.predictor.conv_layers
        for blk in blocks:
            if blk.conv.kernel_size != (3,):
                raise ValueError("m2tts_b200 duration_predictor: only kernel_size=3 is supported")
        w = nat.weight
        st = nat.DurPredWeights()
        for i, blk in enumerate(blocks):
            st.conv_w[i] = w(blk.conv.weight, "conv.weight")
            st.conv_b[i] = w(blk.conv.bias, "conv.bias")
            st.bn_w[i] = w(blk.norm.weight, "norm.weight")
            st.bn_b[i] = w(blk.norm.bias, "norm.bias")
            st.bn_mean[i] = w(blk.norm.running_mean, "norm.running_mean")
            st.bn_var[i] = w(blk.norm.running_var, "norm.running_var")
        st.proj_w = w(self.predictor.projection.weight, "projection.weight")
        st.proj_b = w(self.predictor.projection.bias, "projection.bias")
        st.bn_eps = blocks[0].norm.eps
        dur = torch.empty((B, S), dtype=torch.float32, device=enc.device)
        with nat.on_device(enc.device):
            rc = nat.lib().m2tts_duration_predictor(C.byref(st), enc.data_ptr(), dur.data_ptr(), B, S, H,
                                                    nat.stream_handle(enc.device))
        nat.check(rc, "duration_predictor")
        return dur


class LengthRegulator(nn.Module):
    """Expand each phoneme's encoder row ``int(duration)`` times, pad/truncate to a common length
    (reference tts_model.py:120-178). Eval: integer scan + gather kernels, one 8-byte host read
    (frame maximum + status). Train: the same index map in differentiable torch ops."""

    def __init__(self):
        super().__init__()
        self.last_frames: Optional[torch.Tensor] = None   # int32 [B], frames per utterance
        self.last_index: Optional[torch.Tensor] = None    # int32 [B,T], source phoneme or -1

    def forward(self, encoder_output: torch.Tensor, durations: torch.Tensor,
                max_length: Optional[int] = None) -> torch.Tensor:
        if self.training or encoder_output.requires_grad:
            return self._forward_train(encoder_output, durations, max_length)
        return self._forward_native(encoder_output, durations, max_length)

    @staticmethod
    def _forward_train(enc, durations, max_length):
        if torch.isnan(durations).any():
            raise ValueError("cannot convert float NaN to integer")
        if torch.isinf(durations).any():
            raise OverflowError("cannot convert float infinity to integer")
        B, S, H = enc.shape
        n = durations.detach().trunc().clamp_(min=0).to(torch.int64)
        cum = n.cumsum(dim=1)
        frames = cum[:, -1]
        T = int(max_length) if max_length is not None else int(frames.clamp(min=1).max().item())
        j = torch.arange(T, device=enc.device).expand(B, T)
        src = torch.searchsorted(cum, j.contiguous(), right=True).clamp_(max=S - 1)
        rows = torch.gather(enc, 1, src.unsqueeze(-1).expand(B, T, H))
        return rows * (j < frames.unsqueeze(1)).unsqueeze(-1).to(enc.dtype)

    @torch.no_grad()
    def _forward_native(self, enc, durations, max_length):
        enc = nat.require_cuda(enc, "encoder_output").contiguous()
        dev = enc.device
        B, S, H = enc.shape
        dur = nat.require_cuda(durations.to(device=dev, dtype=torch.float32), "durations").contiguous()
        if dur.shape != (B, S):
            raise ValueError(f"durations must be [{B}, {S}], got {tuple(dur.shape)}")
        lib = nat.lib()
        cum = torch.empty((B, S), dtype=torch.int32, device=dev)
        frames = torch.empty((B,), dtype=torch.int32, device=dev)
        meta = torch.empty((2,), dtype=torch.int32, device=dev)    # [t_max, status]
        with nat.on_device(dev):
            return self._regulate(lib, enc, dur, cum, frames, meta, max_length, B, S, H, dev)

    def _regulate(self, lib, enc, dur, cum, frames, meta, max_length, B, S, H, dev):
        rc = lib.m2tts_length_regulate_count(dur.data_ptr(), B, S, cum.data_ptr(), frames.data_ptr(),
                                             meta.data_ptr(), meta.data_ptr() + 4, nat.stream_handle(dev))
        nat.check(rc, "length_regulate_count")
        if max_length is not None and nat.status_deferred():
            # no host read inside a deferred region (throughput pipelines, CUDA-graph capture): the regulator's own status
            # bits join the device status word and surface at check_status()
            nat.status_word(dev).bitwise_or_(meta[1:2] << 2)
            t_max, status = int(max_length), 0
        else:
            t_max, status = meta.tolist()           # the path's only host read
        if status & 1:
            raise ValueError("cannot convert float NaN to integer")
        if status & 2:
            raise OverflowError("cannot convert float infinity to integer")
        if status & 4:
            raise OverflowError("length regulator: frame count exceeds int32")
        T = int(max_length) if max_length is not None else int(t_max)
        if T <= 0:
            raise ValueError(f"max_length must be positive, got {T}")
        out = torch.empty((B, T, H), dtype=torch.float32, device=dev)
        index = torch.empty((B, T), dtype=torch.int32, device=dev)
        rc = lib.m2tts_length_regulate_gather(enc.data_ptr(), cum.data_ptr(), frames.data_ptr(),
                                              out.data_ptr(), index.data_ptr(), B, S, H, T,
                                              nat.stream_handle(dev))
        nat.check(rc, "length_regulate_gather")
        self.last_frames, self.last_index = frames, index
        return out


class MelDecoder(nn.Module):
    """N unmasked pre-LN layers, LayerNorm, Linear(H -> mel) (reference tts_model.py:181-228)."""

    def __init__(self, hidden_dim: int = 64, mel_channels: int = 64, num_layers: int = 2,
                 num_heads: int = 2, dropout: float = 0.1):
        super().__init__()
        self.layers = nn.ModuleList(
            TransformerEncoderLayer(hidden_dim=hidden_dim, num_heads=num_heads,
                                    ffn_dim=hidden_dim * 2, dropout=dropout)
            for _ in range(num_layers))
        self.norm = nn.LayerNorm(hidden_dim)
        self.mel_projection = nn.Linear(hidden_dim, mel_channels)
        self.apply(initialize_weights)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training:
            for layer in self.layers:
                x = layer(x)
            return self.mel_projection(self.norm(x))
        return self._forward_native(x)

    @torch.no_grad()
    def _forward_native(self, x: torch.Tensor) -> torch.Tensor:
        x = nat.require_cuda(x, "x").contiguous()
        B, T, H = x.shape
        dev = x.device
        M = self.mel_projection.out_features
        lib = nat.lib()
        mel = torch.empty((B, T, M), dtype=torch.float32, device=dev)
        proj = self.mel_projection

        def run(prec: int):
            h = _native_layer_stack(list(self.layers), x, None, prec=prec)
            key = nat.params_key((proj.weight,)) + (H, M, prec)

            def pack(buf: torch.Tensor) -> None:
                nat.check(lib.m2tts_ln_proj_pack(nat.weight(proj.weight, "mel_projection.weight"), H, M, prec, buf.data_ptr(),
                                                 buf.numel(), nat.status_ptr(dev), nat.stream_handle(dev)), "ln_proj_pack")

            packed = nat.packed_weights(self, f"ln_proj{prec}", key, lib.m2tts_ln_proj_pack_bytes(H, M, prec), pack, dev)
            ws = nat.workspace(dev, lib.m2tts_ln_proj_rows_workspace_bytes(B * T, H, M), tag="ln_proj")
            rc = lib.m2tts_layernorm_proj(h.data_ptr(), nat.weight(self.norm.weight, "norm.weight"),
                                          nat.weight(self.norm.bias, "norm.bias"),
                                          nat.weight(proj.weight, "mel_projection.weight"),
                                          nat.weight(proj.bias, "mel_projection.bias"), packed.data_ptr(),
                                          mel.data_ptr(), B * T, H, M, LN_EPS, prec, nat.status_ptr(dev),
                                          ws.data_ptr(), ws.numel(), nat.stream_handle(dev))
            nat.check(rc, "layernorm_proj")

        with nat.on_device(dev):
            nat.guarded(dev, run, "mel_decoder", owner=self)
        return mel


class SimpleVocoder(nn.Module):
    """input conv -> 4 x [ConvTranspose1d(k=2r, s=r, p=r/2) + LeakyReLU(0.1) + ResBlock] ->
    output conv -> tanh; mel [B, M, T] -> waveform [B, 1, 64 T] (reference tts_model.py:231-297)."""

    def __init__(self, mel_channels: int = 64, hidden_channels: int = 128, kernel_size: int = 3,
                 n_layers: int = 4):
        super().__init__()
        self.input_conv = nn.Conv1d(mel_channels, hidden_channels, kernel_size, padding=kernel_size // 2)
        self.upsamples = nn.ModuleList()
        self.resblocks = nn.ModuleList()
        ch = hidden_channels
        for rate in UPSAMPLE_RATES:
            self.upsamples.append(nn.ConvTranspose1d(ch, ch // 2, kernel_size=rate * 2, stride=rate,
                                                     padding=rate // 2))
            ch = ch // 2
            self.resblocks.append(LightweightResBlock(ch, kernel_size))
        self.output_conv = nn.Conv1d(ch, 1, kernel_size, padding=kernel_size // 2)
        self.apply(initialize_weights)

    def forward(self, mel: torch.Tensor) -> torch.Tensor:
        if self.training:
            x = self.input_conv(mel)
            for up, res in zip(self.upsamples, self.resblocks):
                x = res(F.leaky_relu(up(x), 0.1))
            return torch.tanh(self.output_conv(x))
        return self._forward_native(mel)

    @torch.no_grad()
    def _forward_native(self, mel: torch.Tensor) -> torch.Tensor:
        nat.require_cuda(mel, "mel")
        if mel.dim() != 3:
            raise ValueError(f"mel must be [B, mel_channels, T], got {tuple(mel.shape)}")
        B, M, T = mel.shape
        if M != self.input_conv.in_channels:
            raise ValueError(f"mel has {M} channels, vocoder expects {self.input_conv.in_channels}")
        if self.input_conv.kernel_size != (3,):
            raise ValueError("m2tts_b200 vocoder: only kernel_size=3 is supported")
        Cch = self.input_conv.out_channels
        dev = mel.device
        w = nat.weight
        st = nat.VocoderWeights()
        st.in_w, st.in_b = w(self.input_conv.weight, "input_conv.weight"), w(self.input_conv.bias, "input_conv.bias")
        for j, (up, res) in enumerate(zip(self.upsamples, self.resblocks)):
            st.up_w[j], st.up_b[j] = w(up.weight, "upsamples.weight"), w(up.bias, "upsamples.bias")
            st.res1_w[j], st.res1_b[j] = w(res.conv1.weight, "conv1.weight"), w(res.conv1.bias, "conv1.bias")
            st.res2_w[j], st.res2_b[j] = w(res.conv2.weight, "conv2.weight"), w(res.conv2.bias, "conv2.bias")
            st.res_dilation[j] = int(res.conv1.dilation[0])
        st.out_w, st.out_b = w(self.output_conv.weight, "output_conv.weight"), w(self.output_conv.bias, "output_conv.bias")
        lib = nat.lib()
        total = 1
        for r in UPSAMPLE_RATES:
            total *= r
        audio = torch.empty((B, 1, total * T), dtype=torch.float32, device=dev)
        sb, sm, stt = mel.stride()
        params = [self.input_conv.weight, self.output_conv.weight]
        for up, res in zip(self.upsamples, self.resblocks):
            params += [up.weight, res.conv1.weight, res.conv2.weight]
        dils = tuple(int(res.conv1.dilation[0]) for res in self.resblocks)

        def run(prec: int):
            key = nat.params_key(params) + (M, Cch, prec) + dils

            def pack(buf: torch.Tensor) -> None:
                nat.check(lib.m2tts_vocoder_pack(C.byref(st), M, Cch, prec, buf.data_ptr(), buf.numel(),
                                                 nat.status_ptr(dev), nat.stream_handle(dev)), "vocoder_pack")

            packed = nat.packed_weights(self, f"vocoder{prec}", key, lib.m2tts_vocoder_pack_bytes(M, Cch, prec), pack, dev)
            ws = nat.workspace(dev, lib.m2tts_vocoder_workspace_bytes(B, T, M, Cch), tag="vocoder")
            rc = lib.m2tts_vocoder_forward(C.byref(st), packed.data_ptr(), mel.data_ptr(), sb, sm, stt, audio.data_ptr(),
                                           B, T, M, Cch, prec, nat.status_ptr(dev), ws.data_ptr(), ws.numel(),
                                           nat.stream_handle(dev))
            nat.check(rc, "vocoder_forward")

        with nat.on_device(dev):
            nat.guarded(dev, run, "vocoder", owner=self)
        return audio


# =============================================================================================
# model
# =============================================================================================
class M2TTSModel(nn.Module):
    """text_encoder -> duration_predictor -> length_regulator -> decoder -> vocoder
    (reference tts_model.py:300-459)."""

    def __init__(self, vocab_size: int = 256, hidden_dim: int = 64, mel_channels: int = 64,
                 text_encoder_layers: int = 2, decoder_layers: int = 2, num_heads: int = 2,
                 dropout: float = 0.1, vocoder_channels: int = 128):
        super().__init__()
        self.text_encoder = TextEncoder(vocab_size=vocab_size, hidden_dim=hidden_dim,
                                        num_layers=text_encoder_layers, num_heads=num_heads, dropout=dropout)
        self.duration_predictor = DurationPredictor(hidden_dim=hidden_dim, dropout=dropout)
        self.length_regulator = LengthRegulator()
        self.decoder = MelDecoder(hidden_dim=hidden_dim, mel_channels=mel_channels,
                                  num_layers=decoder_layers, num_heads=num_heads, dropout=dropout)
        self.vocoder = SimpleVocoder(mel_channels=mel_channels, hidden_channels=vocoder_channels)

    def _acoustic(self, phoneme_ids, phoneme_lengths, target_durations, max_target_length):
        enc, mask = self.text_encoder(phoneme_ids, phoneme_lengths)
        dur = self.duration_predictor(enc)
        used = target_durations if target_durations is not None else dur
        reg = self.length_regulator(enc, used, max_target_length)
        mel = self.decoder(reg)
        return {"encoder_output": enc, "duration_pred": dur, "regulated_output": reg,
                "mel_output": mel, "audio_output": None, "padding_mask": mask}

    def forward(self, phoneme_ids: torch.Tensor, phoneme_lengths: Optional[torch.Tensor] = None,
                target_durations: Optional[torch.Tensor] = None,
                max_target_length: Optional[int] = None) -> Dict[str, Optional[torch.Tensor]]:
        out = self._acoustic(phoneme_ids, phoneme_lengths, target_durations, max_target_length)
        if not self.training:   # audio only outside training (reference tts_model.py:388-391)
            out["audio_output"] = self.vocoder(out["mel_output"].transpose(1, 2))
        return out

    def inference(self, phoneme_ids: torch.Tensor, phoneme_lengths: Optional[torch.Tensor] = None,
                  duration_scale: float = 1.0) -> Tuple[torch.Tensor, torch.Tensor]:
        """(mel [B,T,M], waveform [B,1,64T]) — same results as the reference's inference()
        (tts_model.py:402-438) without its redundant second vocoder pass."""
        self.eval()
        with torch.no_grad():
            out = self._acoustic(phoneme_ids, phoneme_lengths, None, None)
            mel = out["mel_output"]
            if duration_scale != 1.0:
                reg = self.length_regulator(out["encoder_output"], out["duration_pred"] * duration_scale)
                mel = self.decoder(reg)
            return mel, self.vocoder(mel.transpose(1, 2))

    def get_model_size(self) -> Dict[str, Any]:
        parts = {}
        for name, child in self.named_children():
            tot, trainable = count_parameters(child)
            parts[name] = {"total": tot, "trainable": trainable, "size_mb": tot * 4 / (1024 * 1024)}
        tot, trainable = count_parameters(self)
        return {"total_params": tot, "trainable_params": trainable,
                "total_size_mb": tot * 4 / (1024 * 1024), "components": parts}
