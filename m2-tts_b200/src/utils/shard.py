"""Batch sharding for multi-GPU synthesis: one process per GPU, utterances split in contiguous
blocks, no data-path collective.

Every utterance is independent in eval mode, PROVIDED all ranks use the same padded phoneme
length and the same `max_target_length`: the decoder attends over zero-padded frames without a
mask (reference tts_model.py:219-221), so a frame's mel depends on how much padding follows it.
The only exchange on the path is therefore one 4-byte all-reduce(MAX) of the frame maximum when
the caller does not fix `max_target_length`; outputs are gathered after the path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous block of `n` items owned by `rank` (first n % world ranks get one more)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors: Sequence[Optional[torch.Tensor]], rank: int, world: int) -> List[Optional[torch.Tensor]]:
    """Slice dim 0 of every tensor to this rank's block."""
    n = next(t.shape[0] for t in tensors if t is not None)
    lo, hi = shard_bounds(n, rank, world)
    return [None if t is None else t[lo:hi] for t in tensors]


def shared_max_target_length(local_frames: torch.Tensor, group=None) -> int:
    """max over ALL ranks of max(1, frames) — the one mid-path exchange (4 bytes)."""
    t = local_frames.clamp(min=1).max().to(torch.int32).reshape(1)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t.item())


def frames_from_durations(durations: torch.Tensor) -> torch.Tensor:
    """Per-utterance frame counts with the reference's int() truncation (tts_model.py:150-151)."""
    return durations.trunc().clamp(min=0).to(torch.int64).sum(dim=1)


def gather_batch(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather dim-0 shards (possibly uneven) back into the full batch order, on every rank."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    cap = max(max(sizes), 1)
    pad = local
    if local.shape[0] < cap:
        pad = torch.cat([local, local.new_zeros((cap - local.shape[0],) + tuple(local.shape[1:]))], 0)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous(), group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], 0)


def synthesize_sharded(model, phoneme_ids: torch.Tensor, phoneme_lengths: Optional[torch.Tensor],
                       target_durations: Optional[torch.Tensor] = None,
                       max_target_length: Optional[int] = None, group=None,
                       gather: bool = True) -> Dict[str, torch.Tensor]:
    """Run `model.forward` on this rank's block of the (replicated) batch description and, if
    `gather`, all-gather mel/audio after the path. Results equal the single-process forward of
    the whole batch."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = phoneme_ids.shape[0]
    ids, lens, dur = shard_batch([phoneme_ids, phoneme_lengths, target_durations], rank, world)
    if max_target_length is None:
        if dur is None:
            # every rank takes this branch (the arguments are replicated), so nobody is left waiting in a collective
            raise ValueError("sharded synthesis with predicted durations needs max_target_length "
                             "(or run the duration predictor first and pass its output as target_durations)")
        local = frames_from_durations(dur) if dur.shape[0] > 0 else torch.ones(1, dtype=torch.int64, device=dur.device)
        max_target_length = shared_max_target_length(local, group)
    if ids.shape[0] > 0:
        out = model(ids, lens, target_durations=dur, max_target_length=max_target_length)
        mel, audio = out["mel_output"], out["audio_output"]
    else:
        # fewer utterances than ranks: this rank owns an empty block. It must still enter the collectives below with
        # zero-row tensors of the agreed shapes, or the other ranks wait in all_gather forever.
        dev = next(model.parameters()).device
        mel_ch = model.decoder.mel_projection.out_features
        mel = torch.zeros((0, max_target_length, mel_ch), dtype=torch.float32, device=dev)
        audio = None if model.training else torch.zeros((0, 1, 64 * max_target_length), dtype=torch.float32, device=dev)
    res = {"mel_output": mel, "audio_output": audio, "max_target_length": max_target_length}
    if gather:
        for k in ("mel_output", "audio_output"):
            if res[k] is not None:
                res[k] = gather_batch(res[k], n, group)
    return res
