"""Host-to-host synthesis with the copies hidden behind compute.

The reference's caller (`scripts/synthesize.py:79-83`, `src/evaluation/metrics.py:336-341`) hands the model host
tensors and wants the waveform back on the host. At B200 speeds the two PCIe copies of a 64 x 10 s batch (85 MB in,
56 MB out) cost as much as two decoder layers, so `HostPipeline` splits the utterance batch into chunks and runs
three streams: H2D of chunk i+1 and D2H of chunk i-1 proceed while chunk i computes. Utterances are independent in
eval mode, so chunking dim 0 does not change any result.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch


class HostPipeline:
    def __init__(self, device: torch.device, n_chunks: int = 2, edge: float = 1.0, sizes: Optional[Sequence[int]] = None):
        """`edge` < 1 makes the first and the last chunk smaller than the inner ones (their H2D / D2H copy is the part of
        the PCIe traffic that nothing overlaps): edge = 0.5 with three chunks splits 64 utterances 16 / 32 / 16.
        `sizes` gives the chunk sizes explicitly (used when they sum to the batch; any other batch falls back to
        `n_chunks` / `edge`): the attention kernel's CTAs come in waves of 148, so a chunk of 10, 16, 21, 27 or 32 ten-second
        utterances fills its last wave and one of 11, 22 or 28 does not."""
        if n_chunks < 1:
            raise ValueError("n_chunks must be >= 1")
        if not 0.0 < edge <= 1.0:
            raise ValueError("edge must be in (0, 1]")
        if sizes is not None and (len(sizes) == 0 or min(sizes) < 1):
            raise ValueError("sizes must be positive")
        self.device = torch.device(device)
        self.n_chunks = n_chunks
        self.edge = edge
        self.sizes = None if sizes is None else [int(v) for v in sizes]
        self.tuned = None      # filled by autotune(): the candidates it measured and their times
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self._in: List[Optional[torch.Tensor]] = []
        self._free: List[Optional[torch.cuda.Event]] = []     # chunk input buffer i may be overwritten

    def chunk_bounds(self, n: int) -> List[Tuple[int, int]]:
        """The chunks of a batch of n utterances: the explicit `sizes` when they fit it, else `bounds(n, n_chunks, edge)`."""
        if self.sizes is not None and sum(self.sizes) == n:
            out, lo = [], 0
            for sz in self.sizes:
                out.append((lo, lo + sz))
                lo += sz
            return out
        return self.bounds(n, self.n_chunks, self.edge)

    @staticmethod
    def bounds(n: int, chunks: int, edge: float = 1.0) -> List[Tuple[int, int]]:
        if n <= 0:
            return []
        chunks = min(chunks, n)
        if chunks < 3 or edge >= 1.0:
            base, extra = divmod(n, chunks)
            sizes = [base + (1 if i < extra else 0) for i in range(chunks)]
        else:
            unit = n / (chunks - 2 + 2 * edge)
            e = max(1, int(round(unit * edge)))
            base, extra = divmod(n - 2 * e, chunks - 2)
            sizes = [e] + [base + (1 if i < extra else 0) for i in range(chunks - 2)] + [e]
            if min(sizes) < 1:
                return HostPipeline.bounds(n, chunks)
        out, lo = [], 0
        for sz in sizes:
            out.append((lo, lo + sz))
            lo += sz
        return out

    def run(self, fn: Callable[[torch.Tensor], torch.Tensor], x_host: torch.Tensor, out_host: torch.Tensor) -> None:
        """out_host[b] = fn(x_host[b].to(device)) for every b. Both host tensors should be pinned; returns after
        enqueueing — call `synchronize()` (or `self.d2h.synchronize()`) before reading `out_host`. `x_host` may also be a tensor
        that is ALREADY on the device (the ids path: the regulated encoder output): then only the results are copied."""
        if out_host.is_cuda:
            raise ValueError("HostPipeline.run writes to a HOST tensor")
        resident = x_host.is_cuda
        compute = torch.cuda.current_stream(self.device)
        bnds = self.chunk_bounds(x_host.shape[0])
        while len(self._in) < len(bnds):
            self._in.append(None)
            self._free.append(None)
        ready = []
        for i, (lo, hi) in enumerate(bnds):
            if resident:
                break
            shape = (hi - lo,) + tuple(x_host.shape[1:])
            buf = self._in[i]
            if buf is None or tuple(buf.shape) != shape or buf.dtype != x_host.dtype:
                buf = torch.empty(shape, dtype=x_host.dtype, device=self.device)
                self._in[i], self._free[i] = buf, None
            with torch.cuda.stream(self.h2d):
                if self._free[i] is not None:
                    self.h2d.wait_event(self._free[i])
                buf.copy_(x_host[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.h2d)
            ready.append(ev)
        for i, (lo, hi) in enumerate(bnds):
            if resident:
                y = fn(x_host[lo:hi])
            else:
                compute.wait_event(ready[i])
                y = fn(self._in[i])
            done = torch.cuda.Event()
            done.record(compute)
            if not resident:
                self._free[i] = done
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(done)
                out_host[lo:hi].copy_(y, non_blocking=True)
            y.record_stream(self.d2h)

    def synchronize(self) -> None:
        self.d2h.synchronize()

    AUTOTUNE_SLOTS = 8      # fixed length of the vector the ranks agree on: the number of collectives never depends on the candidates

    def _sm_count(self) -> int:
        return torch.cuda.get_device_properties(self.device).multi_processor_count

    def _utterance_ms(self, fn: Callable[[torch.Tensor], torch.Tensor], x_host: torch.Tensor) -> float:
        """Compute time per utterance of the whole batch with the input resident (after one warm-up call)."""
        import time
        x_dev = x_host.to(self.device)
        fn(x_dev)
        torch.cuda.synchronize(self.device)
        t0 = time.perf_counter()
        fn(x_dev)
        torch.cuda.synchronize(self.device)
        return (time.perf_counter() - t0) * 1e3 / max(int(x_host.shape[0]), 1)

    def _agree_max(self, values: List[float], collective: bool = True) -> List[float]:
        """Element-wise maximum over the ranks (one all-reduce of a fixed-length vector); the identity without torch.distributed
        or when the caller tunes every rank on its own."""
        import torch.distributed as dist
        if not (collective and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return list(values)
        t = torch.tensor(values, dtype=torch.float64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(v) for v in t.tolist()]

    def autotune(self, fn: Callable[[torch.Tensor], torch.Tensor], x_host: torch.Tensor, out_host: torch.Tensor, frames: int, heads: int,
                 reps: int = 3, collective: bool = False) -> List[int]:
        """Measure a handful of chunk layouts (`candidate_layouts`) on the real step and keep the fastest in `self.sizes`. Which
        layout wins depends on what the copy streams get: a GPU with the PCIe link to itself likes few large chunks with small
        edges (5 / 27 / 27 / 5 for the C3 batch), eight ranks of one host copying at once (20 GB/s each instead of 55) like more,
        smaller ones (5 / 10 / 14 / 15 / 15 / 5). By default every rank tunes ON ITS OWN — no collective at all: the path has no
        data-path collective either, ranks that call this at the same point of their program measure each other's copy traffic
        anyway, and nothing can dead-lock. `collective=True` (every rank must then call it) makes all ranks keep the SAME layout
        with EXACTLY TWO all-reduces whatever the candidates are: the first makes every rank build the same candidate list (the
        list depends on the measured compute time per utterance, which differs from rank to rank: the maximum is used), the second
        takes, per candidate, the time of the slowest rank. Either way the ranks run their candidates free-running (a barrier per
        candidate puts the ranks in lock-step, where all copy at the same instant, which a service does not — at 8 GPUs
        5 / 27 / 27 / 5 takes 9.9 ms in lock-step and 8.2 free-running — and ranks whose candidate lists differed in length
        dead-locked an 8-GPU run of the first version). Leaves `out_host` filled with a valid result."""
        import time
        n = int(x_host.shape[0])
        sms = self._sm_count()
        in_b = 0.0 if x_host.is_cuda else float(x_host[0].numel() * x_host.element_size())      # a resident input is not copied
        out_b = float(out_host[0].numel() * out_host.element_size())
        utt_ms = self._agree_max([self._utterance_ms(fn, x_host)], collective)[0]                # collective 1 of 2
        cands = candidate_layouts(n, frames, heads, sms, utt_ms, in_b, out_b)[:self.AUTOTUNE_SLOTS]
        times = [float("inf")] * self.AUTOTUNE_SLOTS
        for ci, sizes in [(i, c) for _ in range(2) for i, c in enumerate(cands)]:      # two interleaved rounds, the better one counts
            self.sizes = list(sizes)
            self.run(fn, x_host, out_host)      # (re)allocates the chunk buffers of this layout
            self.synchronize()
            t0 = time.perf_counter()
            for _ in range(reps):
                self.run(fn, x_host, out_host)
                self.synchronize()
            times[ci] = min(times[ci], (time.perf_counter() - t0) / reps)
        agreed = self._agree_max([v if v != float("inf") else 1e30 for v in times], collective)  # collective 2 of 2
        best = min(range(len(cands)), key=lambda i: agreed[i])
        self.sizes = list(cands[best])
        self.tuned = {"candidates": cands, "ms": [round(agreed[i] * 1e3, 3) for i in range(len(cands))], "utt_ms": round(utt_ms, 4)}
        self.run(fn, x_host, out_host)
        self.synchronize()
        return self.sizes


def attention_makespan(k: int, frames: int, heads: int, sms: int = 148) -> float:
    """Time of the decoder's attention launch for a chunk of k utterances, in units of one two-query-tile CTA. The kernel
    (csrc/attention_h.cu) runs one CTA per SM, longest first: k * heads * (tiles // 2) CTAs with two 128-query tiles, then — when
    the tile count is odd — k * heads single-tile CTAs of half the duration."""
    tiles = (frames + 127) // 128
    doubles, singles = k * heads * (tiles // 2), k * heads * (tiles % 2)
    full, rem = divmod(doubles, sms)
    if rem == 0:
        return full + 0.5 * ((singles + sms - 1) // sms)
    # the last wave of two-tile CTAs leaves sms - rem SMs to the single-tile ones (two of them fit next to it)
    left = singles - 2 * (sms - rem)
    return full + 1.0 + (0.5 * ((left + sms - 1) // sms) if left > 0 else 0.0)


def wave_chunk_sizes(n: int, frames: int, heads: int, sms: int = 148) -> List[int]:
    """Chunk sizes for `HostPipeline(sizes=...)` on a GPU that has the PCIe link to itself: a small first and last chunk (their
    H2D / D2H copy is what nothing overlaps) around two inner ones, every chunk chosen so that the attention launch — a third of
    the step, one CTA per SM — fills its last wave of CTAs. For 64 utterances of 3446 frames on 148 SMs this gives 5 / 27 / 27 / 5
    (tools/e2e_sweep.py: 7.30 ms per host-to-host step against 7.85 for three equal chunks, whose 22-utterance chunk needs 4.5
    waves for 4.0 waves of work). With several ranks of one host copying at once the copies are slower and more, smaller chunks
    win: `HostPipeline.autotune` measures."""
    if n < 8:
        return [n]
    per_utt = attention_makespan(1024, frames, heads, sms) / 1024.0      # asymptotic cost of one utterance

    edges = [k for k in range(max(1, n // 16), max(2, n // 6) + 1)]
    best, best_score = [n], -1.0
    for e in edges:
        inner = n - 2 * e
        if inner < 2:
            continue
        for a in range(inner // 2, inner // 2 + 4):
            b = inner - a
            if a < 1 or b < 1:
                continue
            # efficiency of the whole layout, the edge chunks' exposed copies counted as lost time of their own size
            work = n * per_utt
            spent = sum(attention_makespan(k, frames, heads, sms) for k in (e, a, b, e)) + 0.25 * 2 * e * per_utt
            score = work / spent
            if score > best_score:
                best, best_score = [e, a, b, e], score
    return best


def predict_ms(sizes: Sequence[int], frames: int, heads: int, sms: int, copy_gbps: float, utt_ms: float, in_bytes: float, out_bytes: float,
               ramp_ms: float = 0.15, att_share: float = 0.32) -> float:
    """Model of one `HostPipeline.run`: three queues (H2D copies, compute, D2H copies), chunk i computes when its input has
    arrived and chunk i-1 is done, and is copied back when it is done and chunk i-1 has been copied. Compute of a chunk =
    `ramp_ms` + its attention launch in whole waves of CTAs (`attention_makespan`) + the rest in proportion to its size;
    `utt_ms` = compute per utterance of a large batch, `in_bytes` / `out_bytes` per utterance, `copy_gbps` what one copy stream
    gets. Against tools/e2e_sweep.py (one GPU, 55-60 GB/s) and tools/e2e_sweep_dist.py (eight ranks at once: 20-22 GB/s each) the
    model ranks 14 / 18 measured layouts with correlation 0.96 / 0.98 and an RMS error of 0.15 / 0.23 ms."""
    n = sum(sizes)
    big = max(n, 8 * sms)
    t_unit = att_share * utt_ms * big / attention_makespan(big, frames, heads, sms)
    t_other = (1.0 - att_share) * utt_ms
    th, td = in_bytes / (copy_gbps * 1e6), out_bytes / (copy_gbps * 1e6)
    h = c = d = 0.0
    for k in sizes:
        h += k * th + 0.02
        c = max(c, h) + ramp_ms + attention_makespan(k, frames, heads, sms) * t_unit + k * t_other
        d = max(d, c) + k * td
    return d


def candidate_layouts(n: int, frames: int, heads: int, sms: int, utt_ms: float, in_bytes: float, out_bytes: float) -> List[List[int]]:
    """A handful of chunk layouts worth measuring: for each assumed copy bandwidth (a link of its own ... eight ranks behind one
    host) the two layouts `predict_ms` likes best among 3-6 chunks built from sizes that fill the attention kernel's waves."""
    if n < 8:
        return [[n]]
    per_utt = attention_makespan(1024, frames, heads, sms) / 1024.0
    good = [k for k in range(1, n) if k * per_utt / attention_makespan(k, frames, heads, sms) >= 0.9] or list(range(1, n))

    def nearest(v: float, hi: int) -> int:
        c = [k for k in good if k <= hi] or [max(1, hi)]
        return min(c, key=lambda k: abs(k - v))

    layouts = {tuple(wave_chunk_sizes(n, frames, heads, sms)), tuple(hi - lo for lo, hi in HostPipeline.bounds(n, 3)),
               tuple(hi - lo for lo, hi in HostPipeline.bounds(n, 4))}
    firsts = sorted({nearest(n / 12.0, n // 3), nearest(n / 6.0, n // 3), nearest(n / 4.0, n // 3)})
    lasts = sorted({nearest(n / 12.0, n // 4), nearest(n / 6.0, n // 4)})
    for chunks in (3, 4, 5, 6):
        for f in firsts:
            for l in lasts:
                inner, m = n - f - l, chunks - 2
                if inner < m:
                    continue
                mid, left = [], inner
                for j in range(m):
                    k = left if j == m - 1 else nearest(left / (m - j), left - (m - j - 1))
                    mid.append(k)
                    left -= k
                if min(mid) >= 1:
                    layouts.add(tuple([f] + sorted(mid) + [l]))
                    layouts.add(tuple([f] + sorted(mid, reverse=True) + [l]))
    out: List[List[int]] = []
    for bw in (56.0, 36.0, 24.0, 16.0):
        ranked = sorted(layouts, key=lambda t: predict_ms(t, frames, heads, sms, bw, utt_ms, in_bytes, out_bytes))
        for t in ranked[:2]:
            if list(t) not in out:
                out.append(list(t))
    return out


def synthesize_to_host(model, ids_host: torch.Tensor, lengths_host: Optional[torch.Tensor], durations_host: Optional[torch.Tensor],
                       max_target_length: int, out_host: torch.Tensor, pipe: HostPipeline) -> None:
    """The whole model from HOST phoneme ids to a HOST waveform: `out_host[b] = model(ids[b], ...)["audio_output"]`.

    The acoustic front (text encoder, duration predictor, length regulator: < 5 % of the work, launch-bound) runs ONCE over the
    whole batch; decoder + vocoder then run in `pipe.n_chunks` utterance chunks so each chunk's waveform copy overlaps the next
    chunk's compute (the inputs are a few hundred KB: nothing to hide on the way in). Same results as one `model.forward`
    (utterances are independent in eval mode given a shared `max_target_length`). Returns after enqueueing —
    call `pipe.synchronize()` before reading `out_host`. Reference call site: scripts/synthesize.py:66-83."""
    if model.training:
        raise RuntimeError("synthesize_to_host runs the eval-mode path: call model.eval() first")
    dev = pipe.device
    ids = ids_host.to(dev, non_blocking=True)
    lens = None if lengths_host is None else lengths_host.to(dev, non_blocking=True)
    durs = None if durations_host is None else durations_host.to(dev, non_blocking=True)
    enc, _ = model.text_encoder(ids, lens)
    pred = model.duration_predictor(enc)
    reg = model.length_regulator(enc, durs if durs is not None else pred, max_target_length)
    pipe.run(lambda r: model.vocoder(model.decoder(r).transpose(1, 2)), reg, out_host)      # resident input: only the waveforms are copied
