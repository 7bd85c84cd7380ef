"""HostPipeline chunk bounds (host logic, no GPU): contiguous, non-empty, cover the batch; `edge` shrinks the first and last
chunk (their copies are the ones nothing overlaps)."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
from utils.host_pipeline import HostPipeline  # noqa: E402


@pytest.mark.parametrize("n,chunks,edge", [(64, 1, 1.0), (64, 3, 1.0), (64, 3, 0.5), (64, 4, 0.4), (64, 5, 0.25), (5, 3, 0.5), (3, 3, 0.5),
                                           (2, 3, 0.5), (1, 4, 0.3), (7, 4, 0.1), (512, 8, 0.5)])
def test_bounds_cover_the_batch(n, chunks, edge):
    b = HostPipeline.bounds(n, chunks, edge)
    assert b[0][0] == 0 and b[-1][1] == n
    assert all(hi > lo for lo, hi in b)
    assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    assert len(b) == min(chunks, n)


def test_edge_chunks_are_smaller():
    sizes = [hi - lo for lo, hi in HostPipeline.bounds(64, 3, 0.5)]
    assert sizes == [16, 32, 16]
    assert [hi - lo for lo, hi in HostPipeline.bounds(64, 3, 1.0)] == [22, 21, 21]


def test_bounds_of_an_empty_batch():
    from utils.host_pipeline import HostPipeline
    assert HostPipeline.bounds(0, 3) == [] and HostPipeline.bounds(0, 3, 0.5) == []


def test_wave_chunk_sizes_cover_the_batch_and_fill_attention_waves():
    """utils.host_pipeline.wave_chunk_sizes: positive sizes that sum to the batch; for the C3 shape (64 x 3446 frames, 2 heads,
    148 SMs) every chunk's attention launch ends on a full wave of CTAs (makespan = work rounded up to whole waves)."""
    from utils.host_pipeline import attention_makespan, wave_chunk_sizes
    for n in (1, 7, 8, 16, 31, 64, 65, 128, 512):
        for frames, heads in ((3446, 2), (314, 2), (2048, 4), (100, 2)):
            sizes = wave_chunk_sizes(n, frames, heads)
            assert sum(sizes) == n and min(sizes) >= 1, (n, frames, heads, sizes)
    assert wave_chunk_sizes(64, 3446, 2) == [5, 27, 27, 5]
    # 27 query tiles per (utterance, head): 13 two-tile CTAs + 1 single-tile CTA of half the duration
    assert attention_makespan(64, 3446, 2) == 12.0      # 1664 two-tile CTAs = 11 waves + 36, the 128 single-tile CTAs fit beside them
    assert attention_makespan(21, 3446, 2) == 4.0 and attention_makespan(22, 3446, 2) == 4.5
    assert attention_makespan(5, 3446, 2) == 1.0 and attention_makespan(27, 3446, 2) == 5.0


def test_explicit_sizes_are_used_only_when_they_fit_the_batch():
    import torch
    from utils.host_pipeline import HostPipeline
    pipe = HostPipeline.__new__(HostPipeline)      # no CUDA streams on the CPU box: only the chunk logic is exercised
    pipe.n_chunks, pipe.edge, pipe.sizes = 3, 1.0, [5, 27, 27, 5]
    assert [hi - lo for lo, hi in pipe.chunk_bounds(64)] == [5, 27, 27, 5]
    assert pipe.chunk_bounds(64)[0] == (0, 5) and pipe.chunk_bounds(64)[-1] == (59, 64)
    assert [hi - lo for lo, hi in pipe.chunk_bounds(10)] == [4, 3, 3]
    del torch


def test_candidate_layouts_and_the_pipeline_model():
    """`candidate_layouts`: every candidate is a positive composition of the batch; `predict_ms` reproduces the two measured
    regimes (tools/e2e_sweep.py on one GPU, tools/e2e_sweep_dist.py with eight ranks of one host copying at once)."""
    from utils.host_pipeline import candidate_layouts, predict_ms
    for n, frames, heads in ((64, 3446, 2), (16, 314, 2), (9, 300, 2), (512, 3446, 2), (3, 100, 2), (65, 2048, 4)):
        cands = candidate_layouts(n, frames, heads, 148, 0.1, 1.3e6, 0.88e6)
        assert 1 <= len(cands) <= 8
        for c in cands:
            assert sum(c) == n and min(c) >= 1, (n, c)
    kw = dict(frames=3446, heads=2, sms=148, utt_ms=0.102, in_bytes=1.3233e6, out_bytes=0.8822e6)
    alone = {t: predict_ms(t, copy_gbps=56.0, **kw) for t in ((5, 27, 27, 5), (22, 21, 21), (10, 16, 16, 17, 5), (5, 54, 5), (64,))}
    assert alone[(5, 27, 27, 5)] == min(alone.values())                       # measured: 7.30 against 7.85 / 7.67 / 8.04 / 8.93 ms
    assert alone[(64,)] == max(alone.values())
    shared = {t: predict_ms(t, copy_gbps=21.0, **kw) for t in ((5, 27, 27, 5), (22, 21, 21), (10, 16, 16, 17, 5), (10, 54))}
    assert shared[(10, 16, 16, 17, 5)] == min(shared.values())                # measured at 8 GPUs: 8.51 against 9.39 / 9.20 / 11.94 ms
    assert shared[(10, 54)] == max(shared.values())
    assert [5, 27, 27, 5] in candidate_layouts(64, 3446, 2, 148, 0.102, 1.3233e6, 0.8822e6)
