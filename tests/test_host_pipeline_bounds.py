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
