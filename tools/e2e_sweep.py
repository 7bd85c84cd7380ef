"""Bring-up helper (GPU box): host -> host time of the C3 step (utils.host_pipeline.HostPipeline) for several chunk layouts.
usage: python tools/e2e_sweep.py"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
from utils.host_pipeline import HostPipeline
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)
x_host = torch.randn(64, 3446, 96).pin_memory()
out = torch.empty(64, 1, 3446 * 64).pin_memory()
step = lambda x: m.vocoder(m.decoder(x).transpose(1, 2))
x_dev = x_host.to(dev)
for _ in range(3):
    step(x_dev)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with nat.deferred_status():
    e0.record()
    for _ in range(5):
        step(x_dev)
    e1.record()
    torch.cuda.synchronize()
    print(f"device-resident step: {e0.elapsed_time(e1) / 5:.3f} ms")
    for chunks, edge in ((1, 1.0), (2, 1.0), (3, 1.0), (3, 0.6), (4, 1.0), (4, 0.5), (4, 0.7), (5, 0.5), (5, 0.7), (6, 0.5)):
        pipe = HostPipeline(dev, n_chunks=chunks, edge=edge)
        for _ in range(2):
            pipe.run(step, x_host, out); pipe.synchronize()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(8):
            pipe.run(step, x_host, out); pipe.synchronize()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 8 * 1e3
        print(f"chunks={chunks} edge={edge}: sizes {[hi - lo for lo, hi in HostPipeline.bounds(64, chunks, edge)]} -> {ms:.3f} ms per step")
    import statistics
    cands = ([22, 21, 21], [16, 16, 16, 16], [5, 27, 27, 5], [10, 16, 27, 11], [5, 16, 27, 16], [5, 22, 32, 5], [5, 27, 22, 10], [5, 16, 16, 22, 5],
             [5, 21, 28, 10], [5, 32, 27], [3, 27, 29, 5], [5, 27, 27, 3, 2], [2, 3, 27, 27, 5], [5, 54, 5], [5, 21, 16, 16, 6], [8, 24, 24, 8])
    res = {tuple(c): [] for c in cands}
    pipes = {tuple(c): HostPipeline(dev, sizes=c) for c in cands}
    for rnd in range(6):
        for c in cands:
            pipe = pipes[tuple(c)]
            pipe.run(step, x_host, out); pipe.synchronize()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(5):
                pipe.run(step, x_host, out); pipe.synchronize()
            torch.cuda.synchronize()
            res[tuple(c)].append((time.perf_counter() - t0) / 5 * 1e3)
    for c in cands:
        v = res[tuple(c)]
        print(f"sizes {c}: median {statistics.median(v):.3f} min {min(v):.3f} max {max(v):.3f}")
nat.check_status(dev, "e2e sweep")
