"""Bring-up helper (GPU box): does splitting the resident C3 batch over two streams fill the kernels' tails? One stream with 64
utterances against two streams with 32 + 32 (and 27 + 37, 21 + 43: sizes whose attention launches end on full waves)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)
x = torch.randn(64, 3446, 96, device=dev)
step = lambda t: m.vocoder(m.decoder(t).transpose(1, 2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run_split(sizes, reps=10):
    parts = list(torch.split(x, sizes))
    streams = [s1, s2]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i, p in enumerate(parts):
            with torch.cuda.stream(streams[i % 2]):
                step(p)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def run_single(reps=10):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        step(x)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


with nat.deferred_status():
    for _ in range(3):
        step(x)
    run_split([32, 32], 2)
    for rnd in range(3):
        print(f"round {rnd}: one stream 64: {run_single():.3f} ms | two streams 32+32: {run_split([32, 32]):.3f} | 27+37: {run_split([27, 37]):.3f} | "
              f"21+43: {run_split([21, 43]):.3f} | 16+16+16+16: {run_split([16, 16, 16, 16]):.3f} | 32+32 on one stream: ", end="")
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            step(x[:32]); step(x[32:])
        torch.cuda.synchronize(); print(f"{(time.perf_counter() - t0) / 10 * 1e3:.3f}")
nat.check_status(dev, "two-stream probe")
