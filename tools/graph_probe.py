"""Bring-up helper (GPU box): can the decoder + vocoder step be captured in a CUDA graph, and what does replay buy for small
utterance chunks (where the ~75 launches of a step are launch-latency bound)?"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
dev = torch.device("cuda:0")
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)


def step(x):
    return m.vocoder(m.decoder(x).transpose(1, 2))


for B in (4, 10, 16, 32, 64):
    x = torch.randn(B, 3446, 96, device=dev)
    for _ in range(3):
        y_ref = step(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        y = step(x)
    g.replay()
    torch.cuda.synchronize()
    err = (y - y_ref).abs().max().item()

    def timeit(f, n=10):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n * 1e3
    te, tg = timeit(lambda: step(x)), timeit(g.replay)
    print(f"B={B:3d}: eager {te:.3f} ms, graph replay {tg:.3f} ms, max|diff| {err:.2e}")
