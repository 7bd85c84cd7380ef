"""Bring-up helper (GPU box): host-side profile (cProfile) of LengthRegulator.forward in a deferred-status region."""
import cProfile
import pstats
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
B, S, FR = 22, 256, 3446
enc = torch.randn(B, S, 96, device="cuda")
dur = torch.full((B, S), 13.5, device="cuda")
for _ in range(3):
    m.length_regulator(enc, dur, FR)
torch.cuda.synchronize()
with nat.deferred_status():
    t0 = time.perf_counter()
    for _ in range(20):
        m.length_regulator(enc, dur, FR)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"deferred: host {1e3 * (t1 - t0) / 20:.3f} ms per call")
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(20):
        m.length_regulator(enc, dur, FR)
    pr.disable()
    torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
t0 = time.perf_counter()
for _ in range(20):
    m.length_regulator(enc, dur, FR)
t1 = time.perf_counter()
print(f"eager (status read per call): host {1e3 * (t1 - t0) / 20:.3f} ms per call")
