"""Bring-up helper (GPU box): the decoder (attention = 2/3 of it) in a sustained loop — ms per call once the board sits at its
power cap. usage: python tools/attn_sustained.py [seconds]"""
import _env  # noqa: F401
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(64, 3446, 96, device="cuda")
secs = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
for _ in range(5):
    m.decoder(x)
torch.cuda.synchronize()
t0 = time.perf_counter(); n = 0
while time.perf_counter() - t0 < secs:
    for _ in range(5):
        m.decoder(x)
    torch.cuda.synchronize(); n += 5
dt = time.perf_counter() - t0
nat.stage_timing_enable(True)
m.decoder(x); torch.cuda.synchronize()
nat.stage_timing_enable(False)
st = nat.stage_timing_read()
print(f"decoder sustained: {dt / n * 1e3:.3f} ms per call; last call attention {st['attention'][0]:.3f} ms")
