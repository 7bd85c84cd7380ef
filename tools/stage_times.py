"""Bring-up helper (GPU box): per-stage device time (the library's per-launch CUDA events) of decoder + vocoder at the C3 size,
several repetitions printed one by one (spread between repetitions = clock / power effects).
usage: python tools/stage_times.py [B] [reps]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(B, 3446, 96, device="cuda")
for _ in range(3):
    m.vocoder(m.decoder(x).transpose(1, 2))
torch.cuda.synchronize()
for r in range(reps):
    nat.stage_timing_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    m.vocoder(m.decoder(x).transpose(1, 2))
    e1.record()
    torch.cuda.synchronize()
    nat.stage_timing_enable(False)
    st = nat.stage_timing_read()
    print(f"rep {r}: step {e0.elapsed_time(e1):.3f} ms | " + " ".join(f"{k}={v[0]:.3f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1][0]) if v[1]))
