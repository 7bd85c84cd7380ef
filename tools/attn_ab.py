"""Bring-up helper (GPU box): one attention flavour of the tools library (the tools library): parity of
the decoder against the TF32 split around the key-tile boundaries, then the attention stage time at the C3 size.
usage: M2TTS_ATT_V=2 python tools/attn_ab.py [B]"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import os
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
worst = 0.0
for L in (1, 63, 64, 65, 127, 128, 129, 192, 193, 256, 257, 300, 512, 700, 1153):
    x = torch.randn(3, L, 96, device="cuda") * 3.0
    with nat.precision("tf32"):
        a = m.decoder(x).clone()
    with nat.precision("split16"):
        b = m.decoder(x).clone()
    torch.cuda.synchronize()
    d = float((a - b).abs().max())
    worst = max(worst, d)
    if d > 1e-4:
        print("MISMATCH L", L, d)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, 3446, 96, device="cuda")
for _ in range(3):
    m.decoder(x)
torch.cuda.synchronize()
nat.stage_timing_enable(True)
for _ in range(5):
    m.decoder(x)
torch.cuda.synchronize()
nat.stage_timing_enable(False)
st = nat.stage_timing_read()
print(f"ATT_V={os.environ.get('M2TTS_ATT_V', 'default')} worst |split16 - tf32| = {worst:.3e}; attention {st['attention'][0] / st['attention'][1]:.4f} ms per launch "
      f"({st['attention'][1]} launches); decoder stages: " + ", ".join(f"{k}={v[0] / 5:.3f}" for k, v in st.items() if v[1]))
