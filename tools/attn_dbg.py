"""Bring-up helper (GPU box): time of one attention launch at the C3 shape with parts of the kernel switched off
(M2TTS_ATT_DBG: 1 no Q K^T UMMAs, 2 no P V UMMAs, 3 neither; results are invalid while set) — run once per setting:
    for d in 0 1 2 3; do M2TTS_ATT_DBG=$d python tools/attn_dbg.py; done"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import os
import sys
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, 3446, 96, device="cuda")
with nat.deferred_status():
    for _ in range(2):
        m.decoder(x)
    torch.cuda.synchronize()
    nat.stage_timing_enable(True)
    for _ in range(3):
        m.decoder(x)
    torch.cuda.synchronize()
    nat.stage_timing_enable(False)
t = nat.stage_timing_read()
nat.read_status(x.device)
ms, n = t["attention"]
print(f"M2TTS_ATT_DBG={os.environ.get('M2TTS_ATT_DBG', '0')} M2TTS_ATT_OLD={os.environ.get('M2TTS_ATT_OLD', '0')}: attention {ms / n:.4f} ms per launch ({n} launches)")
