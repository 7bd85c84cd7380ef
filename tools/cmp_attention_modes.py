"""Bring-up helper (GPU box): decoder output of the 16-bit split (default precision) against the TF32 split for sequence
lengths around the key-tile boundaries."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
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
for L in (1, 64, 65, 128, 129, 192, 193, 256, 257, 300, 512, 700):
    x = torch.randn(2, L, 96, device="cuda")
    with nat.precision("tf32"):
        a = m.decoder(x).clone()
    with nat.precision("split16"):
        b = m.decoder(x).clone()
    torch.cuda.synchronize()
    d=(a-b).abs()
    print(L, float(d.max()), "rows bad:", (d.amax(dim=(0,2))>1e-3).nonzero().flatten().tolist()[:12])
