"""Bring-up helper (GPU box): phase timestamps of the warp-specialised attention (CTA 0, query tile A)."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from oracle import m2tts_oracle as oracle
lib = nat.lib()
lib.m2tts_attention_set_prof.argtypes = [C.c_void_p]
torch.manual_seed(1234)
m = M2TTSModel(**oracle.STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(64, 3446, 96, device="cuda")
prof = torch.zeros(48 * 8, dtype=torch.int64, device="cuda")
m.decoder(x)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.decoder(x)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
p = prof.cpu().view(48, 8)
d = p[8:40]
names = ["S ld", "softmax+P st"]
segs = [(d[:, i + 1] - d[:, i]).float().mean().item() for i in range(2)] + [(p[9:41, 0] - p[8:40, 2]).float().mean().item()]
print("per-tile cycles " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names + ["wait next S"], segs)) + f"  total={(p[40, 0] - p[8, 0]).item() / 32:.0f}")
