"""Bring-up helper (GPU box): phase timestamps of the fused ResBlock kernel (voc_res_h.cu), CTA 0, first epilogue warp, per tile."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
lib = nat.tools_lib()
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
mel = torch.randn(64, 3446, 80, device="cuda").transpose(1, 2)
prof = torch.zeros(2 * 48 * 8, dtype=torch.int64, device="cuda")
m.vocoder(mel)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.vocoder(mel)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
p = prof.cpu().view(2, 48, 8)[0]
d = p[4:44]
names = ["wait conv1", "EPI2 (V)", "wait conv2", "EPI3 (residual, staging)", "fence + barrier + store issue"]
segs = [(d[:, k + 1] - d[:, k]).float().mean().item() for k in range(5)]
print("cycles per tile: " + ", ".join(f"{a}={v:.0f}" for a, v in zip(names, segs)) + f", tile period={(d[-1, 0] - d[0, 0]).item() / (len(d) - 1):.0f}")
