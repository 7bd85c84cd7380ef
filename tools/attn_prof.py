"""Bring-up helper (GPU box): phase timestamps of the warp-specialised attention (CTA 0: softmax warpgroup of query tile A
and the UMMA issuer), key tiles 8..39."""
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
lib = nat.lib()
lib.m2tts_attention_set_prof.argtypes = [C.c_void_p]
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(64, 3446, 96, device="cuda")
prof = torch.zeros(2 * 48 * 8, dtype=torch.int64, device="cuda")
m.decoder(x)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.decoder(x)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
p = prof.cpu().view(2, 48, 8)
w, i = p[0, :32], p[1, :32]
names = ["wait S", "S ld", "max (+rescale)", "exp + pack + P st issue", "P st wait", "wait PV(t-1)"]
segs = [(w[:, k + 1] - w[:, k]).float().mean().item() for k in range(6)]
loop = (w[1:, 0] - w[:-1, 6]).float().mean().item()
print("softmax warpgroup A, per key tile (cycles): " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, segs)) +
      f", arrive + loop={loop:.0f}, total={(w[31, 0] - w[0, 0]).item() / 31:.0f}")
inames = ["wait P(A)", "issue PV(A)+QK(A)", "gap", "wait P(B)", "issue PV(B)+QK(B)"]
isegs = [(i[:, k + 1] - i[:, k]).float().mean().item() for k in range(5)]
iloop = (i[1:, 0] - i[:-1, 5]).float().mean().item()
print("issuer, per key tile (cycles): " + ", ".join(f"{n}={v:.0f}" for n, v in zip(inames, isegs)) + f", loop={iloop:.0f}, total={(i[31, 0] - i[0, 0]).item() / 31:.0f}")
# offset between the softmax arrival and the issuer seeing it
print("P(A) arrive -> issuer wake (cycles):", ((i[:, 1] - w[:, 6]).float().mean().item()))
