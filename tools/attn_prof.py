"""Bring-up helper (GPU box): phase timestamps of the warp-specialised attention (attention_h.cu), CTA 0: the first softmax
warpgroup of query tile A, key tiles 8..39 (of 54 at the C3 length)."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
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
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = torch.randn(B, 3446, 96, device="cuda")
prof = torch.zeros(2 * 48 * 8, dtype=torch.int64, device="cuda")
m.decoder(x)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.decoder(x)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
w = prof.cpu()[:256].view(32, 8)      # [0] before wait S, [1] S ready, [2] scores loaded, [3] max + exchange (+ rescale), [4] exps + P stores issued, [5] stores done, [6] PV(t-1) seen
names = ["wait S", "S ld", "max + exchange", "exp + pack + P st issue", "P st wait", "wait PV(t-1)"]
segs = [(w[:, k + 1] - w[:, k]).float().mean().item() for k in range(6)]
print("softmax warpgroup A0, per 64-key tile (cycles): " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, segs)) +
      f", total={(w[31, 0] - w[0, 0]).item() / 31:.0f}")
