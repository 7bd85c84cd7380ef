"""Bring-up helper (GPU box): phase timestamps of the warp-specialised attention (attention_h.cu), CTA 0: softmax warpgroup 0 of
query tile A (the even key tiles from 8 on)."""
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
# softmax warpgroup 0 of query tile A owns the even key tiles; rows = tiles 8, 10, ... (54 key tiles at L = 3446 -> 23 rows)
# [0] before wait S, [1] S ready, [2] scores loaded + row maximum, [3] hand-off + m_ref decision, [4] exps + split + P stores issued, [5] stores done
n = 23
w = prof.cpu()[:8 * n].view(n, 8)
names = ["wait S", "S ld + max", "hand-off + decision", "exp + split + P st issue", "P st wait"]
segs = [(w[:, k + 1] - w[:, k]).float().mean().item() for k in range(5)]
per = (w[n - 1, 0] - w[0, 0]).item() / (n - 1)
print("softmax warpgroup A0, per own key tile = 2 key tiles of the CTA (cycles): " + ", ".join(f"{n_}={v:.0f}" for n_, v in zip(names, segs)) +
      f", period={per:.0f} (= {per / 2:.0f} per key tile)")
