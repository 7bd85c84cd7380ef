"""Bring-up helper (GPU box): phase timestamps of the persistent linear kernel's first epilogue warp (CTA 0), per unit (= one
N pass of one 128-row tile). M2TTS_LIN_PROF_STAGE=<stage id> selects the layer (2 ln_qkv, 4 out_proj, 5 ffn1, 6 ffn2, 7 ln_proj;
see include/m2tts_b200.h); the last launch of that stage in one decoder call is reported."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch
from models import _native as nat
lib = nat.tools_lib()
# one QKV-shaped GEMM through the layernorm_proj entry (mode 0) is not mode 3; run a full layer and keep the QKV launch by
# stopping after it: simplest is to run the decoder and read the buffer after the first layer's QKV only -> we use a 1-layer model
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(64, 3446, 96, device="cuda")
prof = torch.zeros(2 * 48 * 8, dtype=torch.int64, device="cuda")
m.decoder(x)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.decoder(x)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
p = prof.cpu().view(2, 48, 8)[0]
n = 1
while n < 48 and p[n, 0] > p[n - 1, 0] and p[n, 0] - p[n - 1, 0] < 10_000_000:      # units of the last launch (later entries are stale)
    n += 1
d = p[1:n]
names = ["staging free (barrier) + residual load issue", "wait accumulator + residual", "chunks", "fence + barrier + store issue"]
segs = [(d[:, k + 1] - d[:, k]).float().mean().item() for k in range(4)]
print(f"{n} units of CTA 0, cycles per unit: " + ", ".join(f"{a}={v:.0f}" for a, v in zip(names, segs)) +
      f", unit period={(p[n - 1, 0] - p[1, 0]).item() / max(n - 2, 1):.0f}, first unit: wait acc={int(p[0, 1] - p[0, 0])}")
