"""Bring-up helper (GPU box): is the attention kernel clock- (power-) bound? Times one decoder layer's attention launch (the
library's per-launch CUDA events) and its effective SM clock (clock64 / globaltimer of every CTA, tools build) back to back
and with idle gaps between launches."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import sys
import time
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
kw = dict(STAGE_KWARGS["stage2"]); kw["decoder_layers"] = 1
m = M2TTSModel(**kw).eval().cuda()
B, T = 64, 3446
x = torch.randn(B, T, 96, device="cuda")
n_cta = ((T + 255) // 256) * 2 * B
prof = torch.zeros(1024 + 5 * n_cta, dtype=torch.int64, device="cuda")
lib.m2tts_attention_set_prof(prof.data_ptr())


def one(gap_s):
    res = []
    for _ in range(8):
        if gap_s:
            time.sleep(gap_s)
        nat.stage_timing_enable(True)
        m.decoder(x)
        torch.cuda.synchronize()
        nat.stage_timing_enable(False)
        st = nat.stage_timing_read()
        w = prof.cpu()[1024:].view(n_cta, 5)
        clk = (w[:, 4] - w[:, 3]).sum().item() / (w[:, 1] - w[:, 0]).sum().item() * 1e3
        res.append((st["attention"][0], clk))
    return res


for _ in range(10):
    m.decoder(x)
torch.cuda.synchronize()
for gap in (0.0, 0.05, 0.5):
    r = one(gap)
    print(f"gap {gap:.2f} s: attention ms " + " ".join(f"{a:.3f}" for a, _ in r) + " | effective SM MHz " + " ".join(f"{c:.0f}" for _, c in r))
# sustained: 200 decoder calls back to back, then one measured
for _ in range(200):
    m.decoder(x)
r = one(0.0)
print("after 200 back-to-back calls: attention ms " + " ".join(f"{a:.3f}" for a, _ in r) + " | effective SM MHz " + " ".join(f"{c:.0f}" for _, c in r))
lib.m2tts_attention_set_prof(None)
