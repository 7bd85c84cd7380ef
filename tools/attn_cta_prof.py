"""Bring-up helper (GPU box): lifetime of every CTA of the persistent attention kernel (attention_h.cu, tools build) against
the steady-state key-tile period. The per-item phase stamps quoted in DESIGN.md section 4.1 (barrier init, Q_hi copy, first scores,
drain, O store of a mid-kernel CTA) were taken with this tool on the one-CTA-per-item kernel that preceded it (commits a5b6e82 / 4fa91de:
`git log -- tools/attn_cta_prof.py`)."""
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
kw = dict(STAGE_KWARGS["stage2"]); kw["decoder_layers"] = 1
m = M2TTSModel(**kw).eval().cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3446
x = torch.randn(B, T, 96, device="cuda")
import os
n_items = ((T + 255) // 256) * 2 * B
persist = True
n_cta = min(n_items, 148) if persist else n_items
prof = torch.zeros(1024 + 5 * n_cta, dtype=torch.int64, device="cuda")
m.decoder(x)
lib.m2tts_attention_set_prof(prof.data_ptr())
m.decoder(x)
torch.cuda.synchronize()
lib.m2tts_attention_set_prof(None)
w = prof.cpu()[1024:].view(n_cta, 5)
n_long = n_cta if persist else (((T + 127) // 128) // 2) * 2 * B
life = (w[:, 1] - w[:, 0]).float()
t0, t1 = w[:, 0].min().item(), w[:, 1].max().item()
print(f"B={B} T={T}: {n_cta} CTAs ({n_long} two-tile), kernel span {(t1 - t0) / 1e3:.1f} us; two-tile CTA lifetime mean {life[:n_long].mean() / 1e3:.2f} us "
      f"(min {life[:n_long].min() / 1e3:.2f}, max {life[:n_long].max() / 1e3:.2f}); single-tile mean {life[n_long:].mean() / 1e3 if n_long < n_cta else 0:.2f} us")
if persist:
    cyc = (w[:, 4] - w[:, 3]).float()
    print(f"persistent CTAs: lifetime mean {cyc.mean():.0f} cycles, effective SM clock {cyc.sum().item() / life.sum().item() * 1e3:.0f} MHz")
    sys.exit(0)
# gap between consecutive CTAs on one SM
gaps = []
for sm in range(int(w[:, 2].max().item()) + 1):
    rows = w[w[:, 2] == sm]
    if len(rows) < 2:
        continue
    rows = rows[rows[:, 0].argsort()]
    gaps.append((rows[1:, 0] - rows[:-1, 1]).float())
g = torch.cat(gaps)
per_sm = torch.bincount(w[:, 2])
print(f"gap between a CTA's exit and the next CTA's start on the same SM: mean {g.mean() / 1e3:.2f} us (min {g.min() / 1e3:.2f}, max {g.max() / 1e3:.2f}); CTAs per SM {per_sm.min().item()}..{per_sm.max().item()}")
first = prof.cpu()[:8 * 23].view(23, 8)
per = (first[22, 0] - first[0, 0]).item() / 22
print(f"steady-state period per key tile (CTA 0, clock cycles): {per / 2:.0f}; x {(T + 63) // 64} key tiles = {per / 2 * ((T + 63) // 64):.0f} cycles")
cyc = (w[:, 4] - w[:, 3]).float()
print(f"two-tile CTA lifetime in SM cycles: mean {cyc[:n_long].mean():.0f}; effective SM clock {cyc[:n_long].sum().item() / life[:n_long].sum().item() * 1e3:.0f} MHz")
print(f"CTA 0: start -> key tile 8 arrives at softmax: {first[0, 1].item() - w[0, 3].item()} cycles; tile 52 wait -> CTA end: {w[0, 4].item() - first[22, 0].item()} cycles; lifetime {w[0, 4].item() - w[0, 3].item()}")
if persist:
    sys.exit(0)
st = prof.cpu()[512:520]
names = ["setup (barrier init, TMEM alloc, sync)", "Q_hi -> TMEM", "first scores ready", "key-tile loop", "last P V landed", "O normalise + store", "final sync"]
print("CTA 600, softmax warpgroup A0, cycles: " + ", ".join(f"{n}={st[k + 1].item() - st[k].item()}" for k, n in enumerate(names)) + f"; total {st[7].item() - st[0].item()}; CTA lifetime {w[600, 4].item() - w[600, 3].item()}")
