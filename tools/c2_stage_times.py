"""Bring-up helper (GPU box): per-stage device time of the C2 configuration (stage-1 model, batch 16, 64 phonemes, T = 314)."""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage1"]).eval().cuda()
g = torch.Generator().manual_seed(0)
ids = torch.randint(0, 256, (16, 64), generator=g).cuda()
lengths = torch.randint(32, 65, (16,), generator=g).cuda()
dur = torch.randint(1, 9, (16, 64), generator=g).float().cuda()
for _ in range(5):
    m(ids, lengths, target_durations=dur)
torch.cuda.synchronize()
for r in range(3):
    nat.stage_timing_enable(True)
    t0 = time.perf_counter()
    m(ids, lengths, target_durations=dur)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    nat.stage_timing_enable(False)
    st = nat.stage_timing_read()
    print(f"rep {r}: wall {1e3 * (t1 - t0):.3f} ms, device sum {sum(v[0] for v in st.values()):.3f} ms | " +
          " ".join(f"{k}={v[0]:.3f}/{v[1]}" for k, v in sorted(st.items(), key=lambda kv: -kv[1][0])))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    m(ids, lengths, target_durations=dur)
e1.record()
torch.cuda.synchronize()
print(f"eager, no timers: {e0.elapsed_time(e1) / 20:.3f} ms per forward")
