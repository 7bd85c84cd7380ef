"""Profiling helper (GPU box): run the stage2 vocoder alone at the C3 size so ncu can capture its kernels.
usage: python tools/voc_only.py [B] [T] [reps]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from models.tts_model import M2TTSModel  # noqa: E402
from models.stage_configs import STAGE_KWARGS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3446
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
mel = torch.randn(B, T, 80, device="cuda")
for _ in range(reps):
    y = m.vocoder(mel.transpose(1, 2))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
y = m.vocoder(mel.transpose(1, 2))
e1.record()
torch.cuda.synchronize()
print(f"vocoder B={B} T={T}: {e0.elapsed_time(e1):.3f} ms, out {tuple(y.shape)}")
