"""C5 (BASELINE.json configs[4]): stage2 vocoder-only mel -> wave sweep, batch 1-256 x 128-2048 frames, on one B200.
Prints one line per (B, T): device ms (CUDA events, median of 5 after 2 warm-ups) and audio-seconds per second.
usage: python tools/sweep_vocoder.py [out.csv]"""
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from models.tts_model import M2TTSModel  # noqa: E402
from models.stage_configs import STAGE_KWARGS  # noqa: E402

torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
rows = ["batch,frames,ms,audio_s_per_s"]
for B in (1, 2, 4, 8, 16, 32, 64, 128, 256):
    for T in (128, 256, 512, 1024, 2048):
        mel = torch.randn(B, 80, T, device="cuda")
        for _ in range(2):
            m.vocoder(mel)
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            m.vocoder(mel)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = statistics.median(ts)
        rows.append(f"{B},{T},{ms:.4f},{B * T * 64 / 22050 / (ms * 1e-3):.1f}")
        print(rows[-1], flush=True)
if len(sys.argv) > 1:
    Path(sys.argv[1]).write_text("\n".join(rows) + "\n")
