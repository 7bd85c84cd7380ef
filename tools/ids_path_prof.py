"""Bring-up helper (GPU box): where the time of the from-phoneme-ids path goes (bench.py's `e2e_from_ids`): per module, device
time by CUDA events and host time to enqueue (a module whose enqueue time exceeds its device time starves the GPU).
usage: python tools/ids_path_prof.py [B] [S] [frames]"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    sys.path.insert(0, p)
import torch  # noqa: E402
from models import _native as nat  # noqa: E402
from models.tts_model import M2TTSModel  # noqa: E402
from models.stage_configs import STAGE_KWARGS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 22
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
FRAMES = int(sys.argv[3]) if len(sys.argv) > 3 else 3446
dev = torch.device("cuda", 0)
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)
g = torch.Generator().manual_seed(100)
ids = torch.randint(0, 256, (B, S), generator=g).to(dev)
lens = torch.full((B,), S, dtype=torch.int64, device=dev)
per = FRAMES // S
dur = torch.full((B, S), float(per))
for b in range(B):
    dur[b, torch.randperm(S, generator=g)[:FRAMES - per * S]] = per + 1.0
dur = (dur + 0.5).to(dev)


def stages():
    enc, mask = m.text_encoder(ids, lens)
    yield "text_encoder"
    d = m.duration_predictor(enc)
    yield "duration_predictor"
    reg = m.length_regulator(enc, dur, FRAMES)
    yield "length_regulator"
    mel = m.decoder(reg)
    yield "decoder"
    m.vocoder(mel.transpose(1, 2))
    yield "vocoder"


for _ in range(3):
    for _ in stages():
        pass
torch.cuda.synchronize()
reps = 10
host = {}
devt = {}
with nat.deferred_status():
    for _ in range(reps):
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True)]
        ev[0].record()
        t0 = time.perf_counter()
        names = []
        for name in stages():
            t1 = time.perf_counter()
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            ev.append(e)
            names.append(name)
            host[name] = host.get(name, 0.0) + (t1 - t0) * 1e3
            t0 = time.perf_counter()
        torch.cuda.synchronize()
        for i, name in enumerate(names):
            devt[name] = devt.get(name, 0.0) + ev[i].elapsed_time(ev[i + 1])
nat.check_status(dev, "ids_path_prof")
print(f"from-ids path, B={B} S={S} frames={FRAMES} (ms per call, mean of {reps}): module: device span | host enqueue")
for name in devt:
    print(f"  {name:20s} {devt[name] / reps:8.3f} | {host[name] / reps:8.3f}")
print(f"  {'total':20s} {sum(devt.values()) / reps:8.3f} | {sum(host.values()) / reps:8.3f}")
