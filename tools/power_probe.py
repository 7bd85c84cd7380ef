"""Bring-up helper (GPU box): SM clock and board power (NVML) while one part of the C3 step runs in a loop — which kernels run
against the power cap."""
import sys
import threading
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import pynvml
import torch
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().cuda()
x = torch.randn(64, 3446, 96, device="cuda")
mel = m.decoder(x).transpose(1, 2).contiguous()


def sample(fn, seconds=2.0):
    stop, clk, pw = [False], [], []

    def poll():
        while not stop[0]:
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            time.sleep(0.02)
    th = threading.Thread(target=poll)
    t0 = time.perf_counter()
    n = 0
    th.start()
    while time.perf_counter() - t0 < seconds:
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        n += 5
    dt = time.perf_counter() - t0
    stop[0] = True
    th.join()
    k = len(clk) // 2
    return dt / n * 1e3, sorted(clk[k:])[len(clk[k:]) // 2], sum(pw[k:]) / len(pw[k:]), max(pw)


for name, fn in (("decoder", lambda: m.decoder(x)), ("vocoder", lambda: m.vocoder(mel)), ("step", lambda: m.vocoder(m.decoder(x).transpose(1, 2)))):
    ms, c, p, pmax = sample(fn)
    print(f"{name}: {ms:.3f} ms per call, SM clock (median of the second half) {c} MHz, power mean {p:.0f} W max {pmax:.0f} W "
          f"(limit {pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3:.0f} W)")
