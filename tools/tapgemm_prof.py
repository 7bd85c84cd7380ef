"""Bring-up helper (GPU box): per-tile wait / epilogue cycles of the tap-GEMM convolution (CTA 0, epilogue group 0)."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
for (CI, CO, L, res) in ((64, 64, 55136, True), (64, 64, 55136, False), (128, 128, 13784, True)):
    B = 64
    x = torch.randn(B, CI, L, device="cuda")
    w = torch.randn(CO, CI, 3, device="cuda") * 0.1
    b = torch.zeros(CO, device="cuda")
    r = torch.randn(B, CO, L, device="cuda") if res else None
    y = torch.empty(B, CO, L, device="cuda")
    ws = torch.empty(lib.m2tts_conv_tc_workspace_bytes(B, CI, CO, L, 1), dtype=torch.uint8, device="cuda")
    prof = torch.zeros(128 * 4 + 64, dtype=torch.int64, device="cuda")
    for i in range(2):
        lib.m2tts_tapgemm_set_prof(prof.data_ptr() if i == 1 else None)
        nat.check(lib.m2tts_conv1d_k3_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), nat.ptr(r), y.data_ptr(), B, CI, CO, L, 1, 0,
                                          ws.data_ptr(), ws.numel(), None), "conv")
    torch.cuda.synchronize()
    lib.m2tts_tapgemm_set_prof(None)
    sp = prof.cpu()[512:512 + 32].view(8, 4)
    p = prof.cpu()[:512].view(128, 4)
    rows = p[(p[:, 0] > 0)][2:40]          # tiles handled by group 0 (every other t)
    wait = (rows[:, 1] - rows[:, 0]).float().mean().item()
    epi = (rows[:, 2] - rows[:, 1]).float().mean().item()
    period = (rows[1:, 0] - rows[:-1, 0]).float().mean().item()
    nc = CI // 16
    print("   splitter (9th tile) per chunk: raw wait", [(int(sp[c, 1] - sp[c, 0])) for c in range(nc)], "slot wait", [(int(sp[c, 2] - sp[c, 1])) for c in range(nc)],
          "split", [(int(sp[c, 3] - sp[c, 2])) for c in range(nc)])
    print(f"CI={CI} CO={CO} L={L} residual={res}: accumulator wait {wait:.0f}, epilogue {epi:.0f}, period per group tile {period:.0f} cycles (2 groups alternate)")
