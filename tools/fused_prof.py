"""Bring-up helper (GPU box): phase timestamps of the fused vocoder stage (CTA 0, context 0)."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
for Cc, final in ((32, False), (16, True)):
    B, L = 64, (55136 if Cc == 32 else 110272)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, L, 2 * Cc, device="cuda")
    up_w = torch.randn(2 * Cc, Cc, 4, device="cuda") * 0.1
    w1 = torch.randn(Cc, Cc, 3, device="cuda") * 0.1
    w2 = torch.randn(Cc, Cc, 3, device="cuda") * 0.1
    b = torch.zeros(Cc, device="cuda")
    ow = torch.randn(1, Cc, 3, device="cuda") * 0.1
    ob = torch.zeros(1, device="cuda")
    y = torch.empty((B, 2 * L) if final else (B, 2 * L, Cc), device="cuda")
    ws = torch.empty(lib.m2tts_vocoder_stage_fused_workspace_bytes(Cc), dtype=torch.uint8, device="cuda")
    prof = torch.zeros(64 * 8, dtype=torch.int64, device="cuda")
    lib.m2tts_vocoder_stage_fused_set_prof(prof.data_ptr())
    for _ in range(2):
        rc = lib.m2tts_vocoder_stage_fused(x.data_ptr(), up_w.data_ptr(), b.data_ptr(), w1.data_ptr(), b.data_ptr(), w2.data_ptr(), b.data_ptr(),
                                           ow.data_ptr() if final else None, ob.data_ptr() if final else None, y.data_ptr(), B, Cc, L,
                                           ws.data_ptr(), ws.numel(), None)
        nat.check(rc, "fused")
    torch.cuda.synchronize()
    lib.m2tts_vocoder_stage_fused_set_prof(None)
    p = prof.cpu().view(64, 8)[:, :6]
    d = p[8:40]
    names = ["EPI1", "conv1 wait", "EPI2", "conv2 wait", "EPI3", "next up wait"]
    segs = [(d[:, i + 1] - d[:, i]).float().mean().item() for i in range(5)] + [(p[9:41, 0] - p[8:40, 5]).float().mean().item()]
    print(f"C={Cc}: per-tile cycles " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, segs)) + f"  total={(p[40, 0] - p[8, 0]).item() / 32:.0f}")
