"""Bring-up helper (GPU box): time the upsampling kernel (voc_up_h.cu) with parts switched off.
usage: python tools/up_h_prof.py [CI] [L] [B]"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
CI = int(sys.argv[1]) if len(sys.argv) > 1 else 128
L = int(sys.argv[2]) if len(sys.argv) > 2 else 13784
B = int(sys.argv[3]) if len(sys.argv) > 3 else 64
x = torch.randn(B, L, CI, device="cuda")
w = torch.randn(CI, CI // 2, 8, device="cuda") * 0.05
b = torch.zeros(CI // 2, device="cuda")
y = torch.empty(B, 4 * L, CI // 2, device="cuda")
ws = torch.empty(lib.m2tts_conv_transpose_x4_h_workspace_bytes(B, CI, L), dtype=torch.uint8, device="cuda")
for mode, name in ((0, "full"), (1, "no stores"), (2, "no UMMAs"), (4, "no TMA loads"), (3, "no stores, no UMMAs"), (6, "no UMMAs, no loads"), (7, "nothing")):
    lib.m2tts_voc_up_h_set_debug(mode)
    for _ in range(2):
        nat.check(lib.m2tts_conv_transpose_x4_h(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, CI, L, ws.data_ptr(), ws.numel(), None), "up")
    torch.cuda.synchronize()
    nat.stage_timing_enable(True)
    for _ in range(5):
        nat.check(lib.m2tts_conv_transpose_x4_h(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), B, CI, L, ws.data_ptr(), ws.numel(), None), "up")
    torch.cuda.synchronize()
    nat.stage_timing_enable(False)
    t = nat.stage_timing_read()["voc_up"]
    print(f"CI={CI} L={L} B={B} mode {mode} ({name}): {t[0] / t[1]:.3f} ms per kernel launch")
lib.m2tts_voc_up_h_set_debug(0)
