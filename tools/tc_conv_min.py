import sys, ctypes as C, torch
sys.path.insert(0, "m2-tts_b200/src")
from models import _native as nat
lib = nat.lib()
g = torch.Generator().manual_seed(31)
CI, CO, L = 128, 128, 300
x = torch.randn(2, CI, L, generator=g).cuda(); w = (torch.randn(CO, CI, 3, generator=g) * 0.05).cuda(); b = torch.randn(CO, generator=g).cuda()
y = torch.empty(2, CO, L, device="cuda")
ws = torch.empty(lib.m2tts_conv_tc_workspace_bytes(2, CI, CO, L, 1), dtype=torch.uint8, device="cuda")
rc = lib.m2tts_conv1d_k3_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), None, y.data_ptr(), 2, CI, CO, L, 1, 0, ws.data_ptr(), ws.numel(), None)
print("rc", rc)
try:
    torch.cuda.synchronize()
except Exception as e:
    print("sync failed:", str(e).splitlines()[0])
    d = (C.c_int * 8)(); lib.m2tts_debug_words(d, 8); print("debug words", list(d)); sys.exit(1)
want = torch.nn.functional.conv1d(x, w, b, padding=1)
print("err", float((y - want).abs().max()))
