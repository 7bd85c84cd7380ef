"""Bring-up tool (GPU box): cycles per tcgen05.mma kind::tf32 (M=128, K=8) for the operand configurations the kernels use."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
out = torch.zeros(2, dtype=torch.int64, device="cuda")
names = {0: "A smem K-major SW128 / B K-major noswz", 1: "A,B smem MN-major SW128-32B", 2: "A TMEM / B K-major SW128",
         3: "A smem K-major SW64 / B K-major noswz", 4: "A,B smem K-major SW128", 5: "kind::f16, A,B smem K-major SW128",
         6: "kind::f16, A row-shifted, SW128"}
n = 512
for mode in ((int(sys.argv[1]),) if len(sys.argv) > 1 else (2, 4, 5, 6)):
    for N in (16, 32, 48, 64, 96, 128, 192, 256):
        for nacc, elect in ((1, 2),):
            if nacc * N > 256:
                continue
            for _ in range(2):
                nat.check(lib.m2tts_mma_bench(mode, N, n, nacc, elect, out.data_ptr(), None), "mma_bench")
                torch.cuda.synchronize()
            issue, total = (int(v) for v in out.cpu())
            print(f"mode {mode} ({names[mode]:40s}) N={N:3d} nacc={nacc} elect={elect}: issue {issue / n:6.1f} cyc/MMA, total {total / n:6.1f} cyc/MMA "
                  f"(math floor {N / 2:.0f})")
