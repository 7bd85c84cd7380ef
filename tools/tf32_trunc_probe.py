"""Bring-up tool (GPU box): does tcgen05.mma kind::tf32 TRUNCATE fp32 operands (ignore the low 13 mantissa bits) or round
them? D = A @ B^T with B = identity-like ones and A holding values whose truncation and rounding differ."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
K, N, rows = 32, 32, 144
A = torch.zeros(rows, K)
vals = [1 + 2 ** -10 + 2 ** -11 + 2 ** -12, 1 + 2 ** -11, 1 + 2 ** -11 + 2 ** -20, -(1 + 2 ** -10 + 2 ** -11 + 2 ** -13), 3.1415926, 1e-3 * 1.2345678]
for i, v in enumerate(vals):
    A[i, 0] = v
Bm = torch.zeros(N, K); Bm[:, 0] = 1.0
D = torch.full((128, N), -7.0, device="cuda")
Ad, Bd = A.cuda(), Bm.cuda()
nat.check(lib.m2tts_rowshift_probe(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), rows, N, K, 128, 0, 0, None), "probe")
torch.cuda.synchronize()
for i, v in enumerate(vals):
    x = torch.tensor([v], dtype=torch.float32)
    trunc = (x.view(torch.int32) & -8192).view(torch.float32).item()
    # round to nearest even on 13 dropped bits
    xi = x.view(torch.int32).item()
    rn = torch.tensor([(xi + 0xFFF + ((xi >> 13) & 1)) & ~0x1FFF], dtype=torch.int32).view(torch.float32).item()
    got = D[i, 0].item()
    print(f"x={x.item():.10f} trunc={trunc:.10f} rn={rn:.10f} mma={got:.10f} -> {'TRUNC' if got == trunc else ('ROUND' if got == rn else '??')}")
