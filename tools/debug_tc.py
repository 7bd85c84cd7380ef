"""Bring-up diagnostics for the tcgen05 attention kernel (run on the GPU box, not a pytest):
feeds hand-built hi/lo planes to m2tts_attention_tc_planes and localises errors in the first
score tile (S = Q K^T) and the first P*V tile."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
from models import _native as nat  # noqa: E402


def split(x):
    hi = (x.view(torch.int32) & -8192).view(torch.float32)
    lo = ((x - hi).view(torch.int32) & -8192).view(torch.float32)
    return hi, lo


def main(hd=48, L=333, B=2, nh=2):
    dev = "cuda:0"
    lib = nat.lib()
    lib.m2tts_attention_tc_planes.restype = C.c_int
    lib.m2tts_attention_tc_planes.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p] * 3
    g = torch.Generator().manual_seed(0)
    Lp = (L + 3) // 4 * 4
    q = torch.randn(B, nh, hd, L, generator=g) * 0.5
    k = torch.randn(B, nh, hd, L, generator=g)
    v = torch.randn(B, nh, hd, L, generator=g)
    planes = torch.zeros(6, B, nh, hd, Lp)
    for i, x in enumerate((q, k, v)):
        hi, lo = split(x)
        planes[2 * i, ..., :L] = hi
        planes[2 * i + 1, ..., :L] = lo
    pd = planes.to(dev)
    ctx = torch.zeros(B, L, nh * hd, device=dev)
    dbg_s = torch.zeros(128, 64, device=dev)
    dbg_o = torch.zeros(128, hd, device=dev)
    rc = lib.m2tts_attention_tc_planes(pd.data_ptr(), ctx.data_ptr(), None, B, L, Lp, nh, hd,
                                       dbg_s.data_ptr(), dbg_o.data_ptr(), None)
    torch.cuda.synchronize()
    print("rc", rc, lib.m2tts_last_error_string())
    q0, k0, v0 = q[0, 0].double(), k[0, 0].double(), v[0, 0].double()     # [hd, L]
    S = (q0[:, :128].T @ k0[:, :64])                                      # [128, 64]
    got = dbg_s.cpu().double()
    print("S tile: max|err| =", float((got - S).abs().max()), " max|S| =", float(S.abs().max()))
    if (got - S).abs().max() > 1e-3:
        # which d-groups (k-steps) contributed? least squares over per-group partial products
        parts = torch.stack([(q0[8 * i:8 * i + 8, :128].T @ k0[8 * i:8 * i + 8, :64]).flatten() for i in range(hd // 8)], 1)
        coef = torch.linalg.lstsq(parts, got.flatten()[:, None]).solution.flatten()
        print("  per-k-step coefficients (want all 1):", [round(float(c), 3) for c in coef])
        for rb in range(4):
            for cb in range(2):
                e = (got[32 * rb:32 * rb + 32, 32 * cb:32 * cb + 32] - S[32 * rb:32 * rb + 32, 32 * cb:32 * cb + 32]).abs().max()
                print(f"  block rows {32*rb}.. cols {32*cb}..: max err {float(e):.3e}")
        hi_only = (split(q[0, 0])[0].double()[:, :128].T @ split(k[0, 0])[0].double()[:, :64])
        print("  err vs hi*hi only:", float((got - hi_only).abs().max()))
        print("  got[0,:8]", got[0, :8].tolist())
        print("  exp[0,:8]", S[0, :8].tolist())
        print("  got[:8,0]", got[:8, 0].tolist())
        print("  exp[:8,0]", S[:8, 0].tolist())
    # first PV tile from the kernel's own S (so an S bug does not mask a PV bug)
    P = torch.exp2(got - got.max(dim=1, keepdim=True).values)
    O = P @ v0[:, :64].T                                                  # [128, hd]
    goto = dbg_o.cpu().double()
    print("O tile: max|err| =", float((goto - O).abs().max()), " max|O| =", float(O.abs().max()))
    if (goto - O).abs().max() > 1e-3:
        parts = torch.stack([(P[:, 8 * i:8 * i + 8] @ v0[:, 8 * i:8 * i + 8].T).flatten() for i in range(8)], 1)
        coef = torch.linalg.lstsq(parts, goto.flatten()[:, None]).solution.flatten()
        print("  per-k-step (8 keys) coefficients (want all 1):", [round(float(c), 3) for c in coef])
        for cb in range(hd // 8):
            e = (goto[:, 8 * cb:8 * cb + 8] - O[:, 8 * cb:8 * cb + 8]).abs().max()
            print(f"  d cols {8*cb}..: max err {float(e):.3e}")
        print("  got[0,:8]", goto[0, :8].tolist())
        print("  exp[0,:8]", O[0, :8].tolist())
    # full result
    s = torch.einsum("bhdq,bhdk->bhqk", q.double(), k.double())
    p = torch.softmax(s * 0.6931471805599453, dim=-1)   # kernel works in log2 domain: 2^s
    want = torch.einsum("bhqk,bhdk->bqhd", p, v.double()).reshape(B, L, nh * hd)
    print("ctx: max|err| =", float((ctx.cpu().double() - want).abs().max()))


if __name__ == "__main__":
    for hd in (48, 32):
        print("==== hd", hd)
        main(hd=hd)
