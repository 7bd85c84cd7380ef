"""Bring-up tool (GPU box): does a K-major swizzled UMMA A operand accept a descriptor start address that is
moved by whole rows inside the swizzle pattern?  usage: python tests/rowshift_probe_run.py all | <rowbytes> <K> <shift> <bo>"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))


def run(rowbytes, K, shift, bo, N=32, rows_total=144):
    import torch
    from models import _native as nat
    lib = nat.tools_lib()
    fn = lib.m2tts_rowshift_probe
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]
    g = torch.Generator().manual_seed(1)
    A = (torch.randint(-8, 9, (rows_total, K), generator=g).float() / 4).cuda()
    Bm = (torch.randint(-8, 9, (N, K), generator=g).float() / 4).cuda()
    D = torch.full((128, N), -777.0, device="cuda")
    rc = fn(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), rows_total, N, K, rowbytes, shift, bo, None)
    torch.cuda.synchronize()
    want = A[shift:shift + 128] @ Bm.T
    err = float((D - want).abs().max())
    # which shift does the result actually correspond to?
    best = min(range(0, rows_total - 127), key=lambda s: float((D - A[s:s + 128] @ Bm.T).abs().max()))
    print(f"rowbytes={rowbytes} K={K} shift={shift} base_offset={bo} rc={rc} max|err|={err:.4g} best_matching_shift={best} "
          f"(err {float((D - A[best:best + 128] @ Bm.T).abs().max()):.3g})")


if __name__ == "__main__":
    if sys.argv[1] == "all":
        for rowbytes, K in ((128, 32), (128, 64), (64, 16)):
            for shift in (0, 1, 2, 3, 7, 8, 9):
                for bo in sorted({0, shift & 7, (shift >> 1) & 3}):
                    r = subprocess.run([sys.executable, __file__, str(rowbytes), str(K), str(shift), str(bo)],
                                       capture_output=True, text=True, timeout=120)
                    out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                    print(out if r.returncode == 0 else f"rowbytes={rowbytes} K={K} shift={shift} bo={bo} FAILED rc={r.returncode}: "
                          f"{(r.stderr.strip().splitlines() or [''])[-1][:200]}")
    else:
        run(*[int(x) for x in sys.argv[1:5]])
