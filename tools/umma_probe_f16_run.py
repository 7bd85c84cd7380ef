"""Bring-up tool (GPU box): kind::f16 UMMA operand layouts. usage: python tests/umma_probe_f16_run.py all | <mode> <N> <K>"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))


def run(mode, N, K):
    import torch
    from models import _native as nat
    lib = nat.tools_lib()
    g = torch.Generator().manual_seed(1)
    A = (torch.randint(-8, 9, (128, K), generator=g).float() / 4).cuda()
    Bm = (torch.randint(-8, 9, (N, K), generator=g).float() / 4).cuda()
    D = torch.full((128, N), -777.0, device="cuda")
    rc = lib.m2tts_umma_probe_f16(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, K, mode, None)
    torch.cuda.synchronize()
    want = A @ Bm.T
    err = float((D - want).abs().max())
    swapped = A.clone()
    swapped[:, 0::2], swapped[:, 1::2] = A[:, 1::2], A[:, 0::2]
    print(f"mode={mode} N={N} K={K} rc={rc} max|err|={err:.4g} nonzero={float((D != 0).float().mean()):.2f} "
          f"err_if_k_pairs_swapped={float((D - swapped @ Bm.T).abs().max()):.3g} err_first_kstep_only={float((D - A[:, :16] @ Bm[:, :16].T).abs().max()):.3g}")


if __name__ == "__main__":
    if sys.argv[1] == "all":
        for mode in (0, 1, 2):
            for N, K in ((64, 32), (128, 48), (96, 64), (48, 64)):
                r = subprocess.run([sys.executable, __file__, str(mode), str(N), str(K)], capture_output=True, text=True, timeout=120)
                out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
                print(out if r.returncode == 0 else f"mode={mode} N={N} K={K} FAILED rc={r.returncode}: {(r.stderr.strip().splitlines() or [''])[-1][:200]}")
    else:
        run(*[int(x) for x in sys.argv[1:4]])
