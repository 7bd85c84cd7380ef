"""Run one tcgen05 operand-layout configuration of m2tts_umma_probe (GPU box bring-up tool).
usage: python tests/umma_probe_run.py <config-name>  |  all  (spawns one process per config)"""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import ctypes as C
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))

N, K = 64, 16


def kmaj_sw(rows):   # mode 0
    return dict(mode=0, lbo=16, sbo=1024, kstep=32, flag=0, bytes=(rows // 8) * 1024)


def mnmaj_sw(rows):  # mode 1: k groups adjacent (SBO 1024), mn atoms after them (LBO = K/8 * 1024)
    return dict(mode=1, lbo=(K // 8) * 1024, sbo=1024, kstep=1024, flag=1, bytes=(rows // 32) * (K // 8) * 1024)


def kmaj_ns(rows):   # mode 2
    return dict(mode=2, lbo=128, sbo=(K // 4) * 128, kstep=256, flag=0, bytes=(rows // 8) * (K // 4) * 128)


def mnmaj_ns(rows):  # mode 3
    return dict(mode=3, lbo=(rows // 4) * 128, sbo=128, kstep=(rows // 4) * 128, flag=1, bytes=(K // 8) * (rows // 4) * 128)


def mnmaj_sw32(rows):  # mode 4: 4-row k groups adjacent (SBO 512), mn atoms after them (LBO = K/4 * 512)
    return dict(mode=4, lbo=(K // 4) * 512, sbo=512, kstep=1024, flag=1, bytes=(rows // 32) * (K // 4) * 512)


def swapped(d):
    e = dict(d); e["dlbo"], e["dsbo"] = d["sbo"], d["lbo"]; return e


CONFIGS = {
    "mm_sw32": (mnmaj_sw32(128), mnmaj_sw32(N)),
    "mk_sw32": (mnmaj_sw32(128), kmaj_sw(N)),
    "km_sw32": (kmaj_sw(128), mnmaj_sw32(N)),
    "mm_sw32_swapped": (swapped(mnmaj_sw32(128)), swapped(mnmaj_sw32(N))),
    "kk_sw": (kmaj_sw(128), kmaj_sw(N)),
    "mm_sw": (mnmaj_sw(128), mnmaj_sw(N)),
    "mk_sw": (mnmaj_sw(128), kmaj_sw(N)),
    "km_sw": (kmaj_sw(128), mnmaj_sw(N)),
    "kk_ns": (kmaj_ns(128), kmaj_ns(N)),
    "mm_ns": (mnmaj_ns(128), mnmaj_ns(N)),
    "mm_sw_swapped": (swapped(mnmaj_sw(128)), swapped(mnmaj_sw(N))),
    "mm_ns_swapped": (swapped(mnmaj_ns(128)), swapped(mnmaj_ns(N))),
    "kk_ns_swapped": (swapped(kmaj_ns(128)), swapped(kmaj_ns(N))),
}


def arr(d):
    v = [d["mode"], d["lbo"], d["sbo"], d["kstep"], d["flag"], d["bytes"], d.get("dlbo", d["lbo"]), d.get("dsbo", d["sbo"])]
    return (C.c_int * 8)(*v)


def run(name):
    import torch
    from models import _native as nat
    lib = nat.tools_lib()
    lib.m2tts_umma_probe.restype = C.c_int
    lib.m2tts_umma_probe.argtypes = [C.c_void_p] * 3 + [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
    g = torch.Generator().manual_seed(1)
    # tf32-exact inputs so the expected product is exact in fp32
    A = (torch.randint(-8, 9, (128, K), generator=g).float() / 4).cuda()
    Bm = (torch.randint(-8, 9, (N, K), generator=g).float() / 4).cuda()
    D = torch.full((128, N), -777.0, device="cuda")
    a, b = CONFIGS[name]
    rc = lib.m2tts_umma_probe(A.data_ptr(), Bm.data_ptr(), D.data_ptr(), N, K, arr(a), arr(b), None)
    torch.cuda.synchronize()
    want = A @ Bm.T
    err = float((D - want).abs().max())
    nz = float((D != 0).float().mean())
    # partial-structure hints
    e_k0 = float((D - A[:, :8] @ Bm[:, :8].T).abs().max())
    print(f"{name:16s} rc={rc} max|err|={err:.4g} nonzero={nz:.2f} err_vs_first_kstep_only={e_k0:.4g} "
          f"rows0-31 err={float((D[:32]-want[:32]).abs().max()):.3g} cols0-31 err={float((D[:, :32]-want[:, :32]).abs().max()):.3g}")


if __name__ == "__main__":
    if sys.argv[1] == "all":
        for n in CONFIGS:
            r = subprocess.run([sys.executable, __file__, n], capture_output=True, text=True, timeout=120)
            out = (r.stdout.strip().splitlines() or ["<no output>"])[-1]
            print(out if r.returncode == 0 else f"{n:16s} FAILED rc={r.returncode}: {(r.stderr.strip().splitlines() or [''])[-1][:200]}")
    else:
        run(sys.argv[1])
