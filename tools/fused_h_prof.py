"""Bring-up helper (GPU box): phase timestamps of the 16-bit split fused narrow vocoder stages (voc_fused_h.cu), CTA 0, context 0,
first epilogue warp, iterations 8..47 of a C3 vocoder pass (the prof buffer is written by both fused launches: the C = 16 stage,
which runs last, is what remains; pass `32` to stop before it)."""
import _env  # noqa: F401  (selects libm2tts_b200_tools.so)
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "m2-tts_b200" / "src"))
import torch
from models import _native as nat
lib = nat.tools_lib()
names = ["wait up", "EPI1", "wait conv1", "EPI2", "wait conv2", "EPI3", "sync + tail"]
for Cc, final in ((32, False), (16, True)):
    B, L = 64, (55136 if Cc == 32 else 110272)
    x = torch.randn(B, L, 2 * Cc, device="cuda")
    up_w = torch.randn(2 * Cc, Cc, 4, device="cuda") * 0.1
    w1 = torch.randn(Cc, Cc, 3, device="cuda") * 0.1
    w2 = torch.randn(Cc, Cc, 3, device="cuda") * 0.1
    b = torch.zeros(Cc, device="cuda")
    ow = torch.randn(1, Cc, 3, device="cuda") * 0.1
    ob = torch.zeros(1, device="cuda")
    y = torch.empty((B, 2 * L) if final else (B, 2 * L, Cc), device="cuda")
    ws = torch.empty(lib.m2tts_vocoder_stage_fused_h_workspace_bytes(B, Cc, L), dtype=torch.uint8, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    prof = torch.zeros(2 * 64 * 8, dtype=torch.int64, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(2):
        if rep == 1:
            lib.m2tts_attention_set_prof(prof.data_ptr())
            e0.record()
        rc = lib.m2tts_vocoder_stage_fused_h(x.data_ptr(), up_w.data_ptr(), b.data_ptr(), w1.data_ptr(), b.data_ptr(), w2.data_ptr(), b.data_ptr(),
                                             ow.data_ptr() if final else None, ob.data_ptr() if final else None, y.data_ptr(), B, Cc, L,
                                             st.data_ptr(), ws.data_ptr(), ws.numel(), None)
        nat.check(rc, "fused_h")
    e1.record()
    torch.cuda.synchronize()
    lib.m2tts_attention_set_prof(None)
    p2 = prof.cpu()[512:].view(64, 8)[8:48]
    p = prof.cpu()[:512].view(64, 8)
    d = p[8:48]
    segs = [(d[:, i + 1] - d[:, i]).float().mean().item() for i in range(6)] + [(p[9:49, 0] - p[8:48, 6]).float().mean().item()]
    print(f"C={Cc}: per iteration (cycles) " + ", ".join(f"{n}={v:.0f}" for n, v in zip(names, segs)) +
          f"  period={(p[48, 0] - p[8, 0]).item() / 40:.0f}; split + stage {e0.elapsed_time(e1):.3f} ms")
    d1 = [(p2[:, i + 1] - p2[:, i]).float().mean().item() for i in range(4)]
    print(f"      EPI1 in detail: TMEM loads + sums={d1[0]:.0f}, lrelu + split + shared-memory stores={d1[1]:.0f}, tcgen05 fence={d1[2]:.0f}, proxy fence={d1[3]:.0f}")
