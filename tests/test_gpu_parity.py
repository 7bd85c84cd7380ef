"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(ctypes -> libm2tts_b200.so), against the CPU oracle on the same seeded inputs and against the
committed golden fixtures.

Tolerances (BASELINE.json north_star): length-regulator indices / frame counts / regulated rows
bit-exact; fp32 mels and waveforms within max-abs 1e-4.
"""
import ctypes as C

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
DEV = "cuda:0"


class _Status:
    """Device status word for the stand-alone C-ABI entry points (include/m2tts_b200.h, "Status word"); every test that
    passes it also checks that it stayed clear."""

    def __init__(self):
        self._t = None

    def data_ptr(self):
        if self._t is None:
            self._t = torch.zeros(1, dtype=torch.int32, device=DEV)
        return self._t.data_ptr()

    def take(self):
        v = int(self._t.item()) if self._t is not None else 0
        if self._t is not None:
            self._t.zero_()
        return v


STATUS = _Status()


@pytest.fixture(autouse=True)
def _status_word_stays_clear():
    yield
    assert STATUS.take() == 0, "a stand-alone kernel raised a status bit"


def cuda_model(stage, perturb=None, **override):
    return H.product_model(stage, perturb=perturb, **override).to(DEV).eval()


def cpu_sd(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


def load(name):
    return {k: v for k, v in np.load(H.GOLDEN / name).items()}


def test_native_library_is_what_runs():
    from models import _native as nat
    assert nat.library_path().exists()
    before = nat.launch_count()
    m = cuda_model("tiny")
    ids, lengths, dur = H.small_inputs(2, 10, 256, seed=11)
    m(ids.to(DEV), lengths.to(DEV), dur.to(DEV), 50)
    torch.cuda.synchronize()
    assert nat.launch_count() - before >= 20  # every stage is one of our kernels


# --------------------------------------------------------------------------- full forward
@pytest.mark.parametrize("name,stage,pert,ml", [("tiny_fwd.npz", "tiny", None, 50),
                                               ("tiny_biased_fwd.npz", "tiny", 7, 50),
                                               ("stage2_biased_fwd.npz", "stage2", 7, 60)])
def test_forward_matches_golden_and_oracle(name, stage, pert, ml):
    g = load(name)
    m = cuda_model(stage, pert)
    ids, lengths, dur = (torch.from_numpy(g[k]) for k in ("ids", "lengths", "dur"))
    out = m(ids.to(DEV), lengths.to(DEV), dur.to(DEV), ml)
    ref = oracle.forward(cpu_sd(m), ids, lengths, dur, ml)
    assert torch.equal(out["padding_mask"].cpu(), torch.from_numpy(g["padding_mask"]))
    for k in ("encoder_output", "duration_pred", "mel_output", "audio_output"):
        assert out[k].shape == g[k].shape, k
        assert H.max_abs(out[k].cpu(), g[k]) <= FP32_TOL, f"{k} vs golden"
        assert H.max_abs(out[k].cpu(), ref[k]) <= FP32_TOL, f"{k} vs oracle"
    # regulated rows are copies of the (GPU) encoder rows: bit-exact gather of the GPU encoder output
    want = oracle.length_regulator(out["encoder_output"].cpu(), dur, ml)
    assert torch.equal(out["regulated_output"].cpu(), want)


def test_c1_hello_world_inference():
    g = load("c1_hello_world.npz")
    m = cuda_model("stage1")
    ids, lengths = torch.from_numpy(g["ids"]).to(DEV), torch.from_numpy(g["lengths"]).to(DEV)
    mel, audio = m.inference(ids, lengths, 1.0)
    assert mel.shape == (1, 1, 64) and audio.shape == (1, 1, 64)  # all durations truncate to 0
    assert H.max_abs(mel.cpu(), g["mel_scale1"]) <= FP32_TOL
    assert H.max_abs(audio.cpu(), g["audio_scale1"]) <= FP32_TOL
    fwd = m(ids, lengths)
    assert H.max_abs(fwd["duration_pred"].cpu(), g["duration_pred"]) <= FP32_TOL
    # non-degenerate case: feed the golden durations x4 to the regulator on both sides
    dur4 = torch.from_numpy(g["duration_pred"]).to(DEV) * 4.0
    reg = m.length_regulator(fwd["encoder_output"], dur4)
    mel4 = m.decoder(reg)
    audio4 = m.vocoder(mel4.transpose(1, 2))
    assert mel4.shape == g["mel_scale4"].shape and audio4.shape == g["audio_scale4"].shape
    assert H.max_abs(mel4.cpu(), g["mel_scale4"]) <= FP32_TOL
    assert H.max_abs(audio4.cpu(), g["audio_scale4"]) <= FP32_TOL


def test_c2_stage1_batch16():
    g = load("c2_stage1.npz")
    m = cuda_model("stage1")
    ids, lengths, dur = H.c2_inputs()
    out = m(ids.to(DEV), lengths.to(DEV), target_durations=dur.to(DEV))
    T = int(g["T"])
    assert out["mel_output"].shape == (16, T, 64) and out["audio_output"].shape == (16, 1, 64 * T)
    assert np.array_equal(m.length_regulator.last_frames.cpu().numpy(), g["frames"])
    index, frames, _ = oracle.length_regulator_indices(dur.numpy())
    assert np.array_equal(m.length_regulator.last_index.cpu().numpy(), index)
    keep = g["keep"].tolist()
    assert H.max_abs(out["encoder_output"].cpu(), g["encoder_output"]) <= FP32_TOL
    assert H.max_abs(out["mel_output"][keep].cpu(), g["mel_keep"]) <= FP32_TOL
    assert H.max_abs(out["audio_output"][keep].cpu(), g["audio_keep"]) <= FP32_TOL
    assert np.allclose(out["audio_output"].double().abs().sum((1, 2)).cpu().numpy(), g["audio_abs"], rtol=1e-5)
    ref = oracle.forward(cpu_sd(m), ids, lengths, dur)
    assert H.max_abs(out["mel_output"].cpu(), ref["mel_output"]) <= FP32_TOL
    assert H.max_abs(out["audio_output"].cpu(), ref["audio_output"]) <= FP32_TOL


# --------------------------------------------------------------------------- length regulator
def test_length_regulator_edges_bit_exact():
    from models.tts_model import LengthRegulator
    g = load("length_regulator_edges.npz")
    d = torch.from_numpy(g["dur"])
    B, S = d.shape
    enc = torch.randn(B, S, 12, generator=torch.Generator().manual_seed(3))
    lr = LengthRegulator().eval()
    for name, ml in (("none", None), ("trunc20", 20), ("pad90", 90)):
        out = lr(enc.to(DEV), d.to(DEV), ml)
        assert np.array_equal(lr.last_index.cpu().numpy(), g[f"index_{name}"]), name
        assert torch.equal(out.cpu(), oracle.length_regulator(enc, d, ml)), name
    index, frames, _ = oracle.length_regulator_indices(d.numpy())
    assert np.array_equal(lr.last_frames.cpu().numpy(), frames)


def test_length_regulator_large_and_ragged():
    from models.tts_model import LengthRegulator
    g = torch.Generator().manual_seed(21)
    B, S, Hd = 64, 256, 96
    d = torch.randint(0, 28, (B, S), generator=g).float() + torch.rand((B, S), generator=g) * 0.999
    d[5] = 0.3                                  # empty utterance inside a big batch
    enc = torch.randn(B, S, Hd, generator=g)
    lr = LengthRegulator().eval()
    out = lr(enc.to(DEV), d.to(DEV))
    index, frames, T = oracle.length_regulator_indices(d.numpy())
    assert out.shape == (B, T, Hd)
    assert np.array_equal(lr.last_frames.cpu().numpy(), frames)
    assert np.array_equal(lr.last_index.cpu().numpy(), index)
    assert torch.equal(out.cpu(), oracle.length_regulator(enc, d))
    # H not a multiple of 4 (scalar copy path), S == 1
    enc2 = torch.randn(3, 1, 7, generator=g)
    d2 = torch.tensor([[3.9], [0.0], [1.0]])
    assert torch.equal(lr(enc2.to(DEV), d2.to(DEV)).cpu(), oracle.length_regulator(enc2, d2))


def test_length_regulator_nan_inf_raise_like_python_int():
    from models.tts_model import LengthRegulator
    lr = LengthRegulator().eval()
    enc = torch.zeros(1, 4, 8, device=DEV)
    d = torch.ones(1, 4, device=DEV)
    d[0, 2] = float("nan")
    with pytest.raises(ValueError):
        lr(enc, d)
    d[0, 2] = float("-inf")
    with pytest.raises(OverflowError):
        lr(enc, d)


# --------------------------------------------------------------------------- encoder / attention
@pytest.mark.parametrize("stage", ["stage1", "stage2"])
def test_text_encoder_masks_and_edge_lengths(stage):
    m = cuda_model(stage, perturb=5)
    sd = cpu_sd(m)
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(0, 256, (5, 150), generator=g)        # > one 128-query tile, ragged key tiles
    for lengths in (torch.tensor([150, 1, 0, 64, 129]), None):
        got, mask = m.text_encoder(ids.to(DEV), None if lengths is None else lengths.to(DEV))
        want, wmask = oracle.text_encoder(sd, ids, lengths, 2)
        assert H.max_abs(got.cpu(), want) <= FP32_TOL
        if lengths is not None:
            assert torch.equal(mask.cpu(), wmask)


@pytest.mark.parametrize("heads,hidden", [(2, 32), (4, 64), (2, 128), (8, 64)])
def test_decoder_head_dims(heads, hidden):
    m = cuda_model("stage1", perturb=2, hidden_dim=hidden, num_heads=heads, mel_channels=20,
                   text_encoder_layers=1, decoder_layers=2)
    x = torch.randn(2, 203, hidden, generator=torch.Generator().manual_seed(1))
    got = m.decoder(x.to(DEV))
    want = oracle.mel_decoder(cpu_sd(m), x, heads)
    assert H.max_abs(got.cpu(), want) <= FP32_TOL


def test_decoder_long_sequence_stage2():
    m = cuda_model("stage2", perturb=4)
    x = torch.randn(1, 1500, 96, generator=torch.Generator().manual_seed(2))
    got = m.decoder(x.to(DEV))
    want = oracle.mel_decoder(cpu_sd(m), x, 2)
    assert H.max_abs(got.cpu(), want) <= FP32_TOL


@pytest.mark.parametrize("stage,B,S", [("stage2", 3, 77), ("stage2", 8, 256), ("stage1", 5, 64), ("stage1", 2, 1), ("tiny", 2, 33)])
def test_duration_predictor_nonzero_bias_padding(stage, B, S):
    """Non-zero biases / BN statistics catch the per-layer zero padding; S = 1, S % 32 != 0 and S = 8 x 32 cover the tile edges,
    stage2 (H = 96) the chunked weight staging (two output-channel chunks), stage1 the whole-layer one."""
    m = cuda_model(stage, perturb=8)
    Hd = H.STAGE_KWARGS[stage]["hidden_dim"]
    enc = torch.randn(B, S, Hd, generator=torch.Generator().manual_seed(4))
    got = m.duration_predictor(enc.to(DEV))
    want = oracle.duration_predictor(cpu_sd(m), enc)
    assert got.shape == (B, S)
    assert H.max_abs(got.cpu(), want) <= FP32_TOL


# --------------------------------------------------------------------------- vocoder
@pytest.mark.parametrize("stage,B,T", [("stage1", 3, 37), ("stage2", 2, 130), ("tiny", 1, 1), ("stage2", 1, 515)])
def test_vocoder_matches_oracle(stage, B, T):
    m = cuda_model(stage, perturb=6)
    M = H.STAGE_KWARGS[stage]["mel_channels"]
    mel = torch.randn(B, M, T, generator=torch.Generator().manual_seed(T))
    want = oracle.vocoder(cpu_sd(m), mel)
    got = m.vocoder(mel.to(DEV))                                   # contiguous [B,M,T]
    assert got.shape == (B, 1, 64 * T)
    assert H.max_abs(got.cpu(), want) <= FP32_TOL
    got_t = m.vocoder(mel.transpose(1, 2).contiguous().to(DEV).transpose(1, 2))  # strided view of [B,T,M]
    assert H.max_abs(got_t.cpu(), want) <= FP32_TOL


@pytest.mark.parametrize("B,T", [(1, 128), (2, 256), (4, 128), (8, 131), (1, 2048)])
def test_c5_vocoder_sweep_points_match_oracle(B, T):
    """BASELINE.json configs[4] (vocoder-only sweep): points the CPU oracle finishes in seconds; the full grid is
    timed by tools/sweep_vocoder.py and covered at large sizes by the batch-independence property below."""
    m = cuda_model("stage2")
    mel = torch.randn(B, 80, T, generator=torch.Generator().manual_seed(B * 10000 + T))
    got = m.vocoder(mel.to(DEV))
    assert got.shape == (B, 1, 64 * T)
    assert H.max_abs(got.cpu(), oracle.vocoder(cpu_sd(m), mel)) <= FP32_TOL


def test_vocoder_large_batch_rows_are_independent_and_deterministic():
    """C5 at sizes the oracle cannot reach: utterance b of a 64 x 1024 batch equals the same mel run alone (bit-exact),
    twice in a row, and a batch of one utterance repeated gives identical rows."""
    m = cuda_model("stage2")
    mel = torch.randn(64, 80, 1024, generator=torch.Generator().manual_seed(77)).to(DEV)
    full = m.vocoder(mel)
    assert torch.equal(full, m.vocoder(mel))
    for b in (0, 31, 63):
        assert torch.equal(full[b:b + 1], m.vocoder(mel[b:b + 1]))
    rep = m.vocoder(mel[5:6].expand(16, -1, -1).contiguous())
    assert all(torch.equal(rep[0], rep[i]) for i in range(1, 16))


def test_vocoder_batch_independence():
    """Size-independent property at a larger size: the vocoder is deterministic and
    batch-independent — utterance b of a batch equals the same utterance run alone."""
    m = cuda_model("stage2")
    mel = torch.randn(4, 80, 700, generator=torch.Generator().manual_seed(5)).to(DEV)
    full = m.vocoder(mel)
    for b in (0, 3):
        assert torch.equal(full[b:b + 1], m.vocoder(mel[b:b + 1]))


def test_conv_entry_points_dilation_and_activation():
    """Per-stage C-ABI entry points, incl. the generic-dilation path no config exercises."""
    from models import _native as nat
    lib = nat.lib()
    g = torch.Generator().manual_seed(12)
    for (CI, CO, L, dil, act) in [(24, 40, 301, 1, 1), (16, 16, 257, 3, 0), (8, 4, 64, 2, 2), (5, 70, 33, 1, 0)]:
        x = torch.randn(2, CI, L, generator=g)
        w = torch.randn(CO, CI, 3, generator=g) * 0.2
        b = torch.randn(CO, generator=g)
        r = torch.randn(2, CO, L, generator=g)
        want = torch.nn.functional.conv1d(x, w, b, padding=dil, dilation=dil)
        want = {0: want, 1: torch.nn.functional.leaky_relu(want, 0.1), 2: torch.tanh(want)}[act] + r
        xd, wd, bd, rd = (t.to(DEV) for t in (x, w, b, r))
        y = torch.empty(2, CO, L, device=DEV)
        ws = torch.empty(lib.m2tts_conv_workspace_bytes(CI, CO, 3), dtype=torch.uint8, device=DEV)
        rc = lib.m2tts_conv1d_k3(xd.data_ptr(), CI * L, L, 1, wd.data_ptr(), bd.data_ptr(), rd.data_ptr(),
                                 y.data_ptr(), 2, CI, CO, L, dil, act, ws.data_ptr(), ws.numel(), None)
        nat.check(rc, "conv1d_k3")
        assert H.max_abs(y.cpu(), want) <= FP32_TOL, (CI, CO, L, dil, act)
    for (CI, CO, L, r_) in [(32, 16, 45, 4), (12, 6, 130, 2), (256, 128, 9, 4), (6, 3, 7, 2)]:
        x = torch.randn(2, CI, L, generator=g)
        w = torch.randn(CI, CO, 2 * r_, generator=g) * 0.2
        b = torch.randn(CO, generator=g)
        want = torch.nn.functional.leaky_relu(
            torch.nn.functional.conv_transpose1d(x, w, b, stride=r_, padding=r_ // 2), 0.1)
        y = torch.empty(2, CO, r_ * L, device=DEV)
        xd, wd, bd = (t.to(DEV) for t in (x, w, b))   # keep the device copies alive across the call
        rc = lib.m2tts_conv_transpose1d_lrelu(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(),
                                              y.data_ptr(), 2, CI, CO, L, r_, None)
        nat.check(rc, "conv_transpose1d")
        assert H.max_abs(y.cpu(), want) <= FP32_TOL, (CI, CO, L, r_)


# --------------------------------------------------------------------------- errors / boundary behaviour
def test_error_paths():
    from models import _native as nat
    m = H.product_model("tiny")  # on CPU
    with pytest.raises(nat.NativeLibraryError):
        m(torch.zeros(1, 4, dtype=torch.long))   # no CPU fallback in eval mode
    lib = nat.lib()
    assert lib.m2tts_vocoder_forward(None, None, None, 0, 0, 0, None, 1, 1, 1, 16, -1, None, None, 0, None) == -5
    assert b"null" in lib.m2tts_last_error_string()
    with pytest.raises(ValueError):   # head_dim 12 is not supported by the attention kernel
        bad = H.product_model("tiny", hidden_dim=24, num_heads=2).to(DEV).eval()
        bad.decoder(torch.zeros(1, 8, 24, device=DEV))


def test_batch_sharding_is_exact():
    """with a shared max_target_length a shard's results equal the full batch's."""
    m = cuda_model("stage1", perturb=1)
    ids, lengths, dur = H.c2_inputs()
    ids, lengths, dur = ids.to(DEV), lengths.to(DEV), dur.to(DEV)
    full = m(ids, lengths, dur, 330)
    half = m(ids[8:], lengths[8:], dur[8:], 330)
    for k in ("regulated_output", "mel_output", "audio_output"):
        assert torch.equal(full[k][8:], half[k]), k


# --------------------------------------------------------------------------- tensor-core attention
@pytest.mark.parametrize("prec", ["split16", "ffma", "tf32"])  # 16-bit split tcgen05 kernels (default), fp32 FFMA, TF32 split
def test_transformer_precisions_all_meet_fp32_tolerance(prec):
    from models import _native as nat
    with nat.precision(prec):
        for stage, heads in (("stage2", 2), ("stage1", 2)):
            m = cuda_model(stage, perturb=4)
            sd = cpu_sd(m)
            Hd = H.STAGE_KWARGS[stage]["hidden_dim"]
            x = torch.randn(2, 333, Hd, generator=torch.Generator().manual_seed(8))
            got = m.decoder(x.to(DEV))
            assert H.max_abs(got.cpu(), oracle.mel_decoder(sd, x, heads)) <= FP32_TOL, (stage, "decoder")
            ids = torch.randint(0, 256, (4, 200), generator=torch.Generator().manual_seed(9))
            lengths = torch.tensor([200, 77, 0, 129])
            enc, _ = m.text_encoder(ids.to(DEV), lengths.to(DEV))
            assert H.max_abs(enc.cpu(), oracle.text_encoder(sd, ids, lengths, heads)[0]) <= FP32_TOL, (stage, "encoder")
        # head_dim 64 (TF32: single-warpgroup kernel) and 16
        for heads, hidden in ((2, 128), (4, 64)):
            m = cuda_model("stage1", perturb=2, hidden_dim=hidden, num_heads=heads, mel_channels=20, text_encoder_layers=1, decoder_layers=1)
            x = torch.randn(2, 203, hidden, generator=torch.Generator().manual_seed(1))
            assert H.max_abs(m.decoder(x.to(DEV)).cpu(), oracle.mel_decoder(cpu_sd(m), x, heads)) <= FP32_TOL, (heads, hidden)


# --------------------------------------------------------------------------- tensor-core vocoder convolutions
def test_tensor_core_conv_entry_points():
    from models import _native as nat
    lib = nat.lib()
    g = torch.Generator().manual_seed(31)
    for (CI, CO, L, dil, act, with_res) in [(128, 128, 300, 1, 1, False), (64, 64, 517, 1, 0, True),
                                            (16, 64, 130, 2, 0, False), (256, 128, 64, 1, 1, True),
                                            (32, 32, 1000, 1, 1, True), (16, 16, 300, 1, 0, True),
                                            (32, 16, 200, 3, 0, False), (48, 48, 121, 4, 1, False)]:
        x = torch.randn(2, CI, L, generator=g)
        w = torch.randn(CO, CI, 3, generator=g) * (1.0 / (3 * CI) ** 0.5)
        b = torch.randn(CO, generator=g)
        r = torch.randn(2, CO, L, generator=g) if with_res else None
        want = torch.nn.functional.conv1d(x, w, b, padding=dil, dilation=dil)
        if act == 1:
            want = torch.nn.functional.leaky_relu(want, 0.1)
        if r is not None:
            want = want + r
        xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
        rd = r.to(DEV) if r is not None else None
        y = torch.empty(2, CO, L, device=DEV)
        ws = torch.empty(lib.m2tts_conv_tc_workspace_bytes(2, CI, CO, L, 1), dtype=torch.uint8, device=DEV)
        rc = lib.m2tts_conv1d_k3_tc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), nat.ptr(rd), y.data_ptr(),
                                    2, CI, CO, L, dil, act, ws.data_ptr(), ws.numel(), None)
        nat.check(rc, "conv1d_k3_tc")
        assert H.max_abs(y.cpu(), want) <= FP32_TOL, ("conv3_tc", CI, CO, L, dil, act)
    for (CI, CO, L, r_) in [(256, 128, 77, 4), (128, 64, 300, 4), (32, 32, 129, 4), (64, 32, 300, 2), (32, 16, 130, 2),
                            (16, 16, 50, 2)]:
        x = torch.randn(2, CI, L, generator=g)
        w = torch.randn(CI, CO, 2 * r_, generator=g) * (1.0 / (2 * CI) ** 0.5)
        b = torch.randn(CO, generator=g)
        want = torch.nn.functional.leaky_relu(
            torch.nn.functional.conv_transpose1d(x, w, b, stride=r_, padding=r_ // 2), 0.1)
        xd, wd, bd = x.to(DEV), w.to(DEV), b.to(DEV)
        y = torch.empty(2, CO, r_ * L, device=DEV)
        ws = torch.empty(lib.m2tts_conv_tc_workspace_bytes(2, CI, CO, L, r_), dtype=torch.uint8, device=DEV)
        rc = lib.m2tts_conv_transpose1d_lrelu_tc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), 2, CI, CO, L, r_,
                                                 ws.data_ptr(), ws.numel(), None)
        nat.check(rc, "conv_transpose1d_lrelu_tc")
        assert H.max_abs(y.cpu(), want) <= FP32_TOL, ("convT_tc", CI, CO, L, r_)


@pytest.mark.parametrize("prec", ["f16", "tf32"])
@pytest.mark.parametrize("C,L,B,final", [(16, 300, 2, False), (16, 300, 2, True), (32, 257, 2, False), (32, 130, 3, True),
                                         (16, 1, 1, True), (16, 2000, 1, True), (32, 1111, 2, False), (16, 61, 5, False)])
def test_fused_vocoder_stage(C, L, B, final, prec):
    """One (upsample x2, ResBlock[, output conv + tanh]) stage as a single channel-last tcgen05 kernel, in both split
    flavours (16-bit: fp16 hi/lo operands, the default of vocoder_forward; TF32)."""
    from models import _native as nat
    import torch.nn.functional as F
    lib = nat.lib()
    g = torch.Generator().manual_seed(100 * C + L)
    x = torch.randn(B, 2 * C, L, generator=g)
    up_w = torch.randn(2 * C, C, 4, generator=g) * (1.0 / (4 * C) ** 0.5)
    w1 = torch.randn(C, C, 3, generator=g) * (1.0 / (3 * C) ** 0.5)
    w2 = torch.randn(C, C, 3, generator=g) * (1.0 / (3 * C) ** 0.5)
    up_b, b1, b2 = (torch.randn(C, generator=g) * 0.3 for _ in range(3))
    ow = torch.randn(1, C, 3, generator=g) * 0.2
    ob = torch.randn(1, generator=g)
    u = F.leaky_relu(F.conv_transpose1d(x, up_w, up_b, stride=2, padding=1), 0.1)
    want = u + F.conv1d(F.leaky_relu(F.conv1d(u, w1, b1, padding=1), 0.1), w2, b2, padding=1)
    if final:
        want = torch.tanh(F.conv1d(want, ow, ob, padding=1))[:, 0]          # [B, 2L]
    else:
        want = want.transpose(1, 2).contiguous()                            # channel-last [B, 2L, C]
    d = [t.to(DEV) for t in (x.transpose(1, 2).contiguous(), up_w, up_b, w1, b1, w2, b2, ow, ob)]
    y = torch.full(want.shape, float("nan"), device=DEV)
    if prec == "f16":
        ws = torch.empty(lib.m2tts_vocoder_stage_fused_h_workspace_bytes(B, C, L), dtype=torch.uint8, device=DEV)
        fn = lib.m2tts_vocoder_stage_fused_h
    else:
        ws = torch.empty(lib.m2tts_vocoder_stage_fused_workspace_bytes(C), dtype=torch.uint8, device=DEV)
        fn = lib.m2tts_vocoder_stage_fused
    st = (STATUS.data_ptr(),) if prec == "f16" else ()
    rc = fn(*(t.data_ptr() for t in d[:7]), d[7].data_ptr() if final else None,
            d[8].data_ptr() if final else None, y.data_ptr(), B, C, L, *st, ws.data_ptr(), ws.numel(), None)
    nat.check(rc, "vocoder_stage_fused")
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "unwritten output rows"
    assert H.max_abs(y.cpu(), want) <= FP32_TOL, (C, L, B, final)


@pytest.mark.parametrize("C,L,B", [(64, 300, 2), (64, 126, 1), (64, 127, 3), (64, 1, 1), (64, 2000, 2),
                                   (32, 300, 2), (32, 126, 1), (32, 127, 3), (32, 1, 1), (32, 5000, 2)])
def test_fused_resblock_c64(C, L, B):
    """A whole ResBlock (C = 64: 128-byte operand rows; C = 32: 64-byte rows, the stage-1 model's second stage) as one
    channel-last 16-bit-split tcgen05 kernel."""
    from models import _native as nat
    import torch.nn.functional as F
    lib = nat.lib()
    g = torch.Generator().manual_seed(7 * L + B)
    x = torch.randn(B, C, L, generator=g)
    w1 = torch.randn(C, C, 3, generator=g) * (1.0 / (3 * C) ** 0.5)
    w2 = torch.randn(C, C, 3, generator=g) * (1.0 / (3 * C) ** 0.5)
    b1, b2 = (torch.randn(C, generator=g) * 0.3 for _ in range(2))
    want = (x + F.conv1d(F.leaky_relu(F.conv1d(x, w1, b1, padding=1), 0.1), w2, b2, padding=1)).transpose(1, 2).contiguous()
    d = [t.to(DEV) for t in (x.transpose(1, 2).contiguous(), w1, b1, w2, b2)]
    y = torch.full(want.shape, float("nan"), device=DEV)
    ws = torch.empty(lib.m2tts_resblock_fused_h_workspace_bytes(B, C, L), dtype=torch.uint8, device=DEV)
    nat.check(lib.m2tts_resblock_fused_h(*(t.data_ptr() for t in d), y.data_ptr(), B, C, L, STATUS.data_ptr(), ws.data_ptr(), ws.numel(), None),
              "resblock_fused_h")
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "unwritten output rows"
    assert H.max_abs(y.cpu(), want) <= FP32_TOL, (C, L, B)


@pytest.mark.parametrize("L,B,act,res,out_cl", [(1, 1, 0, False, 0), (127, 2, 1, False, 1), (128, 2, 0, True, 0), (129, 3, 1, True, 1),
                                                (1000, 5, 0, True, 0), (13784, 3, 1, False, 1), (13784, 3, 0, True, 0)])
def test_conv1d_k3_h_c128_matches_torch(L, B, act, res, out_cl):
    """One convolution of the widest ResBlock (C = 128) on channel-last 16-bit split operands (voc_conv_h.cu) against
    torch's conv1d (components.py:196-200): both activations, with and without the residual, both output layouts;
    lengths around the 128-row tile, a length that leaves CTAs idle and the C3 stage-0 length."""
    from models import _native as nat
    import torch.nn.functional as F
    lib = nat.lib()
    C = 128
    g = torch.Generator().manual_seed(11 * L + B)
    x = torch.randn(B, C, L, generator=g)
    r = torch.randn(B, C, L, generator=g)
    w = torch.randn(C, C, 3, generator=g) * (1.0 / (3 * C) ** 0.5)
    b = torch.randn(C, generator=g) * 0.3
    want = F.conv1d(x, w, b, padding=1)
    if act:
        want = F.leaky_relu(want, 0.1)
    if res:
        want = want + r
    if out_cl:
        want = want.transpose(1, 2).contiguous()
    xd, rd, wd, bd = (t.to(DEV) for t in (x.transpose(1, 2).contiguous(), r.transpose(1, 2).contiguous(), w, b))
    y = torch.full(want.shape, float("nan"), device=DEV)
    ws = torch.empty(lib.m2tts_conv1d_k3_h_workspace_bytes(B, C, L), dtype=torch.uint8, device=DEV)
    nat.check(lib.m2tts_conv1d_k3_h(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), rd.data_ptr() if res else None, y.data_ptr(), B, C, L,
                                    act, out_cl, STATUS.data_ptr(), ws.data_ptr(), ws.numel(), None), "conv1d_k3_h")
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "unwritten output rows"
    assert H.max_abs(y.cpu(), want) <= FP32_TOL, (L, B, act, res, out_cl)


@pytest.mark.parametrize("CI,L,B", [(256, 1, 1), (256, 127, 2), (256, 128, 1), (256, 129, 3), (256, 3446, 2), (128, 1, 2), (128, 128, 2),
                                    (128, 1000, 3), (128, 13784, 2), (64, 1, 1), (64, 127, 2), (64, 129, 3), (64, 2000, 2)])
def test_conv_transpose_x4_h_matches_torch(CI, L, B):
    """One upsampling layer + leaky_relu of the wide vocoder stages (voc_up_h.cu, polyphase ConvTranspose1d on channel-last
    16-bit split operands) against torch's conv_transpose1d (components.py:225-241): both shapes, lengths around the
    128-row tile and the C3 lengths of stages 0 and 1."""
    from models import _native as nat
    import torch.nn.functional as F
    lib = nat.lib()
    CO = CI // 2
    g = torch.Generator().manual_seed(13 * L + B + CI)
    x = torch.randn(B, CI, L, generator=g)
    w = torch.randn(CI, CO, 8, generator=g) * (1.0 / (2 * CI) ** 0.5)
    b = torch.randn(CO, generator=g) * 0.3
    want = F.leaky_relu(F.conv_transpose1d(x, w, b, stride=4, padding=2), 0.1).transpose(1, 2).contiguous()
    xd, wd, bd = (t.to(DEV) for t in (x.transpose(1, 2).contiguous(), w, b))
    y = torch.full(want.shape, float("nan"), device=DEV)
    ws = torch.empty(lib.m2tts_conv_transpose_x4_h_workspace_bytes(B, CI, L), dtype=torch.uint8, device=DEV)
    nat.check(lib.m2tts_conv_transpose_x4_h(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), y.data_ptr(), B, CI, L, STATUS.data_ptr(), ws.data_ptr(),
                                            ws.numel(), None), "conv_transpose_x4_h")
    torch.cuda.synchronize()
    assert not torch.isnan(y).any(), "unwritten output rows"
    assert H.max_abs(y.cpu(), want) <= FP32_TOL, (CI, L, B)


@pytest.mark.parametrize("prec", ["split16", "ffma", "tf32"])  # channel-last 16-bit split chain (default), FFMA everywhere, TF32 split chain
def test_vocoder_precisions_all_meet_fp32_tolerance(prec):
    from models import _native as nat
    with nat.precision(prec):
        for stage, B, T in (("stage2", 2, 203), ("stage1", 3, 130)):
            m = cuda_model(stage, perturb=6)
            M = H.STAGE_KWARGS[stage]["mel_channels"]
            mel = torch.randn(B, M, T, generator=torch.Generator().manual_seed(T))
            got = m.vocoder(mel.to(DEV))
            assert H.max_abs(got.cpu(), oracle.vocoder(cpu_sd(m), mel)) <= FP32_TOL, stage


def test_stage1_and_stage2_vocoders_run_the_16_bit_split_chain():
    """Both shipped configurations (stage1_poc: C = 128, stage2_quality: C = 256) run every vocoder stage on the channel-last
    16-bit split tcgen05 kernels — no FFMA or TF32 tap-GEMM stage (the stage-1 smoke ran conv3 / convT / tapgemm)."""
    import ctypes as C
    from models import _native as nat
    for stage in ("stage1", "stage2"):
        kw = H.STAGE_KWARGS[stage]
        kinds = (C.c_int * 5)()
        nat.check(nat.lib().m2tts_vocoder_plan(kw["mel_channels"], kw["vocoder_channels"], -1, None, kinds), "vocoder_plan")
        assert kinds[0] == 2, (stage, list(kinds))
        assert all(k in (3, 4, 5) for k in list(kinds)[1:]), (stage, list(kinds))
        m = cuda_model(stage)
        before = nat.launch_count()
        m.vocoder(torch.randn(2, kw["mel_channels"], 50, device=DEV))
        torch.cuda.synchronize()
        nat.stage_timing_enable(True)
        m.vocoder(torch.randn(2, kw["mel_channels"], 50, device=DEV))
        torch.cuda.synchronize()
        nat.stage_timing_enable(False)
        st = nat.stage_timing_read()
        assert "voc_out" not in st, st      # the output conv is fused into the last stage
