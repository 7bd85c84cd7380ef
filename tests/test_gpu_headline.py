"""GPU parity on the configuration bench.py measures (BASELINE.json configs[2], "C3": stage2 decoder + vocoder on 3446-frame
utterances) and on the code paths around it: the host-to-host pipeline behind the `e2e` number, the N>1 output gather, the
status word (fp16 range of the 16-bit split, out-of-range ids), the packed-weight cache, deferred status / CUDA-graph replay.

Reference semantics: tts_model.py:211-228 (decoder), :279-297 (vocoder), :126-178 (length regulator), :350-400 (forward).
Tolerances: fp32 mel / waveform max-abs <= 1e-4 (BASELINE.json north_star); copies and shards bit-exact.
"""
import os
import sys

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
DEV = "cuda:0"
C3_T = 3446          # 10.002 s at 64 samples per frame, 22 050 Hz


def cuda_model(stage, perturb=None, **override):
    return H.product_model(stage, perturb=perturb, **override).to(DEV).eval()


def cpu_sd(model):
    return {k: v.detach().cpu() for k, v in model.state_dict().items()}


# --------------------------------------------------------------------------- the benchmarked shape itself
def test_c3_decoder_then_vocoder_at_3446_frames_matches_oracle():
    """bench.py's step on its own shape: x [2, 3446, 96] -> decoder (3446-key softmax per query, 54 key tiles) -> vocoder
    -> [2, 1, 220544], perturbed weights (non-zero biases), against the CPU oracle."""
    m = cuda_model("stage2", perturb=4)
    sd = cpu_sd(m)
    x = torch.randn(2, C3_T, 96, generator=torch.Generator().manual_seed(0))
    mel = m.decoder(x.to(DEV))
    audio = m.vocoder(mel.transpose(1, 2))
    assert mel.shape == (2, C3_T, 80) and audio.shape == (2, 1, 64 * C3_T)
    want_mel = oracle.mel_decoder(sd, x, 2)
    want_audio = oracle.vocoder(sd, want_mel.transpose(1, 2))
    assert H.max_abs(mel.cpu(), want_mel) <= FP32_TOL
    assert H.max_abs(audio.cpu(), want_audio) <= FP32_TOL
    # same rows inside a larger batch (what the bench runs is 64 of these): bit-identical
    big = torch.cat([x, torch.randn(6, C3_T, 96, generator=torch.Generator().manual_seed(1))]).to(DEV)
    mel8 = m.decoder(big)
    assert torch.equal(mel8[:2], mel)
    assert torch.equal(m.vocoder(mel8.transpose(1, 2))[:2], audio)


def test_c3_through_the_real_length_regulator_s256():
    """SURVEY §8d C3, end-to-end variant: stage2, S = 256 phonemes, 13-14 frames per phoneme summing to exactly 3446 frames,
    ids -> encoder -> duration predictor -> length regulator -> decoder -> vocoder."""
    m = cuda_model("stage2", perturb=9)
    sd = cpu_sd(m)
    g = torch.Generator().manual_seed(5)
    B, S = 2, 256
    ids = torch.randint(0, 256, (B, S), generator=g)
    lengths = torch.tensor([256, 256])
    dur = torch.full((B, S), 13.0)
    extra = C3_T - 13 * S                      # 118 phonemes get a 14th frame
    for b in range(B):
        dur[b, torch.randperm(S, generator=g)[:extra]] = 14.0
    dur += torch.rand((B, S), generator=g) * 0.9          # fractional parts are truncated (tts_model.py:150)
    out = m(ids.to(DEV), lengths.to(DEV), target_durations=dur.to(DEV))
    index, frames, T = oracle.length_regulator_indices(dur.numpy())
    assert T == C3_T and (frames == C3_T).all()
    assert np.array_equal(m.length_regulator.last_frames.cpu().numpy(), frames)
    assert np.array_equal(m.length_regulator.last_index.cpu().numpy(), index)
    assert torch.equal(out["regulated_output"].cpu(), oracle.length_regulator(out["encoder_output"].cpu(), dur))
    ref = oracle.forward(sd, ids, lengths, dur)
    for k in ("encoder_output", "duration_pred", "mel_output", "audio_output"):
        assert out[k].shape == ref[k].shape, k
        assert H.max_abs(out[k].cpu(), ref[k]) <= FP32_TOL, k


def test_host_pipeline_autotune_keeps_a_valid_layout_and_the_result():
    """HostPipeline.autotune measures its candidate layouts on the real step and keeps the fastest: whatever it picks must be a
    composition of the batch, and the result left in the output buffer (and of a later run) must equal the direct call."""
    from models import _native as nat
    from utils.host_pipeline import HostPipeline
    m = cuda_model("stage2")
    x_host = torch.randn(12, 260, 96, generator=torch.Generator().manual_seed(5)).pin_memory()
    out_host = torch.empty((12, 1, 64 * 260)).pin_memory()

    def step(x):
        return m.vocoder(m.decoder(x).transpose(1, 2))

    want = step(x_host.to(DEV)).cpu()
    pipe = HostPipeline(torch.device(DEV), n_chunks=2)
    out_host.fill_(float("nan"))
    with nat.deferred_status():
        sizes = pipe.autotune(step, x_host, out_host, frames=260, heads=2, reps=1)
    nat.check_status(torch.device(DEV))
    assert sum(sizes) == 12 and min(sizes) >= 1 and pipe.sizes == sizes
    assert pipe.tuned is not None and len(pipe.tuned["ms"]) == len(pipe.tuned["candidates"]) and sizes in pipe.tuned["candidates"]
    assert torch.equal(out_host, want)
    out_host.fill_(float("nan"))
    with nat.deferred_status():
        pipe.run(step, x_host, out_host)
    pipe.synchronize()
    nat.check_status(torch.device(DEV))
    assert torch.equal(out_host, want)


# --------------------------------------------------------------------------- host-to-host pipeline (the e2e number)
@pytest.mark.parametrize("n_chunks", [1, 3, 5])
def test_host_pipeline_equals_direct_call(n_chunks):
    """HostPipeline.run(step, x_host, out_host) == step(x.cuda()).cpu(), bit for bit, for 1 / 3 / 5 chunks (utterances are
    independent in eval mode, so chunking the batch must not change a single bit), twice in a row through the same buffers."""
    from models import _native as nat
    from utils.host_pipeline import HostPipeline
    m = cuda_model("stage2")
    x_host = torch.randn(7, 300, 96, generator=torch.Generator().manual_seed(3)).pin_memory()
    out_host = torch.empty((7, 1, 64 * 300)).pin_memory()

    def step(x):
        return m.vocoder(m.decoder(x).transpose(1, 2))

    want = step(x_host.to(DEV)).cpu()
    pipe = HostPipeline(torch.device(DEV), n_chunks=n_chunks)
    for rep in range(2):
        out_host.fill_(float("nan"))
        with nat.deferred_status():
            pipe.run(step, x_host, out_host)
        pipe.synchronize()
        nat.check_status(torch.device(DEV))
        assert torch.equal(out_host, want), (n_chunks, rep)


def test_host_pipeline_with_explicit_chunk_sizes_equals_direct_call():
    """The wave-filling layout the bench uses (small edge chunks around larger inner ones, `HostPipeline(sizes=...)`) must not
    change a bit either; sizes that do not sum to the batch fall back to the equal split."""
    from models import _native as nat
    from utils.host_pipeline import HostPipeline
    m = cuda_model("stage2")
    x_host = torch.randn(9, 300, 96, generator=torch.Generator().manual_seed(4)).pin_memory()
    out_host = torch.empty((9, 1, 64 * 300)).pin_memory()

    def step(x):
        return m.vocoder(m.decoder(x).transpose(1, 2))

    want = step(x_host.to(DEV)).cpu()
    for sizes in ([1, 4, 3, 1], [2, 7], [5, 27, 27, 5]):
        pipe = HostPipeline(torch.device(DEV), n_chunks=2, sizes=sizes)
        out_host.fill_(float("nan"))
        with nat.deferred_status():
            pipe.run(step, x_host, out_host)
        pipe.synchronize()
        nat.check_status(torch.device(DEV))
        assert torch.equal(out_host, want), sizes
    # an input that is already on the device (the ids path of synthesize_to_host): chunked views, only the results are copied
    pipe = HostPipeline(torch.device(DEV), sizes=[2, 4, 3])
    out_host.fill_(float("nan"))
    with nat.deferred_status():
        pipe.run(step, x_host.to(DEV), out_host)
    pipe.synchronize()
    nat.check_status(torch.device(DEV))
    assert torch.equal(out_host, want)


# --------------------------------------------------------------------------- status word
def test_activations_beyond_the_fp16_range_fall_back_to_tf32():
    """the 16-bit split carries operands as fp16 hi + lo. Weights scaled so that (a) the FFN hidden activations
    and (b) the vocoder's first activation exceed 65 504 must not give a silently wrong result: the stage raises the status
    bit and runs again with the TF32 split, <= 1e-4 against the oracle; under deferred_status the error surfaces instead."""
    from models import _native as nat
    m = cuda_model("stage2", perturb=4)
    with torch.no_grad():
        for layer in m.decoder.layers:
            layer.ffn.linear1.weight.mul_(3e5); layer.ffn.linear1.bias.mul_(3e5)
            layer.ffn.linear2.weight.mul_(1.0 / 3e5)
        m.vocoder.input_conv.weight.mul_(1e5); m.vocoder.input_conv.bias.mul_(1e5)
        m.vocoder.upsamples[0].weight.mul_(1e-5)
    sd = cpu_sd(m)
    x = torch.randn(2, 200, 96, generator=torch.Generator().manual_seed(2))
    want_mel = oracle.mel_decoder(sd, x, 2)
    want_audio = oracle.vocoder(sd, want_mel.transpose(1, 2))
    with pytest.warns(RuntimeWarning, match="fp16 range"):
        nat._range_warned = False
        mel = m.decoder(x.to(DEV))
    assert H.max_abs(mel.cpu(), want_mel) <= FP32_TOL
    audio = m.vocoder(want_mel.to(DEV).transpose(1, 2))
    assert H.max_abs(audio.cpu(), want_audio) <= FP32_TOL
    assert m.decoder.__dict__.get("_m2tts_tf32_only") and m.vocoder.__dict__.get("_m2tts_tf32_only")
    assert H.max_abs(m.decoder(x.to(DEV)).cpu(), want_mel) <= FP32_TOL          # second call starts on the TF32 split
    # explicit 16-bit split, deferred check: a clean error, never a silent result
    nat.drop_packed(m)
    with nat.precision("split16"), nat.deferred_status():
        m.decoder(x.to(DEV))
    torch.cuda.synchronize()
    with pytest.raises(nat.Fp16RangeError):
        nat.check_status(torch.device(DEV))
    # non-finite input: flagged as well (a clamp would have turned NaN into -65000)
    bad = x.clone(); bad[0, 7, 3] = float("nan")
    with nat.precision("split16"), nat.deferred_status():
        cuda_model("stage2").decoder(bad.to(DEV))
    torch.cuda.synchronize()
    with pytest.raises(nat.Fp16RangeError):
        nat.check_status(torch.device(DEV))


def test_weights_beyond_the_fp16_range_fall_back_to_tf32():
    from models import _native as nat
    m = cuda_model("stage2", perturb=4)
    with torch.no_grad():
        m.decoder.mel_projection.weight.mul_(1e6)          # |w| up to ~2e5 > 65504
    sd = cpu_sd(m)
    x = torch.randn(1, 150, 96, generator=torch.Generator().manual_seed(6))
    want = oracle.mel_decoder(sd, x, 2)
    nat._range_warned = True
    got = m.decoder(x.to(DEV)).cpu()
    assert torch.isfinite(got).all()
    assert H.max_abs(got, want) <= 1e-5 * float(want.abs().max())


def test_out_of_range_phoneme_id_raises_index_error():
    """nn.Embedding raises on an id outside the table (tts_model.py:78); the kernel used to clamp silently."""
    m = cuda_model("tiny")
    ids = torch.randint(0, 256, (2, 9), generator=torch.Generator().manual_seed(1))
    for bad in (256, -1, 10 ** 9):
        t = ids.clone(); t[1, 4] = bad
        with pytest.raises(IndexError):
            m.text_encoder(t.to(DEV), torch.tensor([9, 9], device=DEV))
    m.text_encoder(ids.to(DEV), torch.tensor([9, 9], device=DEV))      # the word is clear again


# --------------------------------------------------------------------------- packed weights / stateless ABI / graphs
def test_packed_weight_cache_follows_the_parameters():
    from models import _native as nat
    m = cuda_model("stage2", perturb=3)
    x = torch.randn(1, 130, 96, generator=torch.Generator().manual_seed(8))

    def step():
        return m.vocoder(m.decoder(x.to(DEV)).transpose(1, 2)).cpu()

    first = step()
    base = nat.launch_count()
    again = step()
    warm = nat.launch_count() - base
    assert torch.equal(first, again)
    nat.drop_packed(m)
    base = nat.launch_count()
    step()
    cold = nat.launch_count() - base
    assert cold >= warm + 8, (cold, warm)            # packing is 3 layer + 1 projection + >= 5 vocoder launches
    assert warm <= 32, warm                            # <= 32 launches per decoder+vocoder step
    with torch.no_grad():                              # in-place update bumps _version: the images must be rebuilt
        m.decoder.layers[0].ffn.linear1.weight.mul_(1.5)
        m.vocoder.resblocks[2].conv1.weight.mul_(0.5)
    sd = cpu_sd(m)
    want = oracle.vocoder(sd, oracle.mel_decoder(sd, x, 2).transpose(1, 2))
    assert H.max_abs(step(), want) <= FP32_TOL
    m2 = cuda_model("stage2", perturb=5)               # load_state_dict into the same module
    m.load_state_dict(m2.state_dict())
    sd = cpu_sd(m)
    want = oracle.vocoder(sd, oracle.mel_decoder(sd, x, 2).transpose(1, 2))
    assert H.max_abs(step(), want) <= FP32_TOL


def test_stateless_entry_points_without_packed_buffers():
    """The C ABI's stateless form (packed == NULL: images written into the workspace per call) gives the same bits."""
    import ctypes as C
    from models import _native as nat
    from models.tts_model import _layer_struct
    lib = nat.lib()
    m = cuda_model("stage2", perturb=3)
    x = torch.randn(2, 77, 96, generator=torch.Generator().manual_seed(4)).to(DEV)
    want = m.decoder.layers[0](x)
    st = _layer_struct(m.decoder.layers[0])
    out = torch.empty_like(x)
    ws = torch.empty(lib.m2tts_transformer_workspace_bytes(2, 77, 96, 192), dtype=torch.uint8, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    nat.check(lib.m2tts_transformer_layer(C.byref(st), None, x.data_ptr(), out.data_ptr(), None, 2, 77, 96, 2, 192, 1e-5, nat.PREC_DEFAULT,
                                          status.data_ptr(), ws.data_ptr(), ws.numel(), None), "transformer_layer")
    torch.cuda.synchronize()
    assert int(status.item()) == 0 and torch.equal(out, want)


def test_cuda_graph_replay_of_the_step_is_bit_identical():
    """Small-batch path: decoder + vocoder captured once in a CUDA graph (weights packed, workspaces allocated by the warm-up,
    status deferred) and replayed on new inputs."""
    from utils.graph import GraphedStep
    m = cuda_model("stage2", perturb=1)

    def step(x):
        return m.vocoder(m.decoder(x).transpose(1, 2))

    g = torch.Generator().manual_seed(12)
    x0 = torch.randn(1, 128, 96, generator=g).to(DEV)
    graphed = GraphedStep(step, x0)
    for _ in range(3):
        x = torch.randn(1, 128, 96, generator=g).to(DEV)
        assert torch.equal(graphed(x), step(x))
    assert graphed.launches_captured <= 32


# --------------------------------------------------------------------------- vocoder options no config exercises
@pytest.mark.parametrize("stage,dils", [("stage2", (2, 3, 1, 1)), ("stage1", (1, 2, 4, 3)), ("stage2", (1, 1, 2, 2))])
def test_vocoder_forward_with_dilated_resblocks(stage, dils):
    """LightweightResBlock accepts a dilation for conv1 (components.py:177-190); SimpleVocoder always passes 1, so build the
    blocks by hand and check m2tts_vocoder_forward's per-stage dispatch against torch."""
    import torch.nn.functional as F
    from models.components import LightweightResBlock
    m = H.product_model(stage, perturb=6)
    for j, d in enumerate(dils):
        old = m.vocoder.resblocks[j]
        new = LightweightResBlock(old.conv1.in_channels, 3, dilation=d)
        new.load_state_dict(old.state_dict())
        m.vocoder.resblocks[j] = new
    m = m.to(DEV).eval()
    sd = cpu_sd(m)
    M = H.STAGE_KWARGS[stage]["mel_channels"]
    mel = torch.randn(2, M, 150, generator=torch.Generator().manual_seed(3))
    x = F.conv1d(mel, sd["vocoder.input_conv.weight"], sd["vocoder.input_conv.bias"], padding=1)
    for j, r in enumerate((4, 4, 2, 2)):
        x = F.leaky_relu(F.conv_transpose1d(x, sd[f"vocoder.upsamples.{j}.weight"], sd[f"vocoder.upsamples.{j}.bias"], stride=r, padding=r // 2), 0.1)
        h = F.conv1d(x, sd[f"vocoder.resblocks.{j}.conv1.weight"], sd[f"vocoder.resblocks.{j}.conv1.bias"], padding=dils[j], dilation=dils[j])
        x = x + F.conv1d(F.leaky_relu(h, 0.1), sd[f"vocoder.resblocks.{j}.conv2.weight"], sd[f"vocoder.resblocks.{j}.conv2.bias"], padding=1)
    want = torch.tanh(F.conv1d(x, sd["vocoder.output_conv.weight"], sd["vocoder.output_conv.bias"], padding=1))
    assert H.max_abs(m.vocoder(mel.to(DEV)).cpu(), want) <= FP32_TOL


# --------------------------------------------------------------------------- N > 1: NCCL gather of sharded outputs
def _nccl_worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      NCCL_DEBUG="WARN")
    import torch.distributed as dist
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from utils.shard import synthesize_sharded
        m = H.product_model("stage1", perturb=1).to(dev).eval()
        ids, lengths, dur = H.c2_inputs()
        ids, lengths, dur = ids[:5].to(dev), lengths[:5].to(dev), dur[:5].to(dev)       # 5 utterances -> uneven shards 3 + 2
        got = synthesize_sharded(m, ids, lengths, dur, None)
        full = m(ids, lengths, target_durations=dur)
        ok = got["max_target_length"] == full["mel_output"].shape[1]
        ok = ok and torch.equal(got["mel_output"], full["mel_output"]) and torch.equal(got["audio_output"], full["audio_output"])
        one = synthesize_sharded(m, ids[:1], lengths[:1], dur[:1], None)               # fewer utterances than ranks
        full1 = m(ids[:1], lengths[:1], target_durations=dur[:1])
        ok = ok and torch.equal(one["audio_output"], full1["audio_output"])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_sharded_synthesis_over_nccl_equals_full_batch():
    """SURVEY §8e on hardware: two ranks, contiguous utterance blocks, one all-reduce(MAX) of the frame maximum, outputs
    all-gathered over NCCL after the path; every rank ends up with exactly the single-process batch."""
    import torch.multiprocessing as mp
    world = 2
    port = 29700 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_nccl_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


@pytest.mark.parametrize("n_chunks", [1, 3])
def test_synthesize_to_host_equals_forward(n_chunks):
    """utils.host_pipeline.synthesize_to_host (what bench.py reports as `e2e_from_ids`: acoustic front once, decoder + vocoder in
    utterance chunks, waveform copies on a side stream) returns bit-for-bit what one M2TTSModel.forward returns
    (reference call site scripts/synthesize.py:66-83; tts_model.py:350-400)."""
    from utils.host_pipeline import HostPipeline, synthesize_to_host
    m = cuda_model("stage2", perturb=5)
    ids, lengths, dur = H.small_inputs(7, 24, 256, seed=3)
    T = int(dur.trunc().sum(dim=1).max())
    want = m(ids.to(DEV), lengths.to(DEV), target_durations=dur.to(DEV), max_target_length=T)["audio_output"].cpu()
    out = torch.full((7, 1, 64 * T), float("nan")).pin_memory()
    pipe = HostPipeline(torch.device(DEV), n_chunks=n_chunks)
    synthesize_to_host(m, ids.pin_memory(), lengths.pin_memory(), dur.pin_memory(), T, out, pipe)
    pipe.synchronize()
    assert torch.equal(out, want)
