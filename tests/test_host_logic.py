"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares, the
eval-mode path refuses to run without CUDA (no fallback), the train-mode (autograd) formulation
matches the oracle, host utilities behave like the reference's."""
import ctypes as C
import re
import wave

import numpy as np
import pytest
import os

import torch

import helpers as H
from helpers import oracle


def test_header_symbols_are_exported_and_bound():
    from models import _native as nat
    header = (H.REPO / "include" / "m2tts_b200.h").read_text()
    declared = set(re.findall(r"\b(m2tts_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    lib = C.CDLL(str(nat.library_path()))
    for sym in declared:
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    assert declared == set(nat.EXPORTED_SYMBOLS), declared ^ set(nat.EXPORTED_SYMBOLS)
    assert nat.lib().m2tts_version() >= 200
    # the product library carries no bring-up hooks: those live in libm2tts_b200_tools.so (include/m2tts_b200_tools.h)
    tools_hdr = (H.REPO / "include" / "m2tts_b200_tools.h").read_text()
    tool_syms = set(re.findall(r"\b(m2tts_[a-z0-9_]+)\s*\(", tools_hdr))
    assert tool_syms == set(nat.TOOL_SYMBOLS)
    if nat.library_path().name == "libm2tts_b200.so":
        for sym in tool_syms:
            assert not hasattr(lib, sym), f"{sym} is a bring-up hook and must not be exported by the product library"


def test_workspace_queries_without_gpu():
    from models import _native as nat
    lib = nat.lib()
    assert lib.m2tts_transformer_workspace_bytes(64, 3446, 96, 192) > 64 * 3446 * 96 * 4 * 9
    assert lib.m2tts_vocoder_workspace_bytes(64, 3446, 80, 256) > 3 * 64 * 256 * 3446 * 4 * 4
    assert lib.m2tts_vocoder_workspace_bytes(1, 1, 80, 8) == 0     # hidden_channels must be >= 16
    assert lib.m2tts_vocoder_forward(None, None, None, 0, 0, 0, None, 1, 1, 1, 16, -1, None, None, 0, None) == -5
    assert b"null" in lib.m2tts_last_error_string()
    assert lib.m2tts_vocoder_pack_bytes(80, 256, -1) > 2 * 530000 * 4      # every weight at least as hi + lo
    assert lib.m2tts_transformer_pack_bytes(96, 192, -1) >= 2 * 4 * (3 * 96 * 96 + 96 * 96 + 2 * 96 * 192)
    assert lib.m2tts_ln_proj_pack_bytes(96, 80, -1) >= 2 * 4 * 96 * 80


def test_eval_mode_has_no_cpu_fallback():
    from models import _native as nat
    m = H.product_model("tiny")
    ids = torch.zeros(1, 4, dtype=torch.long)
    with pytest.raises(nat.NativeLibraryError):
        m(ids)
    with pytest.raises(nat.NativeLibraryError):
        m.inference(ids)
    with pytest.raises(nat.NativeLibraryError):
        m.vocoder(torch.zeros(1, 32, 5))
    with pytest.raises(nat.NativeLibraryError):
        m.length_regulator(torch.zeros(1, 4, 32), torch.ones(1, 4))


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    from models import _native as nat
    monkeypatch.setattr(nat, "_lib", None)
    monkeypatch.setattr(nat, "_LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(nat.NativeLibraryError, match="not built"):
        nat.lib()


def test_train_mode_matches_oracle_on_cpu():
    """The autograd formulation (what training/train.py drives) is the same math as the oracle."""
    m = H.product_model("stage1", perturb=3, dropout=0.0).train()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    ids, lengths, dur = H.small_inputs(3, 20, 256, seed=5, lo=-0.5, hi=4.0)
    out = m(ids, lengths, target_durations=dur, max_target_length=70)
    assert out["audio_output"] is None          # no vocoder in training forward (reference :388)
    # BatchNorm uses batch statistics in train mode, so compare the stages around it
    enc, mask = oracle.text_encoder(sd, ids, lengths, 2)
    assert H.max_abs(out["encoder_output"].detach(), enc) <= 5e-6 and torch.equal(out["padding_mask"], mask)
    reg = oracle.length_regulator(out["encoder_output"].detach(), dur, 70)
    assert torch.equal(out["regulated_output"].detach(), reg)
    assert H.max_abs(out["mel_output"].detach(), oracle.mel_decoder(sd, reg, 2)) <= 5e-6
    out["mel_output"].sum().backward()
    assert m.text_encoder.embedding.weight.grad is not None   # gradients flow through the regulator
    wav = m.vocoder(out["mel_output"].detach().transpose(1, 2))
    assert H.max_abs(wav.detach(), oracle.vocoder(sd, out["mel_output"].detach().transpose(1, 2))) <= 5e-6


def test_length_regulator_train_path_edges():
    from models.tts_model import LengthRegulator
    lr = LengthRegulator().train()
    d = H.lr_edge_durations()
    enc = torch.randn(d.shape[0], d.shape[1], 8, generator=torch.Generator().manual_seed(3))
    for ml in (None, 20, 90):
        assert torch.equal(lr(enc, d, ml), oracle.length_regulator(enc, d, ml))
    bad = d.clone(); bad[0, 0] = float("nan")
    with pytest.raises(ValueError):
        lr(enc, bad)


def test_state_dict_layout_is_the_reference_layout():
    m = H.product_model("stage2")
    sd = m.state_dict()
    assert sd["text_encoder.pos_encoding.pe"].shape == (1, 1000, 96)
    assert sd["text_encoder.layers.2.self_attn.qkv.weight"].shape == (288, 96)
    assert "text_encoder.layers.0.self_attn.qkv.bias" not in sd
    assert sd["duration_predictor.predictor.conv_layers.1.norm.num_batches_tracked"].dtype == torch.int64
    assert sd["decoder.mel_projection.weight"].shape == (80, 96)
    assert [tuple(sd[f"vocoder.upsamples.{j}.weight"].shape) for j in range(4)] == \
        [(256, 128, 8), (128, 64, 8), (64, 32, 4), (32, 16, 4)]
    assert sd["vocoder.output_conv.weight"].shape == (1, 16, 3)
    assert m.get_model_size()["total_params"] == 1066610
    assert H.product_model("stage1").get_model_size()["total_params"] == 321154


def test_config_loader_and_wav_writer(tmp_path):
    from utils.audio import save_audio
    from utils.config import AttrDict, model_kwargs
    cfg = AttrDict.wrap({"model": {"text_encoder": {"vocab_size": 256, "hidden_dim": 64, "num_layers": 2,
                                                   "num_heads": 2, "dropout": 0.1},
                                   "decoder": {"mel_channels": 64}, "vocoder": {"hidden_channels": 128}}})
    assert model_kwargs(cfg) == H.STAGE_KWARGS["stage1"]   # stage1 YAML has no decoder.num_layers -> 2
    p = tmp_path / "a.wav"
    save_audio(torch.linspace(-1.2, 1.2, 100).reshape(1, 1, 100), p, 22050)
    with wave.open(str(p)) as f:
        assert (f.getnchannels(), f.getsampwidth(), f.getframerate(), f.getnframes()) == (1, 2, 22050, 100)
        pcm = np.frombuffer(f.readframes(100), dtype="<i2")
    assert pcm[0] == -32767 and pcm[-1] == 32767


def test_shard_bounds_cover_and_partition():
    from utils.shard import shard_bounds
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_cpulist_parsing_and_binding_without_a_gpu():
    from utils.device import _parse_cpulist, bind_host_to_gpu
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    if not torch.cuda.is_available():
        before = os.sched_getaffinity(0)
        assert bind_host_to_gpu(torch.device("cuda", 0))["bound"] is False      # no device (or one NUMA node): nothing is touched
        assert os.sched_getaffinity(0) == before


def test_entry_scripts_compile():
    """bench.py, __graft_entry__.py and the bring-up tools are only ever run on the GPU box: a syntax error there costs a round.
    Compile every one of them here."""
    import py_compile
    from pathlib import Path
    root = Path(__file__).resolve().parents[1]
    files = [root / "bench.py", root / "__graft_entry__.py"] + sorted((root / "tools").glob("*.py")) + sorted((root / "m2-tts_b200").rglob("*.py"))
    assert len(files) > 10
    for f in files:
        py_compile.compile(str(f), doraise=True)


def test_persistent_attention_item_schedule_covers_every_item_once():
    """Restatement of `AhItems` (m2-tts_b200/csrc/attention_h.cu): the list of work items a persistent attention CTA walks. Every
    item must appear in exactly one CTA's list, two-tile items first, and the single-tile items (0.5-0.66 of a two-tile one) must go
    to the CTAs that got one two-tile item less before anyone gets a third."""
    def items_of(n_long, n_single, G, c):
        out, r = [], n_long % G
        nd = G - r
        m = 0
        while c + m * G < n_long:
            out.append(c + m * G); m += 1
        lim = min(n_single, 2 * nd)
        for m in range(2):
            j = (c - r) + m * nd
            if c >= r and j < lim:
                out.append(n_long + j)
        m = 0
        while 2 * nd + c + m * G < n_single:
            out.append(n_long + 2 * nd + c + m * G); m += 1
        return out

    for n_long, n_single, sms in ((1664, 128, 148), (0, 5, 148), (0, 300, 148), (10, 5, 148), (13, 1, 148), (1000, 0, 148), (147, 147, 148),
                                  (148, 148, 148), (149, 1, 148), (296, 600, 148), (7, 3, 4), (1, 1, 148)):
        G = min(n_long + n_single, sms)
        seen = []
        loads = []
        for c in range(G):
            it = items_of(n_long, n_single, G, c)
            assert it == sorted(it), "two-tile items come first, every list is ascending"
            seen += it
            loads.append(sum(1.0 if i < n_long else 0.5 for i in it))
        assert sorted(seen) == list(range(n_long + n_single)), (n_long, n_single, G)
        # balance: no CTA carries more than one two-tile item above the least loaded one
        assert max(loads) - min(loads) <= 1.0 + 1e-9, (n_long, n_single, G, max(loads), min(loads))
    # the C3 launch: 36 CTAs with 12 two-tile items, 112 with 11 and one or two single-tile ones: makespan 12 two-tile units
    loads = [sum(1.0 if i < 1664 else 0.5 for i in items_of(1664, 128, 148, c)) for c in range(148)]
    assert max(loads) == 12.0 and min(loads) == 11.5
