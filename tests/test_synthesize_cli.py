"""The step right after the path: checkpoint loader, PCM16 wav writer and the
synthesize.py-compatible CLI. CPU tests cover the host logic; the GPU tests run the CLI end to end against the oracle."""
import importlib.util
import sys
import wave
from pathlib import Path

import numpy as np
import pytest
import torch

import helpers as H
from helpers import oracle
from utils.audio import save_audio, to_pcm16
from utils.checkpoint import load_model, read_checkpoint
from utils.text import TextProcessor

ROOT = Path(__file__).resolve().parents[1]


def _cli():
    spec = importlib.util.spec_from_file_location("b200_synthesize", ROOT / "m2-tts_b200" / "scripts" / "synthesize.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _read_wav(path):
    with wave.open(str(path), "rb") as f:
        assert f.getnchannels() == 1 and f.getsampwidth() == 2
        return f.getframerate(), np.frombuffer(f.readframes(f.getnframes()), dtype="<i2")


def test_pcm16_cpu_rounding_and_clipping(tmp_path):
    x = np.array([0.0, 1.0, -1.0, 1.5, -2.0, 0.5 / 32767, 1.5 / 32767, 2.5 / 32767, -0.25], dtype=np.float32)
    want = np.array([0, 32767, -32767, 32767, -32767, 0, 2, 2, -8192], dtype=np.int16)
    assert (to_pcm16(x) == want).all()
    save_audio(torch.from_numpy(x).reshape(1, 1, -1), tmp_path / "a.wav", 16000)
    sr, pcm = _read_wav(tmp_path / "a.wav")
    assert sr == 16000 and (pcm == want).all()


def test_checkpoint_loader_plain_and_missing(tmp_path):
    m = H.product_model("tiny")
    cfg = {"model": {"text_encoder": {"vocab_size": 256, "hidden_dim": 32, "num_layers": 1, "num_heads": 2, "dropout": 0.1},
                     "decoder": {"mel_channels": 16, "num_layers": 1}, "vocoder": {"hidden_channels": 32}}}
    kw = H.STAGE_KWARGS["tiny"]
    cfg["model"]["text_encoder"].update(vocab_size=kw["vocab_size"], hidden_dim=kw["hidden_dim"], num_layers=kw["text_encoder_layers"],
                                        num_heads=kw["num_heads"], dropout=kw["dropout"])
    cfg["model"]["decoder"].update(mel_channels=kw["mel_channels"], num_layers=kw["decoder_layers"])
    cfg["model"]["vocoder"].update(hidden_channels=kw["vocoder_channels"])
    torch.save({"model_state_dict": m.state_dict(), "config": cfg, "step": 7}, tmp_path / "c.pt")
    ck = read_checkpoint(tmp_path / "c.pt")
    assert ck["step"] == 7 and ck["config"]["model"]["decoder"]["mel_channels"] == kw["mel_channels"]
    m2, _ = load_model(tmp_path / "c.pt", torch.device("cpu"))
    assert not m2.training
    for k, v in m.state_dict().items():
        assert torch.equal(v, m2.state_dict()[k]), k
    with pytest.raises(FileNotFoundError):
        read_checkpoint(tmp_path / "nope.pt")


def test_checkpoint_loader_survives_pickled_config_objects(tmp_path):
    """A trainer checkpoint carries an OmegaConf DictConfig; omegaconf is not installed here, so emulate a foreign
    config class living in a module named `omegaconf` and check the stub unpickler recovers the values."""
    import types
    fake = types.ModuleType("omegaconf")

    class DictConfig:                       # minimal stand-in: payload under _content like the real class
        def __init__(self, content):
            self.__dict__["_content"] = {k: DictConfig(v) if isinstance(v, dict) else v for k, v in content.items()}
    DictConfig.__module__ = "omegaconf"
    DictConfig.__qualname__ = "DictConfig"
    fake.DictConfig = DictConfig
    sys.modules["omegaconf"] = fake
    try:
        kw = H.STAGE_KWARGS["tiny"]
        cfg = DictConfig({"model": {"text_encoder": {"vocab_size": kw["vocab_size"], "hidden_dim": kw["hidden_dim"],
                                                     "num_layers": kw["text_encoder_layers"], "num_heads": kw["num_heads"],
                                                     "dropout": kw["dropout"]},
                                    "decoder": {"mel_channels": kw["mel_channels"], "num_layers": kw["decoder_layers"]},
                                    "vocoder": {"hidden_channels": kw["vocoder_channels"]}}})
        torch.save({"model_state_dict": H.product_model("tiny").state_dict(), "config": cfg}, tmp_path / "o.pt")
    finally:
        del sys.modules["omegaconf"]
    m, ck = load_model(tmp_path / "o.pt", torch.device("cpu"))
    assert ck["config"]["model"]["vocoder"]["hidden_channels"] == kw["vocoder_channels"]
    assert m.hidden_dim == kw["hidden_dim"] if hasattr(m, "hidden_dim") else True


def test_checkpoint_loader_refuses_code_execution_and_corrupt_files(tmp_path):
    """the OmegaConf fallback must not become a generic unpickler. A checkpoint that smuggles a callable next to an
    omegaconf object is refused (the payload never runs); a truncated file is reported as such, not retried."""
    import pickle
    import types
    marker = tmp_path / "pwned"

    class Evil:
        def __reduce__(self):
            import os
            return (os.system, (f"touch {marker}",))

    fake = types.ModuleType("omegaconf")

    class DictConfig:
        def __init__(self):
            self.__dict__["_content"] = {}
    DictConfig.__module__ = "omegaconf"
    DictConfig.__qualname__ = "DictConfig"
    fake.DictConfig = DictConfig
    sys.modules["omegaconf"] = fake
    try:
        torch.save({"model_state_dict": {}, "config": DictConfig(), "extra": Evil()}, tmp_path / "evil.pt")
    finally:
        del sys.modules["omegaconf"]
    with pytest.raises(pickle.UnpicklingError):
        read_checkpoint(tmp_path / "evil.pt")
    assert not marker.exists()
    # no omegaconf object at all: weights_only's own refusal is passed on, nothing is retried
    torch.save({"model_state_dict": {}, "extra": Evil()}, tmp_path / "evil2.pt")
    with pytest.raises(pickle.UnpicklingError):
        read_checkpoint(tmp_path / "evil2.pt")
    assert not marker.exists()
    good = tmp_path / "good.pt"
    torch.save({"model_state_dict": H.product_model("tiny").state_dict()}, good)
    (tmp_path / "cut.pt").write_bytes(good.read_bytes()[: good.stat().st_size // 2])
    with pytest.raises(Exception) as ei:
        read_checkpoint(tmp_path / "cut.pt")
    assert not isinstance(ei.value, FileNotFoundError)


def test_cli_rejects_ambiguous_input(tmp_path):
    with pytest.raises(SystemExit):
        _cli().main(["--checkpoint", str(tmp_path / "x.pt")])


@pytest.mark.gpu
def test_pcm16_kernel_matches_numpy():
    g = torch.Generator().manual_seed(3)
    for n in (1, 3, 4, 1001, 220544):
        x = (torch.rand(n, generator=g) * 2.4 - 1.2)
        assert (to_pcm16(x.cuda()) == to_pcm16(x)).all(), n


@pytest.mark.gpu
def test_cli_end_to_end_matches_oracle(tmp_path):
    """`synthesize.py --text ... --checkpoint ...` on the GPU == oracle inference on the CPU, to one PCM16 step."""
    m = H.product_model("stage1", perturb=3)
    torch.save({"model_state_dict": m.state_dict(), "step": 1,
                "config": {"model": {"text_encoder": {"vocab_size": 256, "hidden_dim": 64, "num_layers": 2, "num_heads": 2, "dropout": 0.1},
                                     "decoder": {"mel_channels": 64}, "vocoder": {"hidden_channels": 128}}}}, tmp_path / "s1.pt")
    text = "Hello world this is a long way down"
    rc = _cli().main(["--text", text, "--checkpoint", str(tmp_path / "s1.pt"), "--output", str(tmp_path / "o.wav"),
                      "--duration-scale", "6.0"])
    assert rc == 0
    sr, pcm = _read_wav(tmp_path / "o.wav")
    r = TextProcessor().process_text(text, max_length=256)
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    mel, audio = oracle.inference(sd, torch.tensor([r["phoneme_ids"]]), torch.tensor([r["length"]]), 6.0)
    want = to_pcm16(audio[0, 0])
    assert sr == 22050 and pcm.shape == want.shape and want.shape[0] > 64
    assert np.abs(pcm.astype(np.int32) - want.astype(np.int32)).max() <= 4     # 1e-4 x 32767 = 3.3 PCM steps
