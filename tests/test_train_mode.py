"""the mirror model drops into the reference's training step (training/train.py:290-342) — train-mode
forward + backward, on the CPU here and on CUDA on the GPU box, against the oracle's torch formulation of the same maths."""
import importlib.util
import sys

import pytest
import torch

import helpers as H
from helpers import oracle


def _train_module():
    spec = importlib.util.spec_from_file_location("m2tts_train", H.REPO / "m2-tts_b200" / "training" / "train.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["m2tts_train"] = mod
    spec.loader.exec_module(mod)
    return mod


def _reference_loss(mel_pred, mel_target, dur_pred, dur_target, mel_lengths, w_mel=1.0, w_dur=0.1):
    """training/train.py:66-105, literally (per-utterance loop with .item())."""
    mel_target = mel_target.transpose(1, 2)
    mel_loss = 0
    for i in range(mel_pred.size(0)):
        n = mel_lengths[i].item()
        mel_loss = mel_loss + torch.nn.functional.l1_loss(mel_pred[i, :n, :], mel_target[i, :n, :])
    mel_loss = mel_loss / mel_pred.size(0)
    return w_mel * mel_loss + w_dur * torch.nn.functional.mse_loss(dur_pred, dur_target)


def _batch(n=4, seed=11, mel_dim=64):
    from data.dataset import DummyDataset, collate_fn
    ds = DummyDataset(size=n, max_text_length=24, max_mel_length=90, mel_dim=mel_dim, seed=seed)
    return collate_fn([ds[i] for i in range(n)])


def test_dummy_dataset_and_collate_follow_the_reference_layout():
    b = _batch()
    B, S = b["phoneme_ids"].shape
    assert b["phoneme_ids"].dtype == torch.long and b["text_lengths"].dtype == torch.long and b["mel_lengths"].dtype == torch.long
    assert b["mel_specs"].shape[:2] == (B, 64) and b["durations"].shape == (B, S) and len(b["texts"]) == B
    assert int(b["text_lengths"].max()) == S and int(b["mel_lengths"].max()) == b["mel_specs"].shape[2]
    for i in range(B):      # padding is zero, durations sum to the mel length (src/data/dataset.py:340-343)
        s, t = int(b["text_lengths"][i]), int(b["mel_lengths"][i])
        assert (b["phoneme_ids"][i, s:] == 0).all() and (b["mel_specs"][i, :, t:] == 0).all() and (b["durations"][i, s:] == 0).all()
        assert abs(float(b["durations"][i].sum()) - t) < 1e-3
    assert torch.equal(_batch()["mel_specs"], b["mel_specs"])      # seeded items are reproducible


def _grads_vs_oracle(device):
    tr = _train_module()
    # fp32 autograd on the GPU for this comparison: by default torch lets cuDNN / cuBLAS use TF32 for convolutions (the duration
    # predictor), which is a training-precision choice of the caller and not part of what is checked here
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = H.product_model("stage1", perturb=3, dropout=0.0).to(device).train()
    for mod in m.duration_predictor.modules():      # BatchNorm batch statistics are not part of the oracle's eval formulation
        if isinstance(mod, torch.nn.BatchNorm1d):
            mod.eval()
    b = _batch()
    bd = {k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    out = m(bd["phoneme_ids"], bd["text_lengths"], target_durations=bd["durations"], max_target_length=bd["mel_specs"].size(2))
    assert out["audio_output"] is None
    losses = tr.TTSLoss()(out["mel_output"], bd["mel_specs"], out["duration_pred"], bd["durations"], bd["mel_lengths"])
    losses["total_loss"].backward()
    # the oracle's functions are differentiable torch ops: same state_dict as leaf tensors on the CPU
    param_names = {n for n, _ in m.named_parameters()}
    sd = {k: v.detach().cpu().clone().requires_grad_(k in param_names) for k, v in m.state_dict().items()}
    enc, mask = oracle.text_encoder(sd, b["phoneme_ids"], b["text_lengths"], 2)
    dur = oracle.duration_predictor(sd, enc)
    idx, _, T = oracle.length_regulator_indices(b["durations"].numpy(), b["mel_specs"].size(2))
    idx = torch.from_numpy(idx.astype("int64"))
    reg = torch.where((idx >= 0)[:, :, None], torch.gather(enc, 1, idx.clamp(min=0)[:, :, None].expand(-1, -1, enc.shape[2])),
                      torch.zeros(()))
    mel = oracle.mel_decoder(sd, reg, 2)
    want = _reference_loss(mel, b["mel_specs"], dur, b["durations"], b["mel_lengths"])
    want.backward()
    assert abs(float(losses["total_loss"].detach()) - float(want.detach())) <= 1e-5 * max(1.0, abs(float(want.detach())))
    checked = 0
    for name, p in m.named_parameters():
        if p.grad is None:
            assert name.startswith("vocoder."), name      # no vocoder in the training forward (tts_model.py:388)
            continue
        g_ref = sd[name].grad
        assert g_ref is not None, name
        scale = float(g_ref.abs().max()) + 1e-8
        assert float((p.grad.cpu() - g_ref).abs().max()) <= 2e-4 * scale + 1e-7, name
        checked += 1
    assert checked >= 40


def test_train_step_gradients_match_the_oracle_formulation_cpu():
    _grads_vs_oracle("cpu")


@pytest.mark.gpu
def test_train_step_gradients_match_the_oracle_formulation_cuda():
    _grads_vs_oracle("cuda:0")


def _loss_goes_down(device):
    tr = _train_module()
    torch.manual_seed(7)
    m = H.product_model("tiny", dropout=0.0).to(device)
    opt = torch.optim.AdamW(m.parameters(), lr=2e-3)
    b = _batch(n=4, mel_dim=32)
    first = last = None
    for _ in range(12):
        metrics = tr.train_step(m, dict(b), tr.TTSLoss(), opt, torch.device(device))
        first = metrics["total_loss"] if first is None else first
        last = metrics["total_loss"]
    assert m.training and last < first
    m.eval()      # and the same module still serves synthesis afterwards (weight-image caches follow the updated parameters)
    return m


def test_training_steps_reduce_the_loss_cpu():
    _loss_goes_down("cpu")


@pytest.mark.gpu
def test_training_then_synthesis_on_cuda():
    m = _loss_goes_down("cuda:0")
    sd = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    ids, lengths, dur = H.small_inputs(2, 10, 256, seed=11)
    out = m(ids.cuda(), lengths.cuda(), dur.cuda(), 40)
    ref = oracle.forward(sd, ids, lengths, dur, 40)
    assert H.max_abs(out["mel_output"].cpu(), ref["mel_output"]) <= 1e-4 and H.max_abs(out["audio_output"].cpu(), ref["audio_output"]) <= 1e-4
    # one more optimiser step must invalidate the cached weight images
    import importlib
    tr = sys.modules["m2tts_train"]
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    tr.train_step(m, dict(_batch(n=4, mel_dim=32)), tr.TTSLoss(), opt, torch.device("cuda:0"))
    m.eval()
    sd2 = {k: v.detach().cpu() for k, v in m.state_dict().items()}
    out2 = m(ids.cuda(), lengths.cuda(), dur.cuda(), 40)
    ref2 = oracle.forward(sd2, ids, lengths, dur, 40)
    assert H.max_abs(out2["mel_output"].cpu(), ref2["mel_output"]) <= 1e-4 and H.max_abs(out2["audio_output"].cpu(), ref2["audio_output"]) <= 1e-4
