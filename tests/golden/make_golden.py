"""Generate the golden fixtures by running the UNMODIFIED reference (imported from
/root/reference/src/models) on seeded weights and inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference cannot travel to the GPU box, these small .npz files can.  Everything here is
seeded; the weights are re-created on the consumer side by the same seeded construction and
checked through `state_digests.json`.
"""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import helpers as H  # noqa: E402

OUT = Path(__file__).resolve().parent


def ref_model(ref, stage, perturb=None, **override):
    kw = dict(H.STAGE_KWARGS[stage]); kw.update(override)
    torch.manual_seed(1234)
    m = ref.M2TTSModel(**kw)
    if perturb is not None:
        H.perturb_(m, perturb)
    return m.eval()


def npify(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            for k, v in d.items() if v is not None}


def main():
    torch.set_num_threads(1)  # one fixed reduction order for the stored fp32 values
    ref = H.load_reference_module()
    digests = {}

    # ---- weight digests: plain seeded init and the perturbed ("biased") second set ----
    for stage in ("tiny", "stage1", "stage2"):
        digests[f"{stage}"] = H.state_digest(ref_model(ref, stage).state_dict())
        digests[f"{stage}+perturb7"] = H.state_digest(ref_model(ref, stage, perturb=7).state_dict())

    with torch.no_grad():
        # ---- tiny model, scripts/test_pipeline.py:67-113 shape: B=2, S=10, max_target_length=50 ----
        for tag, pert in (("tiny_fwd", None), ("tiny_biased_fwd", 7)):
            m = ref_model(ref, "tiny", perturb=pert)
            ids, lengths, dur = H.small_inputs(2, 10, 256, seed=11)
            out = m(ids, lengths, target_durations=dur, max_target_length=50)
            np.savez(OUT / f"{tag}.npz", ids=ids.numpy(), lengths=lengths.numpy(), dur=dur.numpy(), **npify(out))

        # ---- stage2 small with biased weights (test_stage2_simple.py:38-66 shape) ----
        m = ref_model(ref, "stage2", perturb=7)
        ids, lengths, dur = H.small_inputs(2, 15, 256, seed=12)
        out = m(ids, lengths, target_durations=dur, max_target_length=60)
        np.savez(OUT / "stage2_biased_fwd.npz", ids=ids.numpy(), lengths=lengths.numpy(), dur=dur.numpy(), **npify(out))
        # predicted-duration path (no target durations, no max length)
        out = m(ids, lengths)
        np.savez(OUT / "stage2_biased_pred.npz", ids=ids.numpy(), lengths=lengths.numpy(), **npify(out))

        # ---- C1: stage1 "Hello world" through inference() ----
        m = ref_model(ref, "stage1")
        ids = torch.full((1, 256), 39, dtype=torch.long)
        ids[0, : len(H.HELLO_WORLD_IDS)] = torch.tensor(H.HELLO_WORLD_IDS)
        lengths = torch.tensor([9])
        mel1, audio1 = m.inference(ids, lengths, 1.0)
        fwd = m(ids, lengths)
        mel4, audio4 = m.inference(ids, lengths, 4.0)
        np.savez(OUT / "c1_hello_world.npz", ids=ids.numpy(), lengths=lengths.numpy(),
                 duration_pred=fwd["duration_pred"].numpy(), encoder_output=fwd["encoder_output"].numpy(),
                 mel_scale1=mel1.numpy(), audio_scale1=audio1.numpy(),
                 mel_scale4=mel4.numpy(), audio_scale4=audio4.numpy())

        # ---- C2: stage1 B=16 S=64 (full tensors for two utterances + per-utterance sums) ----
        ids, lengths, dur = H.c2_inputs()
        out = m(ids, lengths, target_durations=dur)
        keep = [0, 5]
        np.savez(OUT / "c2_stage1.npz",
                 encoder_output=out["encoder_output"].numpy(), duration_pred=out["duration_pred"].numpy(),
                 frames=dur.long().sum(1).numpy().astype(np.int32), T=np.int32(out["mel_output"].shape[1]),
                 keep=np.array(keep), mel_keep=out["mel_output"][keep].numpy(),
                 audio_keep=out["audio_output"][keep].numpy(),
                 mel_sum=out["mel_output"].double().sum((1, 2)).numpy(),
                 mel_abs=out["mel_output"].double().abs().sum((1, 2)).numpy(),
                 audio_sum=out["audio_output"].double().sum((1, 2)).numpy(),
                 audio_abs=out["audio_output"].double().abs().sum((1, 2)).numpy())

        # ---- length regulator edge cases through the reference's own loop ----
        lr = ref.LengthRegulator()
        d = H.lr_edge_durations()
        B, S = d.shape
        enc = (torch.arange(S, dtype=torch.float32) + 1.0)[None, :, None].expand(B, S, 4).contiguous()
        cases = {}
        for name, ml in (("none", None), ("trunc20", 20), ("pad90", 90)):
            o = lr(enc, d, ml)
            cases[f"index_{name}"] = (o[:, :, 0].round().long() - 1).numpy().astype(np.int32)  # -1 = zero row
        np.savez(OUT / "length_regulator_edges.npz", dur=d.numpy(), **cases)

    (OUT / "state_digests.json").write_text(json.dumps(digests, indent=1, sort_keys=True) + "\n")
    tot = sum(p.stat().st_size for p in OUT.glob("*.npz"))
    print(f"wrote {len(list(OUT.glob('*.npz')))} fixtures, {tot/1e6:.2f} MB")


if __name__ == "__main__":
    main()
