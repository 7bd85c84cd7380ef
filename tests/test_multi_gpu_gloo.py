"""world_size-2 `gloo` test of the N>1 host logic on CPU: contiguous batch sharding, the shared
max_target_length all-reduce, and the post-path all-gather reproduce the single-process batch.
(The CUDA kernels need a GPU; here the mirror model runs its train-mode torch formulation with
dropout 0, which is the same math — see test_host_logic.py.)"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from utils.shard import synthesize_sharded
        torch.set_num_threads(1)
        m = H.product_model("tiny", perturb=2, dropout=0.0).train()
        for p in m.duration_predictor.modules():   # eval BatchNorm statistics so utterances stay independent
            if isinstance(p, torch.nn.BatchNorm1d):
                p.eval()
        ids, lengths, dur = H.small_inputs(5, 12, 256, seed=3)   # 5 utterances -> uneven shards 3 + 2
        with torch.no_grad():
            got = synthesize_sharded(m, ids, lengths, dur, None)
            full = m(ids, lengths, target_durations=dur)
        ok = (got["max_target_length"] == full["mel_output"].shape[1]
              and torch.equal(got["mel_output"], full["mel_output"]))
        # fewer utterances than ranks: rank 1 owns an empty block and must still enter the collectives
        with torch.no_grad():
            one = synthesize_sharded(m, ids[:1], lengths[:1], dur[:1], None)
            full1 = m(ids[:1], lengths[:1], target_durations=dur[:1])
        ok = ok and one["mel_output"].shape == full1["mel_output"].shape and torch.equal(one["mel_output"], full1["mel_output"])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_sharded_synthesis_equals_full_batch_world2():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
