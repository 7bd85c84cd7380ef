"""Bring-up helper (GPU box, torchrun): host -> host time of the C3 step for several HostPipeline chunk layouts with ALL ranks of
the node copying at once (the PCIe / host-memory side is shared: what is best for one GPU is not best for eight).
usage: python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/e2e_sweep_dist.py"""
import os
import statistics
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'm2-tts_b200' / 'src'))
import torch
import torch.distributed as dist
from models import _native as nat
from models.tts_model import M2TTSModel
from models.stage_configs import STAGE_KWARGS
from utils.host_pipeline import HostPipeline
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(1234)
m = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)
x_host = torch.randn(64, 3446, 96).pin_memory()
out = torch.empty(64, 1, 3446 * 64).pin_memory()
step = lambda x: m.vocoder(m.decoder(x).transpose(1, 2))


def barrier():
    if world > 1:
        dist.barrier()


cands = ([5, 27, 27, 5], [10, 22, 27, 5], [10, 27, 22, 5], [16, 21, 22, 5], [10, 16, 27, 11], [16, 16, 27, 5], [10, 21, 28, 5], [5, 16, 21, 17, 5],
         [10, 16, 16, 17, 5], [22, 21, 21], [16, 16, 16, 16], [10, 27, 27], [16, 27, 21], [21, 27, 16], [10, 54], [16, 43, 5], [10, 49, 5], [21, 38, 5])
res = {tuple(c): [] for c in cands}
pipes = {tuple(c): HostPipeline(dev, sizes=c) for c in cands}
with nat.deferred_status():
    for _ in range(3):
        step(x_host.to(dev))
    for rnd in range(4):
        for c in cands:
            pipe = pipes[tuple(c)]
            pipe.run(step, x_host, out); pipe.synchronize()
            torch.cuda.synchronize(); barrier()
            t0 = time.perf_counter()
            for _ in range(4):
                pipe.run(step, x_host, out); pipe.synchronize()
            torch.cuda.synchronize(); barrier()
            res[tuple(c)].append((time.perf_counter() - t0) / 4 * 1e3)
nat.check_status(dev, "e2e sweep")
if rank == 0:
    for c in sorted(cands, key=lambda c: statistics.median(res[tuple(c)])):
        v = res[tuple(c)]
        print(f"world {world} sizes {c}: median {statistics.median(v):.3f} min {min(v):.3f} max {max(v):.3f}", flush=True)
if world > 1:
    dist.destroy_process_group()
