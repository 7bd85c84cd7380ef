"""CUDA-graph replay of a fixed-shape synthesis step — the small-batch path (SURVEY.md §7 "hard parts": a 2 MB model at
batch 1 is launch-bound; BASELINE.json configs[0], [4]).

``GraphedStep(fn, example)`` warms `fn` up (workspaces and weight images get allocated and packed outside the capture),
captures one call on a side stream with the status read deferred, and replays it on new inputs of the same shape: one graph
launch instead of ~30 kernel launches with their host-side tensor-map encodes. The status word is read after every replay
(`check=True`) exactly like the eager path: a range violation raises ``Fp16RangeError`` — capture again under
``precision("tf32")`` in that case.
"""
from __future__ import annotations

from typing import Callable

import torch

from models import _native as nat


class GraphedStep:
    def __init__(self, fn: Callable[[torch.Tensor], torch.Tensor], example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise ValueError("GraphedStep needs a CUDA example input")
        self.device = example.device
        self.static_in = example.clone()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):          # eager: allocates workspaces, packs weights, validates the status word
                fn(self.static_in)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        before = nat.launch_count()
        with nat.deferred_status(), torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = fn(self.static_in)
        self.launches_captured = nat.launch_count() - before
        nat.read_status(self.device)                 # the capture itself executes nothing; start from a clear word

    def __call__(self, x: torch.Tensor, check: bool = True) -> torch.Tensor:
        """Replay on `x` (same shape/dtype as the example). Returns the graph's static output buffer: copy it if it has to
        survive the next call."""
        if x.shape != self.static_in.shape or x.dtype != self.static_in.dtype:
            raise ValueError(f"GraphedStep captured {tuple(self.static_in.shape)}, got {tuple(x.shape)}")
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        if check:
            nat.check_status(self.device, "graph replay")
        return self.static_out
