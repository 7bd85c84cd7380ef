#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 synthesis path (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "C3"): stage2_quality VAE-less transformer mel decoder +
HiFi-GAN-style vocoder, 64 utterances x 3446 frames (10.0 s at 22.05 kHz) PER GPU (weak scaling:
C4 = 8 x 64 = 512 utterances at N=8), seeded random-init weights, synthetic `regulated_output`.
One step = decoder + vocoder over the batch. Metric = audio-seconds synthesised per second.

  value : inputs resident in HBM, CUDA events per step, max over ranks
  e2e   : same step through the module API from PINNED HOST input to PINNED HOST waveform
          (H2D + D2H inside the timed region)
  roofline     : dominant kernel, timed live by the library's per-launch CUDA events
  cpu_baseline : the oracle port (torch CPU, all host threads) on a bounded sample (rank 0, N=1)
  --impl reference : the reference's CPU implementation of the path (the oracle port: the
          reference is pure Python and cannot travel to the GPU box), same metric/config.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for _p in (str(ROOT), str(ROOT / "m2-tts_b200" / "src")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

STAGE = "stage2"
BATCH = 64          # utterances per GPU
FRAMES = 3446       # 10.002 s at 64 samples/frame, 22 050 Hz
HIDDEN, MEL, LAYERS, VOC = 96, 80, 3, 256
SAMPLES_PER_FRAME, SAMPLE_RATE = 64, 22050
METRIC = "audio-sec/sec (synthesis RTF^-1)"
UNIT = "audio-s/s"


def audio_seconds(n_utt: int, frames: int = FRAMES) -> float:
    return n_utt * frames * SAMPLES_PER_FRAME / SAMPLE_RATE


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d.get("hbm_gbs"), "bf16_tflops": d.get("bf16_tflops"),
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml, else nvidia-smi)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self._stop = threading.Event()
        self._thr = None
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical(index))
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    @staticmethod
    def _physical(index: int) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[index])
            except Exception:
                return index
        return index

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        legacy = {"hw_slowdown": "nvmlClocksThrottleReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksThrottleReasonHwThermalSlowdown",
                  "sw_thermal_slowdown": "nvmlClocksThrottleReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksThrottleReasonSwPowerCap"}
        while not self._stop.is_set():
            try:
                if nv is not None:
                    self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for k in names:
                        bit = getattr(nv, names[k], None) or getattr(nv, legacy[k], None)
                        if bit and (mask & bit):
                            self.reasons.add(k)
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm,"
                                          "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                                          "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                                          "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                    f = [x.strip() for x in out.strip().split(",")]
                    self.samples.append(int(f[0])); self.max_mhz = int(f[1])
                    for k, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        self._thr = threading.Thread(target=self._loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        med = statistics.median(self.samples) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------------
def cpu_port_step(sd, x, oracle):
    mel = oracle.mel_decoder(sd, x, 2)
    return oracle.vocoder(sd, mel.transpose(1, 2))


def time_cpu_port(n_utt: int, reps: int, warm: int, seed: int = 0):
    """The reference's CPU path (oracle port: torch CPU fp32, scores materialised exactly like
    components.py:75-87) on `n_utt` utterances of the C3 workload. Returns (audio-s/s list, threads)."""
    from oracle import m2tts_oracle as oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    from models.tts_model import M2TTSModel
    sd = M2TTSModel(**oracle.STAGE_KWARGS[STAGE]).eval().state_dict()
    x = torch.randn(n_utt, FRAMES, HIDDEN, generator=torch.Generator().manual_seed(seed))
    vals = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            cpu_port_step(sd, x, oracle)
            dt = time.perf_counter() - t0
            if i >= warm:
                vals.append(audio_seconds(n_utt) / dt)
    return vals, threads


def run_reference(args, rank: int):
    if rank != 0:
        return
    n_utt = 8
    vals, threads = time_cpu_port(n_utt, reps=args.steps, warm=args.warmup)
    value = audio_seconds(n_utt) * len(vals) / sum(audio_seconds(n_utt) / v for v in vals)
    sample = f"{n_utt} of {BATCH} utterances x {FRAMES} frames per step (decoder+vocoder), torch CPU fp32"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * audio_seconds(n_utt) / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int):
    return {"workload": "C3 stage2_quality decoder+vocoder, 64 utterances x 3446 frames (10 s @ 22.05 kHz) per GPU"
                        + (f"; C4-style batch sharding, {64 * n_gpus} utterances total" if n_gpus > 1 else ""),
            "utterances_per_gpu": BATCH, "frames": FRAMES, "hidden": HIDDEN, "mel": MEL, "decoder_layers": LAYERS,
            "vocoder_channels": VOC, "weights": "random-init seed 1234", "parallelism": f"dp{n_gpus} (batch sharding, no data-path collective)",
            "l2": "256 MiB buffer rewritten between timed steps; a step streams >3 GB of intermediates (L2 = 126 MB)"}


# --------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    from models import _native as nat
    from models.stage_configs import STAGE_KWARGS
    from models.tts_model import M2TTSModel

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "NONE"      # keep stdout to the one JSON line (from level VERSION up NCCL prints a banner there)
        dist.init_process_group("nccl", device_id=dev)

    torch.manual_seed(1234)
    model = M2TTSModel(**STAGE_KWARGS[STAGE]).eval().to(dev)
    x_host = torch.randn(BATCH, FRAMES, HIDDEN, generator=torch.Generator().manual_seed(rank)).pin_memory()
    x_dev = x_host.to(dev)
    audio_host = torch.empty((BATCH, 1, FRAMES * SAMPLES_PER_FRAME), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(x):
        mel = model.decoder(x)
        return model.vocoder(mel.transpose(1, 2))

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step(x_dev)
    torch.cuda.synchronize(dev)

    # fp32 FFMA peak of this GPU right now (no such number in MEASURED_PEAKS.json)
    sink = torch.zeros(4, device=dev)
    import ctypes as C
    flops = C.c_double(0.0)
    nat.check(nat.lib().m2tts_ffma_probe(sink.data_ptr(), 4096, C.byref(flops), nat.stream_handle(dev)), "probe")
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.check(nat.lib().m2tts_ffma_probe(sink.data_ptr(), 65536, C.byref(flops), nat.stream_handle(dev)), "probe")
    e1.record()
    torch.cuda.synchronize(dev)
    ffma_peak_tflops = flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12

    # ---- timed region: K steps, device-resident input ----
    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.start()
    nat.stage_timing_enable(True)
    launches0 = nat.launch_count()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)            # evict L2 between steps (outside the event pair)
        starts[i].record()
        step(x_dev)
        ends[i].record()
    barrier()
    launches = nat.launch_count() - launches0
    nat.stage_timing_enable(False)
    clocks = sampler.stop()
    stage_ms = nat.stage_timing_read()
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))

    # ---- e2e: pinned host in -> pinned host out, copies inside the timed region. The caller-facing helper
    # (utils.host_pipeline.HostPipeline) chunks the utterance batch so the PCIe copies overlap compute. ----
    from utils.host_pipeline import HostPipeline
    pipe = HostPipeline(dev, n_chunks=args.e2e_chunks, edge=args.e2e_edge)
    for _ in range(2):
        pipe.run(step, x_host, audio_host)
        pipe.synchronize()
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe.run(step, x_host, audio_host)
        pipe.synchronize()               # the caller needs the waveform on the host
    barrier()
    e2e_s = time.perf_counter() - t0

    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        n_utt_total = BATCH * world
        aud = audio_seconds(n_utt_total)
        value = aud * args.steps / (total_ms * 1e-3)
        e2e_value = aud * args.steps / (e2e_ms * 1e-3)
        peaks = measured_peaks()
        # dominant stage by summed device time; algorithmic FLOPs per step from SURVEY §8d's model
        fl = stage_flops_per_step()
        dom = max(stage_ms.items(), key=lambda kv: kv[1][0]) if stage_ms else ("none", (0.0, 1))
        dom_ms_per_step = dom[1][0] / args.steps
        tensor_stages = {"attention": "tcgen05 kind::f16, 16-bit split (fp16 hi/lo, 3 product terms, fp32 accumulate): issued MMA FLOPs = 3x algorithmic",
                         "voc_in": "tcgen05 3xTF32 tap-GEMM (after a strided -> channel-first copy of the mel), writes fp16 hi/lo planes channel-last",
                         "voc_up": "stages 0-1 transposed convs: polyphase channel-last tcgen05 kind::f16 16-bit split kernel with TMA stores (voc_up_h.cu)",
                         "voc_res1": "stage 0 conv1 (C=128, channel-last 16-bit split conv kernel) + the whole stage-1 ResBlock (C=64, one fused 16-bit split kernel)",
                         "voc_res2": "stage 0 conv2 + residual (C=128, channel-last 16-bit split conv kernel, writes fp32 channel-first)",
                         "voc_fused": "stages 2-3: upsample + ResBlock (+ output conv + tanh) fused, channel-last tcgen05 kind::f16 16-bit split"}
        roof = {"kernel": dom[0], "bound": "tensor", "unit": "TFLOP/s", "stage_ms_per_step": dom_ms_per_step,
                "launches_per_step": dom[1][1] // args.steps, "launch_ms": dom[1][0] / max(dom[1][1], 1),
                "peak": peaks["bf16_tflops_sustained"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                "traffic": read_traffic(dom[0]), "fp32_ffma_peak_tflops": ffma_peak_tflops,
                "path": tensor_stages.get(dom[0], "fp32 FFMA")}
        if dom[0] in fl and dom_ms_per_step > 0:
            ach = fl[dom[0]] / (dom_ms_per_step * 1e-3) / 1e12
            roof.update(achieved=ach, frac=ach / peaks["bf16_tflops_sustained"],
                        algorithmic_flops_per_step=fl[dom[0]],
                        note="achieved = algorithmic (useful fp32-equivalent) FLOPs / device time of the stage; an fp32-faithful "
                             "split-precision kernel issues 3 products per algorithmic one: ceiling 1/3 of the bf16 peak with fp16 "
                             "halves (attention, linear layers, ResBlocks, narrow stages), 1/6 with TF32 halves (input conv)")
        else:
            roof.update(achieved=None, frac=None)
        all_stage_tflops = {k: round(fl[k] / (v[0] / args.steps * 1e-3) / 1e12, 2) for k, v in stage_ms.items()
                            if k in fl and v[0] > 0}
        # every stage as a fraction of ITS roofline: GEMM-shaped stages against the measured bf16 tensor peak (an
        # fp32-faithful 3xTF32 kernel tops out at 1/6 of it), streaming stages against the measured HBM copy bandwidth
        by = stage_bytes_per_step()
        stage_roofline = {}
        for k, v in stage_ms.items():
            sec = v[0] / args.steps * 1e-3
            if sec <= 0:
                continue
            if k in by:
                ach = by[k] / sec / 1e9
                stage_roofline[k] = {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": round(ach / peaks["hbm_gbs"], 4)}
            elif k in fl:
                ach = fl[k] / sec / 1e12
                stage_roofline[k] = {"bound": "tensor" if k in tensor_stages or k in TENSOR_LINEAR else "ffma",
                                     "achieved": round(ach, 2), "unit": "TFLOP/s",
                                     "peak": peaks["bf16_tflops_sustained"] if (k in tensor_stages or k in TENSOR_LINEAR) else round(ffma_peak_tflops, 1),
                                     "frac": round(ach / (peaks["bf16_tflops_sustained"] if (k in tensor_stages or k in TENSOR_LINEAR) else ffma_peak_tflops), 4)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": audio_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                        "api": f"utils.host_pipeline.HostPipeline(n_chunks={args.e2e_chunks}, edge={args.e2e_edge}).run(decoder+vocoder, pinned host in, pinned host out)"},
                "roofline": roof, "stage_roofline": stage_roofline, "stage_tflops": all_stage_tflops,
                "stage_ms_per_step": {k: round(v[0] / args.steps, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1][0])},
                "x_realtime_per_gpu": value / world}
        if world == 1 and not args.skip_cpu_baseline:
            n_s = 8
            vals, threads = time_cpu_port(n_s, reps=2, warm=0)
            vals = [max(vals)]
            line["cpu_baseline"] = {"value": vals[0], "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n_s} of {BATCH} utterances x {FRAMES} frames, best of 2 passes, torch CPU fp32"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def stage_flops_per_step():
    """Algorithmic FLOPs of one C3 step per library stage (SURVEY.md §8d formulas)."""
    B, T, H, M, C, L = BATCH, FRAMES, HIDDEN, MEL, VOC, LAYERS
    F = 2 * H
    rows = B * T
    fl = {"attention": L * 4 * T * T * H * B, "ln_qkv": L * 2 * H * 3 * H * rows, "out_proj": L * 2 * H * H * rows,
          "ffn1": L * 2 * H * F * rows, "ffn2": L * 2 * H * F * rows, "ln_proj": 2 * H * M * rows,
          "voc_in": 2 * 3 * M * C * rows, "voc_up": 0, "voc_res1": 0, "voc_res2": 0}
    c_in, Lc = C, T
    fl["voc_fused"] = 0
    for j, r in enumerate((4, 4, 2, 2)):
        c, Lc = c_in // 2, Lc * r
        up, res = 2 * 2 * c_in * c * Lc * B, 2 * 3 * c * c * Lc * B       # two taps per output sample; one k=3 conv
        if j < 2:      # wide stages: upsampling tap-GEMM, then two conv launches (C = 128) or one fused ResBlock launch (C = 64,
            fl["voc_up"] += up          # accounted under voc_res1)
            fl["voc_res1"] += res if j == 0 else 2 * res
            fl["voc_res2"] += res if j == 0 else 0
        else:          # narrow stages: one fused kernel each (the last one includes the output conv)
            fl["voc_fused"] += up + 2 * res
        c_in = c
    fl["voc_fused"] += 2 * 3 * c_in * Lc * B
    return fl


TENSOR_LINEAR = {"ln_qkv", "out_proj", "ffn1", "ffn2", "ln_proj"}      # tcgen05 16-bit split linear layers (lin_h.cu)


def stage_bytes_per_step():
    """Algorithmic HBM bytes of one C3 step for the streaming (bandwidth-bound) stages: every operand once."""
    rows = BATCH * FRAMES
    F = 2 * HIDDEN
    per = rows * 4
    return {"layernorm": (2 * LAYERS) * 2 * rows * HIDDEN * 4 + 2 * rows * HIDDEN * 4,   # 2 per layer + the final one: x in, LN(x) out
            # K = 96 linear layers sit at the 3xTF32 ridge (36 FLOP/B): reported against HBM, every operand once in fp32
            "ln_qkv": LAYERS * per * (HIDDEN + 3 * HIDDEN), "out_proj": LAYERS * per * 3 * HIDDEN,
            "ffn1": LAYERS * per * (HIDDEN + F), "ffn2": LAYERS * per * (F + 2 * HIDDEN), "ln_proj": per * (HIDDEN + MEL),
            "pack": 2 * 4 * sum(x * 4 for x in (3 * HIDDEN * HIDDEN, HIDDEN * HIDDEN, 2 * HIDDEN * HIDDEN, 2 * HIDDEN * HIDDEN)) * LAYERS}


def read_traffic(kernel: str):
    p = ROOT / "profiles" / "dram_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs: omit the CPU leg")
    ap.add_argument("--e2e-chunks", type=int, default=3, help="utterance chunks of the host-to-host pipeline (1 = no overlap)")
    ap.add_argument("--e2e-edge", type=float, default=1.0, help="relative size of the first and last chunk (their copies are the unhidden ones)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU path)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
