#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 synthesis path (contract: see DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c1|c2|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Headline workload (BASELINE.json configs[2], "C3"): stage2_quality transformer mel decoder + HiFi-GAN-style vocoder,
64 utterances x 3446 frames (10.0 s at 22.05 kHz) PER GPU (weak scaling: C4 = 8 x 64 = 512 utterances at N=8), seeded
random-init weights, synthetic `regulated_output`. One step = decoder + vocoder over the batch. Metric = audio-seconds
synthesised per second.

  value        : inputs resident in HBM, CUDA events per step, max over ranks; nothing but the step inside the timed region
                 (no stage timers, status word read once after the region)
  e2e          : same step through the module API from PINNED HOST input to PINNED HOST waveform (H2D + D2H inside the
                 timed region, utils.host_pipeline.HostPipeline)
  e2e_from_ids : the whole model from host phoneme ids (S = 256, 3446 frames) to a host waveform (SURVEY §8d C3 e2e variant)
  parity       : utterance 0 of the timed step's own output against the CPU oracle; the run fails above 1e-4
  roofline     : dominant kernel, timed by the library's per-launch CUDA events in a SEPARATE instrumented pass
  cpu_baseline : the reference's CPU path (oracle/_ref = the unmodified reference modules when vendored by
                 oracle/build_ref.py, else the oracle port) on a bounded sample (rank 0, N=1)
  gather       : at N > 1, the all-gather of the sharded waveforms after the path (utils.shard.gather_batch), timed apart
  --impl reference : the reference's CPU implementation of the path, same metric/config, all host threads
  --config c1|c2|c5 : secondary lines for BASELINE.json configs[0], [1], [4] (headline stays C3)
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
PARITY_TOL = 1e-4


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
            self._stop.wait(0.02)

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
# CPU arm: the reference's own modules (oracle/_ref, vendored by oracle/build_ref.py where /root/reference exists) or,
# without them, the oracle port. Only this leg and the parity check touch oracle/.
class CpuPath:
    def __init__(self, stage: str = STAGE):
        from oracle import build_ref
        from oracle import m2tts_oracle as oracle
        self.oracle = oracle
        self.threads = os.cpu_count() or 1
        torch.set_num_threads(self.threads)
        kw = oracle.STAGE_KWARGS[stage]
        self.heads = kw["num_heads"]
        if build_ref.available():
            ref = build_ref.load()
            torch.manual_seed(1234)
            self.model = ref.M2TTSModel(**kw).eval()      # unmodified reference code, its own seeded init
            self.kind = "reference"
            self.sd = None
        else:
            from models.tts_model import M2TTSModel
            torch.manual_seed(1234)
            self.sd = M2TTSModel(**kw).eval().state_dict()
            self.model = None
            self.kind = "port"

    @torch.no_grad()
    def decoder_vocoder(self, x: torch.Tensor, chunk: int = 8) -> torch.Tensor:
        """reference tts_model.py:211-228, 279-297 on utterance chunks (utterances are independent in eval mode; chunking
        bounds the materialised [chunk, 2, T, T] score tensors of components.py:75-87 to 0.76 GB per layer)."""
        outs = []
        for lo in range(0, x.shape[0], chunk):
            xs = x[lo:lo + chunk]
            if self.model is not None:
                mel = self.model.decoder(xs)
                outs.append(self.model.vocoder(mel.transpose(1, 2)))
            else:
                mel = self.oracle.mel_decoder(self.sd, xs, self.heads)
                outs.append(self.oracle.vocoder(self.sd, mel.transpose(1, 2)))
        return torch.cat(outs, 0)

    def time_c3(self, n_utt: int, reps: int, warm: int, seed: int = 0):
        x = torch.randn(n_utt, FRAMES, HIDDEN, generator=torch.Generator().manual_seed(seed))
        vals = []
        for i in range(warm + reps):
            t0 = time.perf_counter()
            self.decoder_vocoder(x)
            dt = time.perf_counter() - t0
            if i >= warm:
                vals.append(audio_seconds(n_utt) / dt)
        return vals


def run_reference(args, rank: int):
    """`--impl reference`: the reference's CPU implementation of the C3 step on all host threads. A step processes as many
    of the 64 utterances as keep the whole run within ~2 minutes (a power of two >= 8; the sample is stated in the line)."""
    if rank != 0:
        return
    cpu = CpuPath()
    x8 = torch.randn(8, FRAMES, HIDDEN, generator=torch.Generator().manual_seed(0))
    cpu.decoder_vocoder(x8[:2])                                   # page in
    t0 = time.perf_counter()
    cpu.decoder_vocoder(x8)
    t8 = time.perf_counter() - t0
    total_steps = max(args.steps + args.warmup, 1)
    n_utt = 8
    while n_utt < BATCH and (2 * n_utt / 8) * t8 * total_steps <= 120.0:
        n_utt *= 2
    vals = cpu.time_c3(n_utt, reps=args.steps, warm=args.warmup)
    value = audio_seconds(n_utt) * len(vals) / sum(audio_seconds(n_utt) / v for v in vals)
    sample = (f"{n_utt} of {BATCH} utterances x {FRAMES} frames per step (decoder+vocoder, same per-utterance work; bounded so that "
              f"{total_steps} steps end within ~2 min), torch CPU fp32, {cpu.threads} threads")
    cfg = workload_config(args.gpus)
    cfg["workload"] += f" [reference arm: {n_utt} of the 64 utterances per step]"
    cfg["utterances_per_step_reference_arm"] = n_utt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * audio_seconds(n_utt) / value,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "gpu_launches": 0,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(n_gpus: int):
    return {"workload": "C3 stage2_quality decoder+vocoder, 64 utterances x 3446 frames (10 s @ 22.05 kHz) per GPU"
                        + (f"; C4-style batch sharding, {64 * n_gpus} utterances total" if n_gpus > 1 else ""),
            "utterances_per_gpu": BATCH, "frames": FRAMES, "hidden": HIDDEN, "mel": MEL, "decoder_layers": LAYERS,
            "vocoder_channels": VOC, "weights": "random-init seed 1234", "parallelism": f"dp{n_gpus} (batch sharding, no data-path collective)",
            "l2": "256 MiB buffer rewritten between timed steps; a step streams >3 GB of intermediates (L2 = 126 MB)"}


# --------------------------------------------------------------------------------------------
def setup_dist(world: int, dev):
    if world <= 1:
        return None
    import torch.distributed as dist
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
        os.environ["NCCL_DEBUG"] = "NONE"      # keep stdout to the one JSON line (from level VERSION up NCCL prints a banner there)
    dist.init_process_group("nccl", device_id=dev)
    return dist


def ffma_peak(nat, dev) -> float:
    """fp32 FFMA peak of this GPU right now (no such number in MEASURED_PEAKS.json)."""
    import ctypes as C
    sink = torch.zeros(4, device=dev)
    flops = C.c_double(0.0)
    nat.check(nat.lib().m2tts_ffma_probe(sink.data_ptr(), 4096, C.byref(flops), nat.stream_handle(dev)), "probe")
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nat.check(nat.lib().m2tts_ffma_probe(sink.data_ptr(), 65536, C.byref(flops), nat.stream_handle(dev)), "probe")
    e1.record()
    torch.cuda.synchronize(dev)
    return flops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12


def run_b200(args, rank: int, world: int, local_rank: int):
    from models import _native as nat
    from models.stage_configs import STAGE_KWARGS
    from models.tts_model import M2TTSModel
    from utils.host_pipeline import HostPipeline

    from utils.device import bind_host_to_gpu

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    all_cores = os.sched_getaffinity(0)
    host_binding = ({"bound": False, "why": "--no-bind"} if args.no_bind else
                    bind_host_to_gpu(dev, local_rank, world))      # before the pinned buffers are allocated (first touch)
    dist = setup_dist(world, dev)

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

    warm = max(args.warmup, 3)
    for _ in range(warm):                 # eager: packs the weight images, validates the status word per stage
        step(x_dev)
    torch.cuda.synchronize(dev)
    ffma_peak_tflops = ffma_peak(nat, dev)

    # ---- timed region: K steps, device-resident input, nothing else on the stream ----
    sampler = ClockSampler(local_rank)
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    barrier()
    sampler.start()
    launches0 = nat.launch_count()
    with nat.deferred_status():
        for i in range(args.steps):
            flush.fill_(i & 0xFF)            # evict L2 between steps (outside the event pair)
            starts[i].record()
            audio_dev = step(x_dev)
            ends[i].record()
    barrier()
    launches = nat.launch_count() - launches0
    clocks = sampler.stop()
    nat.check_status(dev, "timed region")     # raises if any step left the fp16 range: the number would be invalid
    total_ms = sum(s.elapsed_time(e) for s, e in zip(starts, ends))
    mel_dev0 = model.decoder(x_dev[:1])       # utterance 0 again, for the parity check below (bit-identical rows)

    # ---- instrumented pass (not part of any headline number): per-kernel CUDA events keyed by stage ----
    prof_steps = 3
    nat.stage_timing_enable(True)
    with nat.deferred_status():
        for i in range(prof_steps):
            flush.fill_(i & 0xFF)
            step(x_dev)
    torch.cuda.synchronize(dev)
    nat.stage_timing_enable(False)
    stage_ms = nat.stage_timing_read()
    nat.check_status(dev, "instrumented pass")

    # ---- e2e: pinned host in -> pinned host out, copies inside the timed region. The caller-facing helper
    # (utils.host_pipeline.HostPipeline) chunks the utterance batch so the PCIe copies overlap compute. ----
    if args.e2e_chunks > 0:
        pipe = HostPipeline(dev, n_chunks=args.e2e_chunks, edge=args.e2e_edge)
    else:
        # default: the pipeline measures a handful of chunk layouts on this step and keeps the fastest (what the copy streams get
        # depends on how many ranks of the host copy at once; every rank tunes on its own, at the same point of the program: no
        # collective — utils.host_pipeline.HostPipeline.autotune)
        pipe = HostPipeline(dev, n_chunks=3)
        with nat.deferred_status():
            pipe.autotune(step, x_host, audio_host, FRAMES, model.decoder.layers[0].self_attn.num_heads)
        nat.check_status(dev, "e2e chunk layout autotune")
    pipe_desc = (f"sizes={pipe.sizes}" if pipe.sizes is not None else f"n_chunks={args.e2e_chunks}, edge={args.e2e_edge}")
    with nat.deferred_status():
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
    nat.check_status(dev, "e2e region")
    e2e_first = audio_host[0].clone()

    # ---- e2e from phoneme ids: the whole model (encoder, duration predictor, length regulator, decoder, vocoder) ----
    S = 256
    g = torch.Generator().manual_seed(100 + rank)
    ids_host = torch.randint(0, 256, (BATCH, S), generator=g).pin_memory()
    len_host = torch.full((BATCH,), S, dtype=torch.int64).pin_memory()
    dur = torch.full((BATCH, S), 13.0)
    for b in range(BATCH):
        dur[b, torch.randperm(S, generator=g)[:FRAMES - 13 * S]] = 14.0
    dur_host = (dur + 0.5).pin_memory()

    from utils.host_pipeline import synthesize_to_host

    # the ids path copies only waveforms: its own pipeline, its own chunk layout (tuned on a resident input of the regulated shape)
    if args.e2e_chunks > 0:
        pipe_ids = pipe
    else:
        pipe_ids = HostPipeline(dev, n_chunks=3)
        with nat.deferred_status():
            pipe_ids.autotune(step, x_dev, audio_host, FRAMES, model.decoder.layers[0].self_attn.num_heads)
        nat.check_status(dev, "e2e-from-ids chunk layout autotune")
    ids_desc = (f"sizes={pipe_ids.sizes}" if pipe_ids.sizes is not None else pipe_desc)

    def ids_step():
        synthesize_to_host(model, ids_host, len_host, dur_host, FRAMES, audio_host, pipe_ids)
        pipe_ids.d2h.synchronize()

    with nat.deferred_status():
        for _ in range(2):
            ids_step()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            ids_step()
        barrier()
        ids_s = time.perf_counter() - t0
    nat.check_status(dev, "e2e-from-ids region")

    # ---- N > 1: all-gather of the sharded waveforms after the path, timed apart from the path ----
    gather_ms = None
    gather_ok = None
    if dist is not None:
        from utils.shard import gather_batch, shard_bounds
        n_total = BATCH * world
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = gather_batch(audio_dev, n_total)
        g1.record()
        barrier()
        lo, hi = shard_bounds(n_total, rank, world)
        gather_ok = bool(full.shape[0] == n_total and torch.equal(full[lo:hi], audio_dev))
        gt = torch.tensor([g0.elapsed_time(g1), 0.0 if gather_ok else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        gather_ms, gather_ok = float(gt[0]), float(gt[1]) == 0.0
        del full

    t = torch.tensor([total_ms, e2e_s * 1e3, ids_s * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, ids_ms = float(t[0]), float(t[1]), float(t[2])

    rc = 0
    if rank == 0:
        # ---- parity of what was just timed: utterance 0 against the CPU oracle (decoder 0.5 s, vocoder 0.2 s of CPU) ----
        from oracle import m2tts_oracle as oracle
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        with torch.no_grad():
            want_mel = oracle.mel_decoder(sd, x_host[:1].clone(), 2)
            want_audio = oracle.vocoder(sd, want_mel.transpose(1, 2))
        got_audio = audio_dev[:1].cpu()
        parity = {"max_abs_mel": float((mel_dev0.cpu() - want_mel).abs().max()),
                  "max_abs_audio": float((got_audio - want_audio).abs().max()),
                  "e2e_equals_device_path": bool(torch.equal(e2e_first, got_audio[0])),
                  "tolerance": PARITY_TOL,
                  "checked": "utterance 0 of rank 0's timed batch (decoder -> mel, vocoder -> waveform) vs oracle/m2tts_oracle.py on the CPU"}
        parity["ok"] = bool(parity["max_abs_mel"] <= PARITY_TOL and parity["max_abs_audio"] <= PARITY_TOL and parity["e2e_equals_device_path"])
        if not parity["ok"]:
            rc = 3

        n_utt_total = BATCH * world
        aud = audio_seconds(n_utt_total)
        value = aud * args.steps / (total_ms * 1e-3)
        e2e_value = aud * args.steps / (e2e_ms * 1e-3)
        ids_value = aud * args.steps / (ids_ms * 1e-3)
        peaks = measured_peaks()
        # dominant stage by summed device time; algorithmic FLOPs per step from SURVEY §8d's model
        fl = stage_flops_per_step()
        dom = max(stage_ms.items(), key=lambda kv: kv[1][0]) if stage_ms else ("none", (0.0, 1))
        dom_ms_per_step = dom[1][0] / prof_steps
        tensor_stages = {"attention": "tcgen05 kind::f16, 16-bit split (fp16 hi/lo, 3 product terms, fp32 accumulate): issued MMA FLOPs = 3x algorithmic",
                         "voc_in": "channel-last 16-bit split conv kernel (CI = 80 zero-padded to 128 by TMA) after a flat fp32 -> fp16 hi/lo split of the mel",
                         "voc_up": "stages 0-1 transposed convs: polyphase channel-last tcgen05 kind::f16 16-bit split kernel with TMA stores (voc_up_h.cu)",
                         "voc_res1": "stage 0 conv1 (C=128, channel-last 16-bit split conv kernel) + the whole stage-1 ResBlock (C=64, one fused 16-bit split kernel)",
                         "voc_res2": "stage 0 conv2 + residual (C=128, channel-last 16-bit split conv kernel)",
                         "voc_fused": "stages 2-3: upsample + ResBlock (+ output conv + tanh) fused, channel-last tcgen05 kind::f16 16-bit split"}
        roof = {"kernel": dom[0], "bound": "tensor", "unit": "TFLOP/s", "stage_ms_per_step": dom_ms_per_step,
                "launches_per_step": dom[1][1] // prof_steps, "launch_ms": dom[1][0] / max(dom[1][1], 1),
                "peak": peaks["bf16_tflops_sustained"],
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']})",
                "traffic": read_traffic(dom[0]), "fp32_ffma_peak_tflops": ffma_peak_tflops,
                "measured_in": f"separate instrumented pass of {prof_steps} steps (per-launch CUDA events on the launching stream)",
                "path": tensor_stages.get(dom[0], "fp32 FFMA")}
        if dom[0] in fl and dom_ms_per_step > 0:
            ach = fl[dom[0]] / (dom_ms_per_step * 1e-3) / 1e12
            roof.update(achieved=ach, frac=ach / peaks["bf16_tflops_sustained"],
                        algorithmic_flops_per_step=fl[dom[0]],
                        note="achieved = algorithmic (useful fp32-equivalent) FLOPs / device time of the stage; an fp32-faithful "
                             "split-precision kernel issues 3 products per algorithmic one: ceiling 1/3 of the bf16 peak with fp16 halves")
        else:
            roof.update(achieved=None, frac=None)
        all_stage_tflops = {k: round(fl[k] / (v[0] / prof_steps * 1e-3) / 1e12, 2) for k, v in stage_ms.items()
                            if k in fl and v[0] > 0}
        by = stage_bytes_per_step()
        stage_roofline = {}
        for k, v in stage_ms.items():
            sec = v[0] / prof_steps * 1e-3
            if sec <= 0:
                continue
            if k in by:
                ach = by[k] / sec / 1e9
                stage_roofline[k] = {"bound": "hbm", "achieved": round(ach, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                     "frac": round(ach / peaks["hbm_gbs"], 4)}
            elif k in fl:
                ach = fl[k] / sec / 1e12
                tens = k in tensor_stages or k in TENSOR_LINEAR
                stage_roofline[k] = {"bound": "tensor" if tens else "ffma", "achieved": round(ach, 2), "unit": "TFLOP/s",
                                     "peak": peaks["bf16_tflops_sustained"] if tens else round(ffma_peak_tflops, 1),
                                     "frac": round(ach / (peaks["bf16_tflops_sustained"] if tens else ffma_peak_tflops), 4)}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": warm, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "clocks": clocks, "gpu_launches": launches,
                "launches_per_step": launches // max(args.steps, 1),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4,
                        "d2h_bytes_per_step": audio_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                        "exposed_copy_ms": e2e_ms / args.steps - total_ms / args.steps,
                        "host_binding": host_binding,
                        "chunk_layout": {"sizes": pipe.sizes, "autotune": pipe.tuned},
                        "api": f"utils.host_pipeline.HostPipeline({pipe_desc}).run(decoder+vocoder, pinned host in, pinned host out)"},
                "e2e_from_ids": {"value": ids_value, "unit": UNIT, "ms_per_step": ids_ms / args.steps,
                                 "h2d_bytes_per_step": ids_host.numel() * 8 + len_host.numel() * 8 + dur_host.numel() * 4,
                                 "d2h_bytes_per_step": audio_host.numel() * 4,
                                 "api": f"utils.host_pipeline.synthesize_to_host(model, ids[{BATCH},{S}], lengths, target_durations, max_target_length={FRAMES}): acoustic front once, decoder + vocoder in utterance chunks ({ids_desc}), pinned host ids in, pinned host waveform out"},
                "parity": parity,
                "roofline": roof, "stage_roofline": stage_roofline, "stage_tflops": all_stage_tflops,
                "stage_ms_per_step": {k: round(v[0] / prof_steps, 4) for k, v in sorted(stage_ms.items(), key=lambda kv: -kv[1][0])},
                "x_realtime_per_gpu": value / world}
        if gather_ms is not None:
            line["gather"] = {"ms": gather_ms, "bytes_per_rank": audio_dev.numel() * 4, "equals_local_shard": gather_ok,
                              "api": "utils.shard.gather_batch (NCCL all_gather of [B/N,1,64T] fp32 waveforms), after the path, not in `value`"}
            if not gather_ok:
                rc = 4
        if world == 1 and not args.skip_cpu_baseline:
            os.sched_setaffinity(0, all_cores)      # the CPU leg gets every host core back
            torch.set_num_threads(len(all_cores))
            cpu = CpuPath()
            n_s = 16
            cpu.decoder_vocoder(torch.randn(2, FRAMES, HIDDEN))           # page in
            vals = cpu.time_c3(n_s, reps=1, warm=0)
            line["cpu_baseline"] = {"value": vals[0], "unit": UNIT, "cores": cpu.threads, "kind": cpu.kind,
                                    "sample": f"{n_s} of {BATCH} utterances x {FRAMES} frames, one pass, torch CPU fp32 ({'oracle/_ref: the unmodified reference modules' if cpu.kind == 'reference' else 'oracle port'})"}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rc:
        raise SystemExit(rc)


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
        if j < 2:      # wide stages: upsampling kernel, then two conv launches (C = 128) or one fused ResBlock launch (C = 64,
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
            # K = 96 linear layers sit at the split-precision ridge: reported against HBM, every operand once in fp32
            "ln_qkv": LAYERS * per * (HIDDEN + 3 * HIDDEN), "out_proj": LAYERS * per * 3 * HIDDEN,
            "ffn1": LAYERS * per * (HIDDEN + F), "ffn2": LAYERS * per * (F + 2 * HIDDEN), "ln_proj": per * (HIDDEN + MEL)}


def read_traffic(kernel: str):
    p = ROOT / "profiles" / "dram_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(kernel)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------
# Secondary lines: BASELINE.json configs[0] (C1), [1] (C2), [4] (C5). One GPU, same JSON shape, `config.workload` names them.
def _time_calls(fn, steps: int, warmup: int, dev) -> float:
    """Median milliseconds of fn() measured with CUDA events around each call (host enqueue included: these small
    configurations are launch-bound, which is the point of measuring them)."""
    for _ in range(max(warmup, 3)):
        fn()
    torch.cuda.synchronize(dev)
    ts = []
    for _ in range(steps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def run_secondary(args, local_rank: int):
    from models import _native as nat
    from models.stage_configs import STAGE_KWARGS
    from models.tts_model import M2TTSModel
    from oracle import m2tts_oracle as oracle
    from utils.graph import GraphedStep
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    cfg = args.config
    base = {"metric": METRIC, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "secondary": True}
    if cfg in ("c1", "c2"):
        torch.manual_seed(1234)
        model = M2TTSModel(**STAGE_KWARGS["stage1"]).eval().to(dev)
        sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    if cfg == "c1":
        # "Hello world" through TextProcessor: ids[0:11] = 39,21,6,24,11,40,35,7,24,17,39, rest 39, length 9.
        from utils.text import TextProcessor
        tp = TextProcessor().process_text("Hello world", max_length=256)
        ids = torch.as_tensor(tp["phoneme_ids"], dtype=torch.int64).reshape(1, -1)
        lens = torch.tensor([int(tp["length"])])
        res = {}
        for scale in (1.0, 4.0):        # scale 1: every random-init duration truncates to 0 -> one zero frame (the reference's edge case)
            ids_h, lens_h = ids.pin_memory(), lens.pin_memory()

            def call():
                mel, audio = model.inference(ids_h.to(dev, non_blocking=True), lens_h.to(dev, non_blocking=True), scale)
                return mel, audio.cpu()
            mel, audio = call()
            want_mel, want_audio = oracle.inference(sd, ids, lens, scale)
            # truncation boundary: compare shapes, and values when the frame counts agree (SURVEY §7 "hard parts")
            ok = mel.shape == want_mel.shape and float((audio - want_audio).abs().max()) <= PARITY_TOL
            ms = _time_calls(call, args.steps, args.warmup, dev)
            frames = int(mel.shape[1])
            res[f"scale{scale:g}"] = {"ms_per_call": ms, "frames": frames, "audio_s_per_s": audio_seconds(1, frames) / (ms * 1e-3), "parity_ok": bool(ok)}
        main = res["scale4"]
        line = dict(base, value=main["audio_s_per_s"], ms_per_step=main["ms_per_call"],
                    config={"workload": "C1 stage1_poc random-init single-utterance synthesis 'Hello world' through M2TTSModel.inference "
                                        "(host ids in, host waveform out), duration_scale 4 (scale 1 gives the reference's zero-frame edge case)"},
                    detail=res, e2e={"value": main["audio_s_per_s"], "unit": UNIT, "h2d_bytes_per_step": 256 * 8 + 8,
                                     "d2h_bytes_per_step": main["frames"] * 64 * 4})
    elif cfg == "c2":
        g = torch.Generator().manual_seed(0)
        ids = torch.randint(0, 256, (16, 64), generator=g)
        lengths = torch.randint(32, 65, (16,), generator=g)
        dur = torch.randint(1, 9, (16, 64), generator=g).float()
        ids_d, len_d, dur_d = ids.to(dev), lengths.to(dev), dur.to(dev)
        out = model(ids_d, len_d, target_durations=dur_d)
        ref = oracle.forward(sd, ids, lengths, dur)
        T = int(out["mel_output"].shape[1])
        valid = float(model.length_regulator.last_frames.sum()) * SAMPLES_PER_FRAME / SAMPLE_RATE
        parity = {"max_abs_mel": float((out["mel_output"].cpu() - ref["mel_output"]).abs().max()),
                  "max_abs_audio": float((out["audio_output"].cpu() - ref["audio_output"]).abs().max()), "tolerance": PARITY_TOL}
        ms = _time_calls(lambda: model(ids_d, len_d, target_durations=dur_d), args.steps, args.warmup, dev)
        before = nat.launch_count()
        model(ids_d, len_d, target_durations=dur_d)
        launches = nat.launch_count() - before
        # the same forward with the frame count given (no host read) as a CUDA-graph replay: the launch-bound small-batch path
        graphed = GraphedStep(lambda d: model(ids_d, len_d, target_durations=d, max_target_length=T)["audio_output"], dur_d)
        g_audio = graphed(dur_d).clone()
        g_ok = bool(torch.equal(g_audio, out["audio_output"]))
        g_ms = _time_calls(lambda: graphed(dur_d, check=False), args.steps, args.warmup, dev)
        nat.check_status(dev, "C2 graph replays")
        line = dict(base, value=valid / (ms * 1e-3), ms_per_step=ms, gpu_launches=launches,
                    config={"workload": f"C2 stage1_poc full pipeline (encoder, duration predictor, length regulator, decoder, vocoder), batch 16, "
                                        f"64 phonemes, T = {T} frames, {valid:.2f} valid audio-s per step, inputs in HBM, one host read (frame maximum)"},
                    padded_value=audio_seconds(16, T) / (ms * 1e-3), parity=parity,
                    cuda_graph={"ms_per_step": g_ms, "audio_s_per_s": valid / (g_ms * 1e-3), "launches_in_graph": graphed.launches_captured,
                                "equals_eager": g_ok,
                                "api": "utils.graph.GraphedStep over model(ids, lengths, target_durations=d, max_target_length=T): the frame count "
                                       "is given, so the forward has no host read and replays as one graph"})
    else:       # c5: vocoder-only sweep
        torch.manual_seed(1234)
        model = M2TTSModel(**STAGE_KWARGS["stage2"]).eval().to(dev)
        rows = []
        for B in (1, 4, 16, 64, 256):
            for T in (128, 512, 2048):
                if B * T > 64 * 2048 * 2:
                    continue
                mel = torch.randn(B, 80, T, device=dev)
                ms = _time_calls(lambda: model.vocoder(mel), max(args.steps // 2, 3), 3, dev)
                rows.append({"B": B, "T": T, "ms": round(ms, 4), "audio_s_per_s": round(audio_seconds(B, T) / (ms * 1e-3), 1)})
        mel1 = torch.randn(1, 80, 128, device=dev)
        graphed = GraphedStep(lambda m_: model.vocoder(m_), mel1)
        ms_graph = _time_calls(lambda: graphed(mel1, check=False), args.steps, args.warmup, dev)
        ms_eager = next(r["ms"] for r in rows if r["B"] == 1 and r["T"] == 128)
        best = max(rows, key=lambda r: r["audio_s_per_s"])
        line = dict(base, value=best["audio_s_per_s"], ms_per_step=best["ms"],
                    config={"workload": f"C5 HiFi-GAN vocoder-only mel->wave sweep, stage2 (80 mel, 256 channels); value = best point (B={best['B']}, T={best['T']})"},
                    sweep=rows, small_batch={"B": 1, "T": 128, "eager_ms": ms_eager, "cuda_graph_ms": round(ms_graph, 4),
                                             "launches_in_graph": graphed.launches_captured,
                                             "audio_s_per_s_graph": round(audio_seconds(1, 128) / (ms_graph * 1e-3), 1)})
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--no-bind", action="store_true", help="do not pin the rank to the cores next to its GPU (A/B of the host -> host number)")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c1", "c2", "c5"], help="c3 = headline (BASELINE configs[2]); c1/c2/c5 = secondary lines")
    ap.add_argument("--skip-cpu-baseline", action="store_true", help="profiling runs: omit the CPU leg")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="utterance chunks of the host-to-host pipeline (1 = no overlap; 0 = measured: HostPipeline.autotune)")
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
    if args.config != "c3":
        if rank == 0:
            run_secondary(args, local_rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
