#!/usr/bin/env python3
"""bench.py - decoded frames/s (48 kHz, 10 ms, mono, 150 B) per GPU, with HBM-roofline fraction and CPU baseline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: launched by torchrun, one rank per GPU; streams are sharded, no collective on the data path)

A "step" is one Lc3Decoder::decode_frame for EVERY stream of the batch (one frame per stream): the batched hot path.
Workload: BASELINE.json config 5's per-GPU shard - 32,768 concurrent 48 kHz / 10 ms / 150 B mono streams per GPU
(weak scaling: every rank owns 32,768 streams).  Bitstreams: tests/golden/bench_c1_frames.npy (1,024 distinct synthetic
streams x 8 frames, oracle-encoded once by tools/make_bench_corpus.py), tiled over the batch.

Printed JSON (one line, rank 0):
  value        frames/s with inputs resident in HBM (CUDA events, max over ranks, whole job)
  e2e          same metric through the host-buffer entry point: pinned host frames in, pinned host PCM out, every step
  roofline     dominant kernel vs the measured HBM copy bandwidth; achieved = 1110 B x frames per launch / kernel time
  cpu_baseline the oracle (C++ restatement of the reference) on the host cores, bounded sample, reported not targeted
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

FS, MS, NBYTES, NF = 48000, 10, 150, 480
STREAMS_PER_GPU = 32768
ALGO_BYTES_PER_FRAME = NBYTES + 2 * NF          # 150 B in + 960 B PCM out = 1110 B
METRIC = "decoded frames/sec (48kHz 10ms mono) per GPU"
WORKLOAD = "decode 48 kHz mono 10 ms, 150 B/frame, 32768 concurrent streams per GPU, 1 frame per stream per step"


def load_frames() -> np.ndarray:
    return np.load(ROOT / "tests" / "golden" / "bench_c1_frames.npy")          # [1024, 8, 150] u8


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(np.max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(frames: np.ndarray, target_s: float = 8.0):
    """Oracle decoder on all host cores over a bounded sample of the same bitstreams."""
    from oracle import pyoracle as O
    cores = O.ncores()
    n = min(frames.shape[0], max(cores * 8, 64))
    sample = np.ascontiguousarray(frames[:n])
    O.decode_streams(sample[:cores], FS, MS)                       # warm (page-in, table init)
    t0 = time.perf_counter()
    reps = 0
    while True:
        O.decode_streams(sample, FS, MS, nthreads=cores)
        reps += 1
        if time.perf_counter() - t0 > target_s:
            break
    dt = time.perf_counter() - t0
    fps = reps * sample.shape[0] * sample.shape[1] / dt
    return {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x ({sample.shape[0]} streams x {sample.shape[1]} frames), {dt:.1f} s, one thread per core, "
                      "C++ restatement of lc3-codec (the Rust reference cannot be built here)"}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on the host cores."""
    if rank != 0:
        return
    from oracle import pyoracle as O
    frames = load_frames()
    cores = O.ncores()
    n = min(frames.shape[0], max(cores * 4, 32))
    sample = np.ascontiguousarray(frames[:n])
    per_step = sample.shape[0] * sample.shape[1]
    for _ in range(args.warmup):
        O.decode_streams(sample, FS, MS, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.decode_streams(sample, FS, MS, nthreads=cores)
    dt = time.perf_counter() - t0
    fps = args.steps * per_step / dt
    cb = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
          "sample": f"each step decodes {sample.shape[0]} streams x {sample.shape[1]} frames of the bench bitstreams on {cores} threads"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU path; a step here is a bounded sample, see cpu_baseline.sample"},
        "cpu_baseline": cb,
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--streams", type=int, default=STREAMS_PER_GPU, help="streams per GPU (default: the named workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="device-resident loop only (for ncu captures; not a bench value)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    import lc3_codec_b200 as L

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: lc3_codec_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    S = args.streams
    frames_np = load_frames()
    U, F, _ = frames_np.shape
    # rank r owns streams [r*S, (r+1)*S) of the job; stream s replays corpus stream s mod U
    idx = (np.arange(S) + rank * S) % U
    dev_frames = torch.from_numpy(frames_np).to(dev)[torch.from_numpy(idx).to(dev)].permute(1, 0, 2).contiguous()  # [F,S,150]
    sf, fd = L.SamplingFrequency.Hz48000, L.FrameDuration.TenMs
    ws_bytes = L.Lc3BatchDecoder.calc_working_buffer_lengths(S, fd, sf, NBYTES)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    dec = L.Lc3BatchDecoder(S, fd, sf, ws, NBYTES)
    pcm = torch.empty((S, NF), dtype=torch.int16, device=dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, finish=None):
        for i in range(warmup):
            fn(i)
        if finish:
            finish()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(steps):
            fn(warmup + i)
        if finish:
            finish()                      # e.g. make the stream wait for the last pipelined device->host copy
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_dev(i):
        dec.decode_frames(16, dev_frames[i % F], pcm)

    # ---- device-resident throughput (value) with clocks sampled during the timed region
    with ClockSampler(local_rank) as clk:
        ms_total = timed(step_dev, args.steps, args.warmup)
    clocks = clk.summary()
    fps = world * S * args.steps / (ms_total * 1e-3)

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "value": fps, "unit": "frames/s", "ms_per_step": ms_total / args.steps}))
        return

    # ---- per-kernel time, each kernel alone (profiling hook), same inputs
    k_steps = max(20, min(args.steps, 100))
    dec.set_stage_mask(1)
    ms_entropy = timed(step_dev, k_steps, 3) / k_steps
    dec.set_stage_mask(2)
    ms_synth = timed(step_dev, k_steps, 3) / k_steps
    dec.set_stage_mask(3)
    dom_name, dom_ms = ("lc3b::entropy_kernel", ms_entropy) if ms_entropy >= ms_synth else ("lc3b::synth_kernel", ms_synth)
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_FRAME * S / (dom_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src,
                "kernels_ms": {"entropy_kernel": ms_entropy, "synth_kernel": ms_synth},
                "algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME,
                "note": "codec stages are issue/latency bound, not HBM bound (DESIGN.md); traffic: see profiles/"}
    traffic_file = ROOT / "profiles" / "traffic.json"
    if traffic_file.exists():
        try:
            roofline["traffic"] = json.loads(traffic_file.read_text()).get(dom_name)
        except Exception:
            pass

    # ---- end to end through the host-buffer entry point: pinned host frames in, pinned host PCM out, every step
    host_frames = dev_frames.cpu().pin_memory()                         # [F,S,150]
    host_pcm = [torch.empty((S, NF), dtype=torch.int16).pin_memory() for _ in range(2)]
    dec.set_host_pipelining(True)       # PCM copy of step i overlaps the kernels of step i+1 (include/lc3b.h)

    def step_host(i):
        dec.decode_frames_host(16, host_frames[i % F], host_pcm[i & 1])

    e2e_steps = max(10, min(args.steps, 200))
    ms_e2e = timed(step_host, e2e_steps, 3, finish=dec.host_fence)
    dec.set_host_pipelining(False)
    e2e = {"value": world * S * e2e_steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": S * NBYTES,
           "d2h_bytes_per_step": S * NF * 2, "ms_per_step": ms_e2e / e2e_steps,
           "api": "lc3b_decode_frames_host (Lc3BatchDecoder.decode_frames_host), host pipelining on, "
                  "host_fence before the end event"}

    if rank == 0:
        cb = None
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(frames_np)
        state_mb = ws_bytes / 1e6
        out = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "streams_per_gpu": S, "frame_bytes": NBYTES, "nf": NF,
                       "l2": f"no explicit flush: each step streams the per-stream codec state + I/O "
                             f"({state_mb:.0f} MB workspace per GPU) which exceeds the 126 MB L2",
                       "parallelism": f"{world} x independent stream shards, no collective on the data path"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": 2 * args.steps, "roofline": roofline, "cpu_baseline": cb,
        }
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
