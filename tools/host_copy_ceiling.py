#!/usr/bin/env python3
"""What the box can move between pinned host memory and its GPUs, with NO kernels: the ceiling of bench.py's `e2e`.

One host thread per GPU (the shape of the sharded front end, include/lc3b.h lc3b_sharded_*), each with its own pinned
buffers and CUDA stream, copying the byte counts of one decode48 step of 262 144 streams: 39.3 MB host->device
(150 B frames) and 251.7 MB device->host (480 i16 samples).  k = 1, 2, 4, 8 GPUs run concurrently; timing by CUDA
events per device, total = bytes over all GPUs / slowest GPU's time.  Three placements of the pinned memory:
  default   cudaHostAlloc from an unbound thread (first touch wherever the allocating thread runs)
  numa      the thread is bound to the GPU's NVML CPU affinity set before it allocates (one NUMA node per GPU group)
  wc        like numa, host->device source allocated write-combined (cudaHostAllocWriteCombined)
Three traffic shapes: h2d alone, d2h alone, both directions at once (two streams per GPU), which is what a pipelined
decode does.  `frames_per_s_ceiling` converts the both-directions figure into decode48 frames/s (1 110 B per frame).

  python tools/host_copy_ceiling.py [--gpus 8] [--reps 20] > gpurun_out/host_copy_ceiling.json
"""
import argparse
import ctypes
import json
import os
import threading

import torch

H2D_BYTES = 262144 * 150
D2H_BYTES = 262144 * 480 * 2


def gpu_cpus(index):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1}
        return cpus & os.sched_getaffinity(0)
    except Exception:
        return set()


def pinned(nbytes, wc=False):
    """Pinned host buffer as a uint8 tensor; wc=True uses cudaHostAllocWriteCombined through the runtime."""
    if not wc:
        return torch.empty(nbytes, dtype=torch.uint8).pin_memory(), None
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04 | 0x01))   # WC | portable
    if rc:
        raise RuntimeError(f"cudaHostAlloc WC failed: {rc}")
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), (rt, p)


def worker(gpu, placement, shape, reps, barrier, out):
    all_cpus = os.sched_getaffinity(0)
    if placement in ("numa", "wc"):
        c = gpu_cpus(gpu)
        if c:
            os.sched_setaffinity(0, c)
    dev = torch.device("cuda", gpu)
    torch.cuda.set_device(dev)
    h_in, keep = pinned(H2D_BYTES, wc=(placement == "wc"))
    h_out, _ = pinned(D2H_BYTES)
    h_in.fill_(1)
    h_out.fill_(0)
    d_in = torch.empty(H2D_BYTES, dtype=torch.uint8, device=dev)
    d_out = torch.ones(D2H_BYTES, dtype=torch.uint8, device=dev)
    s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def run(n):
        for _ in range(n):
            if shape in ("h2d", "both"):
                with torch.cuda.stream(s0):
                    d_in.copy_(h_in, non_blocking=True)
            if shape in ("d2h", "both"):
                with torch.cuda.stream(s1):
                    h_out.copy_(d_out, non_blocking=True)
    run(3)
    torch.cuda.synchronize(dev)
    barrier.wait()
    e0, e1a, e1b = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(s0)
    s1.wait_event(e0)
    run(reps)
    e1a.record(s0)
    e1b.record(s1)
    torch.cuda.synchronize(dev)
    out[gpu] = max(e0.elapsed_time(e1a), e0.elapsed_time(e1b)) * 1e-3
    barrier.wait()
    os.sched_setaffinity(0, all_cpus)
    del keep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    n = args.gpus or torch.cuda.device_count()
    ks = [k for k in (1, 2, 4, 8) if k <= n]
    rows = []
    for placement in ("default", "numa", "wc"):
        for shape in ("h2d", "d2h", "both"):
            for k in ks:
                out = {}
                barrier = threading.Barrier(k)
                th = [threading.Thread(target=worker, args=(g, placement, shape, args.reps, barrier, out)) for g in range(k)]
                [t.start() for t in th]
                [t.join() for t in th]
                t_max = max(out.values())
                per_rep = (H2D_BYTES if shape != "d2h" else 0) + (D2H_BYTES if shape != "h2d" else 0)
                row = {"placement": placement, "shape": shape, "gpus": k, "seconds_slowest_gpu": t_max,
                       "total_gbs": k * per_rep * args.reps / t_max / 1e9,
                       "per_gpu_gbs": [per_rep * args.reps / out[g] / 1e9 for g in range(k)]}
                if shape == "both":
                    row["frames_per_s_ceiling"] = k * 262144 * args.reps / t_max
                rows.append(row)
                print(json.dumps(row), flush=True)
    doc = {"what": "bare pinned cudaMemcpyAsync bandwidth, no kernels, one host thread per GPU (tools/host_copy_ceiling.py)",
           "h2d_bytes_per_rep": H2D_BYTES, "d2h_bytes_per_rep": D2H_BYTES, "reps": args.reps, "n_gpus_visible": n,
           "cpu_count": os.cpu_count(), "affinity_cpus": len(os.sched_getaffinity(0)),
           "gpu_cpu_affinity_sizes": [len(gpu_cpus(g)) for g in range(n)], "rows": rows}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/host_copy_ceiling.json", "w") as f:
        json.dump(doc, f, indent=1)


if __name__ == "__main__":
    main()
