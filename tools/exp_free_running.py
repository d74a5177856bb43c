#!/usr/bin/env python3
"""Experiment: K quarter-size handles on K streams, NOT joined between steps (each stream runs its own sequence of calls):
what per-call fork/join costs against free-running lanes.  ms per 262 144 frames."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import lc3_codec_b200 as L
from bench import load_frames

dev = torch.device("cuda:0")
S, NB, NF, F, U = 262144, 150, 480, 8, 1024
sf, fd = L.SamplingFrequency.Hz48000, L.FrameDuration.TenMs
fr_u = torch.from_numpy(load_frames()).to(dev).permute(1, 0, 2).contiguous()


def make(n, first):
    ws = torch.empty(L.Lc3BatchDecoder.calc_working_buffer_lengths(n, fd, sf, NB), dtype=torch.uint8, device=dev)
    d = L.Lc3BatchDecoder(n, fd, sf, ws, NB)
    d.set_min_nbytes(NB)
    d.set_split(1)
    d.set_graph_mode(False)
    idx = torch.from_numpy((np.arange(n) + first) % U).to(dev)
    return d, ws, fr_u[:, idx].contiguous(), torch.empty((n, NF), dtype=torch.int16, device=dev)


for K in (4, 8):
    n = S // K
    parts = [make(n, k * n) for k in range(K)]
    streams = [torch.cuda.Stream(dev) for _ in range(K)]
    main = torch.cuda.current_stream(dev)

    def run(steps, joined):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(main)
        for s in streams:
            s.wait_event(e0)
        for i in range(steps):
            if joined:
                ev = torch.cuda.Event(); ev.record(main)
            for k, (d, _, fr, pcm) in enumerate(parts):
                with torch.cuda.stream(streams[k]):
                    if joined:
                        streams[k].wait_event(ev)
                    d.decode_frames(16, fr[i % F], pcm)
                    if joined:
                        e = torch.cuda.Event(); e.record(streams[k]); main.wait_event(e)
        if not joined:
            for s in streams:
                e = torch.cuda.Event(); e.record(s); main.wait_event(e)
        e1.record(main)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    run(10, True); run(10, False)
    print("K=%d joined every step: %.4f ms   free-running lanes: %.4f ms" % (K, run(100, True), run(100, False)))
    del parts
    torch.cuda.empty_cache()
