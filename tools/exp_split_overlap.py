#!/usr/bin/env python3
"""Experiment: does splitting the decode48 batch into K independent sub-batches on K CUDA streams (so that one sub-batch's
kernels fill the ragged end of the other's) beat one batch of 262 144?  Uses only the public handles and the stage-mask
profiling hook.  Prints ms per 262 144 frames for each schedule."""
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
    idx = torch.from_numpy((np.arange(n) + first) % U).to(dev)
    return d, ws, fr_u[:, idx].contiguous(), torch.empty((n, NF), dtype=torch.int16, device=dev)


def timed(fn, steps=100, warm=10):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warm + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


full = make(S, 0)
print("one batch of %d: %.4f ms" % (S, timed(lambda i: full[0].decode_frames(16, full[2][i % F], full[3]))))
del full
torch.cuda.empty_cache()
for K in (2, 4):
    n = S // K
    parts = [make(n, k * n) for k in range(K)]
    streams = [torch.cuda.Stream(dev) for _ in range(K)]
    main = torch.cuda.current_stream(dev)

    def together(i):
        ev = torch.cuda.Event(); ev.record(main)
        dones = []
        for k, (d, _, fr, pcm) in enumerate(parts):
            with torch.cuda.stream(streams[k]):
                streams[k].wait_event(ev)
                d.decode_frames(16, fr[i % F], pcm)
                e = torch.cuda.Event(); e.record(streams[k]); dones.append(e)
        for e in dones:
            main.wait_event(e)

    def staggered(i):
        # sub-batch k + 1 starts when the entropy kernel of sub-batch k is done
        ev = torch.cuda.Event(); ev.record(main)
        dones = []
        for k, (d, _, fr, pcm) in enumerate(parts):
            with torch.cuda.stream(streams[k]):
                streams[k].wait_event(ev)
                d.set_stage_mask(1); d.decode_frames(16, fr[i % F], pcm)
                ev = torch.cuda.Event(); ev.record(streams[k])
                d.set_stage_mask(6); d.decode_frames(16, fr[i % F], pcm)
                d.set_stage_mask(7)
                e = torch.cuda.Event(); e.record(streams[k]); dones.append(e)
        for e in dones:
            main.wait_event(e)

    def serial(i):
        for d, _, fr, pcm in parts:
            d.decode_frames(16, fr[i % F], pcm)

    print("K=%d sub-batches, one stream (serial): %.4f ms" % (K, timed(serial)))
    print("K=%d sub-batches, K streams, all at once: %.4f ms" % (K, timed(together)))
    print("K=%d sub-batches, K streams, staggered by the entropy kernel: %.4f ms" % (K, timed(staggered)))
    del parts
    torch.cuda.empty_cache()
