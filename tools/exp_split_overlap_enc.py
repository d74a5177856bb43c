#!/usr/bin/env python3
"""Experiment: K independent encoder handles on K streams against one handle of 262 144 streams (48 kHz / 150 B)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
import lc3_codec_b200 as L
from tools.corpus import make_pcm_window

dev = torch.device("cuda:0")
S, NB, NF, F, U = 262144, 150, 480, 8, 256
sf, fd = L.SamplingFrequency.Hz48000, L.FrameDuration.TenMs
pcm_u = torch.from_numpy(make_pcm_window(U, F, 48000, NF, lead=4)).to(dev)[:, 4:]          # [U, F, nf]


def make(n, first):
    ws = torch.empty(L.Lc3BatchEncoder.calc_working_buffer_lengths(n, fd, sf, NB), dtype=torch.uint8, device=dev)
    e = L.Lc3BatchEncoder(n, fd, sf, ws, NB)
    idx = torch.from_numpy((np.arange(n) + first) % U).to(dev)
    return e, ws, pcm_u[idx].permute(1, 0, 2).contiguous(), torch.empty((n, NB), dtype=torch.uint8, device=dev)


def timed(fn, steps=30, warm=5):
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
print("one batch of %d: %.4f ms" % (S, timed(lambda i: full[0].encode_frames(full[2][i % F], full[3]))))
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
        for k, (e, _, pcm, out) in enumerate(parts):
            with torch.cuda.stream(streams[k]):
                streams[k].wait_event(ev)
                e.encode_frames(pcm[i % F], out)
                d = torch.cuda.Event(); d.record(streams[k]); dones.append(d)
        for d in dones:
            main.wait_event(d)

    print("K=%d encoder sub-batches, K streams: %.4f ms" % (K, timed(together)))
    del parts
    torch.cuda.empty_cache()
