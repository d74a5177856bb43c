#!/usr/bin/env python3
"""Dev-time generator of tests/golden/bench_c1_frames.npy: the bitstreams bench.py decodes.

1024 distinct synthetic streams (tools/corpus.py, classes sweep/noise/speech-like) x 8 consecutive frames,
48 kHz / 10 ms / 150 bytes, encoded by the ORACLE encoder (the reference encoder cannot run here).  bench.py tiles
these streams up to the batch size, so the product benchmark itself never executes oracle code.
Usage: python tools/make_bench_corpus.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import pyoracle as O  # noqa: E402
from tools.corpus import make_pcm  # noqa: E402

N_STREAMS, WARM_FRAMES, N_FRAMES = 1024, 4, 8
pcm = make_pcm(N_STREAMS, WARM_FRAMES + N_FRAMES, 48000, 480)
frames = O.encode_streams(pcm, 48000, 10, 150)[:, WARM_FRAMES:]   # skip the encoder's start-up frames
out = ROOT / "tests" / "golden" / "bench_c1_frames.npy"
np.save(out, np.ascontiguousarray(frames))
print(out, frames.shape, frames.dtype, out.stat().st_size)

# 8 kHz bitstreams for the mixed-rate bench (BASELINE config 4): the reference encoder (and therefore the GPU
# encoder) cannot be constructed at 8 kHz, so these come from the oracle encoder's spec-following 8 kHz path.
for ms, nb, nf in ((7.5, 20, 60), (10, 26, 80)):
    pcm8 = make_pcm(512, WARM_FRAMES + N_FRAMES, 8000, nf)
    fr8 = O.encode_streams(pcm8, 8000, ms, nb)[:, WARM_FRAMES:]
    out8 = ROOT / "tests" / "golden" / f"bench_mixed_8k_{str(ms).replace('.', 'p')}ms.npy"
    np.save(out8, np.ascontiguousarray(fr8))
    print(out8, fr8.shape, out8.stat().st_size)
