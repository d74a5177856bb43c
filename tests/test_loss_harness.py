"""Loss / corruption at scale: tools/loss_harness.py fault models over thousands of streams; the engine's
concealment flags, PCM and recovery behaviour against the oracle, frame-by-frame and through the time-parallel path."""
from __future__ import annotations

import numpy as np
import pytest

from tools.loss_harness import inject, recovery_distance


def test_fault_injection_is_deterministic_and_covers_every_model():
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, size=(512, 40, 60), dtype=np.uint8)
    a = inject(frames, seed=3)
    b = inject(frames, seed=3)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    damaged, lens, kind = a
    assert set(np.unique(kind)) == {0, 1, 2, 3, 4}
    assert (lens[kind == 1] == 0).all() and ((lens[kind == 3] >= 2) & (lens[kind == 3] < 60)).all()
    assert (lens[(kind != 1) & (kind != 3)] == 60).all()
    untouched = (kind == 0) | (kind == 1) | (kind == 3)
    assert np.array_equal(damaged[untouched], frames[untouched])
    assert (damaged[kind == 2] != frames[kind == 2]).any(-1).mean() > 0.99      # two flips of one bit cancel
    # burst model: mean run of drops is about 1 / p_bg = 2.5 frames
    runs = np.diff(np.flatnonzero(np.diff(np.pad((kind == 1).astype(np.int8), ((0, 0), (1, 1))).ravel()) != 0))[::2]
    assert 1.8 < runs.mean() < 3.2


def test_recovery_distance_counts_frames_until_the_clean_decode_is_matched():
    clean = np.zeros((1, 10, 4), np.int16)
    dam = clean.copy()
    dam[0, 3:6] = 100                      # frames 3 (concealed), 4 and 5 (still recovering) differ
    conc = np.zeros((1, 10), bool)
    conc[0, 3] = True
    assert recovery_distance(clean, dam, conc).tolist() == [3]


@pytest.mark.gpu
@pytest.mark.parametrize("fs,ms,nbytes,streams,frames", [(48000, 10, 150, 4096, 48), (16000, 7.5, 30, 4096, 64),
                                                         (32000, 10, 60, 2048, 48)])
def test_loss_and_corruption_at_scale(fs, ms, nbytes, streams, frames):
    from tools.loss_harness import run
    rep = run(fs, ms, nbytes, streams, frames, seed=11)
    assert rep["concealed_frames"] > 0.05 * streams * frames
    assert all(v > 0 for v in rep["damaged_frames"].values())
    assert rep["time_parallel_identical"]
    assert rep["recovery_frames"]["runs"] > 100
