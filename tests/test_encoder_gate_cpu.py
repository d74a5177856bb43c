"""The encoder parity gate itself (tests/common.py encoder_gate; SURVEY.md 8d iii), exercised on the CPU with
deliberately perturbed bitstreams: a checker that every GPU encoder test leans on must be shown to fire.

  * identical bitstreams                      -> (1.0, 0.0)
  * one frame in 1 440 differs, benignly      -> passes: identical fraction 99.93 %, |dSNR| <= 0.1 dB measured and > 0
  * one frame differs grossly                 -> fails on the SNR clause
  * two benign frames in 1 440 (99.86 %)      -> fails on the 99.9 % clause
Benign / gross single-bit flips are found by search with the oracle decoder (a flipped residual or sign bit of a small
coefficient moves the frame's SNR by a few hundredths of a dB; a flipped arithmetic-coded byte destroys it).
"""
import numpy as np
import pytest

from common import codec_delay, corpus, encoder_gate, snr_db
from oracle import pyoracle as O

FS, MS, NB = 48000, 10, 150


def _flip_effects(pcm, frames, s, f, delay):
    """|dSNR| of frame f of stream s for every single-bit flip of that frame (decode of the whole stream each time)."""
    nbits = NB * 8
    batch = np.repeat(frames[s:s + 1], nbits, axis=0)                  # [nbits, F, NB]
    idx = np.arange(nbits)
    batch[idx, f, idx // 8] ^= (1 << (idx % 8)).astype(np.uint8)
    dec = O.decode_streams(batch, FS, MS)
    base = O.decode_streams(frames[s:s + 1], FS, MS)[0]
    F, nf = pcm.shape[1], pcm.shape[2]
    x = np.concatenate([np.zeros(delay, np.int16), pcm[s].reshape(-1)])[:F * nf].reshape(F, nf)
    ref = x[f].astype(np.float64)
    s0 = snr_db(ref, base[f])
    d = np.array([abs(snr_db(ref, dec[b, f]) - s0) for b in range(nbits)])
    changed = (dec[:, f] != base[f][None]).any(-1)
    return d, changed


@pytest.fixture(scope="module")
def material():
    pcm, frames = corpus(FS, MS, NB, 48, 30)
    delay = codec_delay(pcm, O.decode_streams(frames, FS, MS))
    assert delay == 480 - 2 * 180                                      # nf - 2 z: LC3's 2.5 ms look-ahead at 10 ms frames
    picks = []
    for s, f in ((2, 14), (5, 20), (8, 11), (11, 17)):                 # speech-like / sweep streams, mid-clip frames
        d, changed = _flip_effects(pcm, frames, s, f, delay)
        benign = np.nonzero(changed & (d > 1e-4) & (d < 0.05))[0]
        gross = np.nonzero(d > 3.0)[0]
        if benign.size and gross.size:
            picks.append((s, f, int(benign[0]), int(gross[0])))
    assert len(picks) >= 2, "search found no usable bit flips"
    return pcm, frames, picks


def _flipped(frames, s, f, bit):
    g = frames.copy()
    g[s, f, bit // 8] ^= 1 << (bit % 8)
    return g


def test_identical_passes(material):
    pcm, frames, _ = material
    assert encoder_gate(pcm, frames, frames.copy(), FS, MS) == (1.0, 0.0)


def test_one_benign_frame_passes_and_is_measured(material):
    pcm, frames, picks = material
    s, f, benign, _ = picks[0]
    frac, worst = encoder_gate(pcm, frames, _flipped(frames, s, f, benign), FS, MS)
    assert 0.999 <= frac < 1.0
    assert 0.0 < worst <= 0.1


def test_gross_frame_fails_on_snr(material):
    pcm, frames, picks = material
    s, f, _, gross = picks[0]
    with pytest.raises(AssertionError, match="decoded SNR differs"):
        encoder_gate(pcm, frames, _flipped(frames, s, f, gross), FS, MS)


def test_two_benign_frames_fail_on_fraction(material):
    pcm, frames, picks = material
    g = frames
    for s, f, benign, _ in picks[:2]:
        g = _flipped(g, s, f, benign)
    with pytest.raises(AssertionError, match="byte-identical"):
        encoder_gate(pcm, frames, g, FS, MS)
