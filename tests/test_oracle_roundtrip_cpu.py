"""Cross-check for the configurations no reference golden vector covers (SURVEY.md 8c "unpinned"): the oracle encoder
followed by the oracle decoder must reproduce the input audio, delayed by the codec's algorithmic delay, at a
signal-to-noise ratio that a working LC3 chain reaches at that bit rate.  A restatement error in either direction
(window, MDCT scaling, SNS/TNS mirror, bit allocation) collapses the SNR to ~0 dB, so modest floors are discriminating.
Runs on CPU (the oracle is test infrastructure)."""
import numpy as np
import pytest

from common import ALL_CONFIGS
from oracle import pyoracle as O
from tools.corpus import MIXED_NBYTES, make_pcm


def _snr_db(ref, out):
    """Best-alignment SNR: the decoder output lags the input by the codec delay (found by cross-correlation)."""
    ref = ref.astype(np.float64)
    out = out.astype(np.float64)
    best = -1e9
    n = len(ref)
    for d in range(0, 400):
        a, b = ref[: n - d], out[d:]
        err = a - b
        snr = 10 * np.log10((a * a).sum() / max((err * err).sum(), 1e-9))
        best = max(best, snr)
    return best


@pytest.mark.parametrize("fs,ms", [c for c in ALL_CONFIGS if c[0] != 8000])
def test_encode_decode_reproduces_the_audio(fs, ms):
    cfg = O.config(fs, ms)
    nf = cfg["nf"]
    pcm = make_pcm(6, 60, fs, nf)                               # sweep, noise, speech-like, ...
    nbytes = MIXED_NBYTES[(fs, ms)]
    frames = O.encode_streams(pcm, fs, ms, nbytes)
    out = O.decode_streams(frames, fs, ms)
    snrs = []
    for s in (0, 2, 3, 5):                                      # sweeps and speech-like streams (noise is not waveform-coded well)
        x = pcm[s, 10:].reshape(-1)
        y = out[s, 10:].reshape(-1)
        snrs.append(_snr_db(x, y))
    assert min(snrs) > 6.0 and np.mean(snrs) > 10.0, snrs


def test_higher_rate_gives_higher_snr():
    pcm = make_pcm(3, 60, 48000, 480)
    res = []
    for nbytes in (60, 100, 150):
        out = O.decode_streams(O.encode_streams(pcm, 48000, 10, nbytes), 48000, 10)
        res.append(_snr_db(pcm[2, 10:].reshape(-1), out[2, 10:].reshape(-1)))
    assert res[0] < res[1] < res[2], res
