"""Full-size property tests (BASELINE.json config sizes) for the CUDA paths.

The oracle cannot decode 262 144 streams in seconds, so these runs rely on size-independent properties:
  * replication invariance - a batch built by tiling a small set of distinct streams must produce, for every replica,
    exactly the bytes / samples of the first block (catches any index arithmetic that depends on the batch size, the
    CTA-level work sorting of the entropy kernel, 32-bit offset overflow, per-warp scratch aliasing);
  * the first block itself is checked against the oracle (bit exact bytes, PCM +-1 LSB);
  * round trip: decode(encode(x)) of the full batch equals the oracle's decode(encode(x)) on the base block.
Run on the B200 box: python -m pytest tests -m gpu
"""
import numpy as np
import pytest

from common import corpus_clip
from oracle import pyoracle as O

pytestmark = pytest.mark.gpu

FULL = 262144        # BASELINE config 5 (and bench.py's default streams per GPU)
BASE = 512           # distinct streams
FRAMES = 8           # consecutive frames per stream, cut from the 200-frame SURVEY 8d clips at a per-stream offset in
                     # [10, 192] (tools.corpus.clip_offsets): the noise class is audible in ~90 % of them, as in bench.py


def _decode_full(frames_base, S, fs=48000, ms=10):
    import torch

    import lc3_codec_b200 as L
    B, F, nb = frames_base.shape
    reps = S // B
    sf, fd = L.SamplingFrequency.from_hz(fs), L.FrameDuration.from_ms(ms)
    ws = torch.empty(L.Lc3BatchDecoder.calc_working_buffer_lengths(S, fd, sf, nb), dtype=torch.uint8, device="cuda:0")
    dec = L.Lc3BatchDecoder(S, fd, sf, ws, nb)
    out = torch.zeros((S, dec.nf), dtype=torch.int16, device="cuda:0")
    first = np.zeros((B, F, dec.nf), np.int16)
    for f in range(F):
        fr = torch.from_numpy(np.ascontiguousarray(frames_base[:, f])).cuda().repeat(reps, 1)      # stream s = base s mod B
        dec.decode_frames(16, fr, out)
        blocks = out.view(reps, B, dec.nf)
        assert bool((blocks == blocks[:1]).all()), f"frame {f}: replicas differ"
        first[:, f] = blocks[0].cpu().numpy()
    return first


def test_decode_262144_streams_replication_invariance():
    _, frames = corpus_clip(48000, 10, 150, BASE, window=FRAMES)
    got = _decode_full(frames, FULL)
    exp = O.decode_streams(frames, 48000, 10)
    assert np.abs(got.astype(np.int32) - exp.astype(np.int32)).max() <= 1


def test_decode_65536_streams_16k_7p5ms():
    """BASELINE config 3's shape, four times its stream count (LTPF and TNS active)."""
    _, frames = corpus_clip(16000, 7.5, 30, BASE, window=FRAMES)
    got = _decode_full(frames, 65536, 16000, 7.5)
    exp = O.decode_streams(frames, 16000, 7.5)
    assert np.abs(got.astype(np.int32) - exp.astype(np.int32)).max() <= 1


def test_encode_262144_streams_byte_identity_and_round_trip():
    import torch

    import lc3_codec_b200 as L
    pcm, _ = corpus_clip(48000, 10, 150, BASE, window=FRAMES)
    o_frames = O.encode_streams(pcm, 48000, 10, 150)          # both encoders start from zero state on the window
    S, reps = FULL, FULL // BASE
    sf, fd = L.SamplingFrequency.Hz48000, L.FrameDuration.TenMs
    ews = torch.empty(L.Lc3BatchEncoder.calc_working_buffer_lengths(S, fd, sf, 150), dtype=torch.uint8, device="cuda:0")
    enc = L.Lc3BatchEncoder(S, fd, sf, ews, 150)
    dws = torch.empty(L.Lc3BatchDecoder.calc_working_buffer_lengths(S, fd, sf, 150), dtype=torch.uint8, device="cuda:0")
    dec = L.Lc3BatchDecoder(S, fd, sf, dws, 150)
    bits = torch.zeros((S, 150), dtype=torch.uint8, device="cuda:0")
    out = torch.zeros((S, 480), dtype=torch.int16, device="cuda:0")
    o_pcm = O.decode_streams(o_frames, 48000, 10)
    for f in range(FRAMES):
        x = torch.from_numpy(np.ascontiguousarray(pcm[:, f])).cuda().repeat(reps, 1)
        enc.encode_frames(x, bits)
        dec.decode_frames(16, bits, out)
        b = bits.view(reps, BASE, 150)
        assert bool((b == b[:1]).all()), f"frame {f}: encoded replicas differ"
        assert np.array_equal(b[0].cpu().numpy(), o_frames[:, f]), f"frame {f}: bytes differ from the oracle encoder"
        y = out.view(reps, BASE, 480)
        assert bool((y == y[:1]).all()), f"frame {f}: decoded replicas differ"
        d = np.abs(y[0].cpu().numpy().astype(np.int32) - o_pcm[:, f].astype(np.int32))
        assert d.max() <= 1


def test_encoder_large_corpus_byte_identity():
    """Parity gate (iii) on a larger corpus: >= 99.9 % of frames byte-identical (we require 100 %), several bit rates."""
    from common import gpu_encode
    total = same = 0
    for nbytes in (60, 100, 150):
        pcm, _ = corpus_clip(48000, 10, nbytes, 768, window=40)       # 40 consecutive frames somewhere in [10, 200)
        o_frames = O.encode_streams(pcm, 48000, 10, nbytes)
        g = gpu_encode(48000, 10, pcm, nbytes)
        eq = (g == o_frames).all(-1)
        total += eq.size
        same += int(eq.sum())
    assert same == total, f"{total - same} of {total} frames differ"
