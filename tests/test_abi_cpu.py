"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/lc3b.h
declares, argument validation mirrors the reference's error behaviour, and the engine's own f32 transcendentals
(host build of csrc/lc3b_math.cuh) agree bit-for-bit with the oracle's msun restatement.  No compute on a GPU here.
"""
import ctypes as C
import re
import zlib
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def L():
    import lc3_codec_b200 as m
    m.lib()
    return m


def test_every_declared_symbol_is_exported(L):
    header = (ROOT / "include" / "lc3b.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = sorted(set(re.findall(r"\b(lc3b_\w+)\s*\(", header)))
    assert len(declared) >= 12
    lib = L.lib()
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, f"declared in include/lc3b.h but not exported: {missing}"
    from lc3_codec_b200 import native
    assert sorted(native.EXPORTS) == declared


def test_config_matches_reference_table(L, oracle):      # src/common/config.rs:42-100
    from lc3_codec_b200 import native
    for hz in (8000, 16000, 24000, 32000, 44100, 48000):
        for ms in (7.5, 10):
            c = native.config(L.SamplingFrequency.from_hz(hz), L.FrameDuration.from_ms(ms))
            o = oracle.config(hz, ms)
            assert (c.fs_ind, c.fs, c.ne, c.nb, c.nf, c.z) == (o["fs_ind"], o["fs"], o["ne"], o["nb"], o["nf"], o["z"])
    bad = native.Config()
    assert L.lib().lc3b_config_new(9, 0, C.byref(bad)) == 2


def test_workspace_bytes_and_validation(L):
    f = L.Lc3BatchDecoder.calc_working_buffer_lengths
    a = f(1024, L.FrameDuration.TenMs, L.SamplingFrequency.Hz48000, 150)
    b = f(2048, L.FrameDuration.TenMs, L.SamplingFrequency.Hz48000, 150)
    assert 0 < a < b < 2.1 * a
    # persistent per-stream state: 2 spectrum slots + overlap + LTPF history dominate (DESIGN.md "Data layout")
    per_stream = (b - a) / 1024
    assert 2 * 400 * 4 + 300 * 4 + 960 * 4 <= per_stream <= 16384
    with pytest.raises(L.Lc3bError):
        f(0, L.FrameDuration.TenMs, L.SamplingFrequency.Hz48000, 150)
    with pytest.raises(L.Lc3bError):
        f(16, L.FrameDuration.TenMs, L.SamplingFrequency.Hz48000, 401)
    n = C.c_size_t(0)
    assert L.lib().lc3b_decoder_workspace_bytes(16, 1, 5, 150, None) == 2
    # null handle / null buffers are rejected, never dereferenced
    assert L.lib().lc3b_decode_frames(None, 16, None, None, 150, 150, None, 480, None, None) == 2
    assert L.lib().lc3b_decoder_set_trace(None, None, None) == 2


def test_tables_crc():
    """lc3_tables.h carries a CRC32 per table (tools/gen_tables.py); recompute from the literals in the header."""
    import struct
    text = (ROOT / "lc3_codec_b200" / "csrc" / "lc3_tables.h").read_text()
    n_checked = 0
    for m in re.finditer(r"crc32=(0x[0-9a-f]+) \*/\nLC3_TABLE\((\w+)\) (\w+)((?:\[\d+\])+) = \{\n(.*?)\n\};", text, flags=re.S):
        crc, ctype, name, dims, body = m.groups()
        vals = [v.strip() for v in body.replace("\n", " ").split(",") if v.strip()]
        if ctype == "float":
            raw = b"".join(struct.pack("<f", float.fromhex(v[:-1])) for v in vals)
        else:
            raw = b"".join(struct.pack("<q", int(v)) for v in vals)
        assert zlib.crc32(raw) == int(crc, 16), name
        n_checked += 1
    assert n_checked == 53


MATH_CASES = [
    # which (engine, oracle index), x sampler, y sampler
    ("powf", 0, 0, lambda r, n: np.full(n, 10.0, np.float32), lambda r, n: (r.integers(-245, 146, n) / np.float32(28.0)).astype(np.float32)),
    ("powf_tilt", 0, 0, lambda r, n: np.full(n, 10.0, np.float32), lambda r, n: (r.integers(0, 64, n).astype(np.float32) * np.float32(30.0 / 630.0))),
    ("log2f", 1, 1, lambda r, n: np.exp(r.uniform(-30, 40, n)).astype(np.float32), None),
    ("log10f", 2, 2, lambda r, n: np.exp(r.uniform(-20, 45, n)).astype(np.float32), None),
    ("exp2f", 3, 3, lambda r, n: r.uniform(-40, 40, n).astype(np.float32), None),
    ("asinf", 4, 4, lambda r, n: r.uniform(-1, 1, n).astype(np.float32), None),
    ("exp2_raw", 5, 6, lambda r, n: r.uniform(-20, 20, n).astype(np.float32), None),
]


@pytest.mark.parametrize("name,which,owhich,xs,ys", MATH_CASES, ids=[c[0] for c in MATH_CASES])
def test_engine_math_equals_oracle_math_host(L, oracle, name, which, owhich, xs, ys):
    """Two independent statements of the msun algorithms (oracle/lc3o_math.cpp, csrc/lc3b_math.cuh) must agree
    bit for bit - the encoder's byte-exactness rests on it."""
    rng = np.random.default_rng(1234 + which)
    n = 200_000
    x = xs(rng, n)
    y = ys(rng, n) if ys else np.zeros(n, np.float32)
    a, b = np.zeros(n, np.float32), np.zeros(n, np.float32)
    assert L.lib().lc3b_selftest_math_host(which, oracle.p(x), oracle.p(y), oracle.p(a), n) == 0
    oracle.lib().lc3o_math_vec(owhich, oracle.p(x), oracle.p(y), oracle.p(b), n)
    bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
    assert bad.size == 0, f"{name}: {bad.size} differ, e.g. x={x[bad[:3]]} y={y[bad[:3]]} -> {a[bad[:3]]} vs {b[bad[:3]]}"


def test_msun_accuracy_vs_float64(oracle):
    """Sanity of the restated constants: every msun-style routine stays within 1 ulp of the float64 result
    (msun documents < 1 ulp); a mistyped constant would blow this by orders of magnitude."""
    rng = np.random.default_rng(5)
    n = 100_000

    def ulp_err(got, ref64):
        ref32 = ref64.astype(np.float32)
        ulp = np.spacing(np.abs(ref32)).astype(np.float64)
        return np.max(np.abs(got.astype(np.float64) - ref64) / ulp)

    def run(which, x, y=None):
        out = np.zeros(len(x), np.float32)
        yy = np.zeros(len(x), np.float32) if y is None else y
        oracle.lib().lc3o_math_vec(which, oracle.p(x), oracle.p(yy), oracle.p(out), len(x))
        return out

    x = np.exp(rng.uniform(-30, 40, n)).astype(np.float32)
    assert ulp_err(run(1, x), np.log2(x.astype(np.float64))) < 1.0
    assert ulp_err(run(2, x), np.log10(x.astype(np.float64))) < 1.0
    x = rng.uniform(-40, 40, n).astype(np.float32)
    assert ulp_err(run(3, x), np.exp2(x.astype(np.float64))) < 0.51
    x = rng.uniform(-1, 1, n).astype(np.float32)
    assert ulp_err(run(4, x), np.arcsin(x.astype(np.float64))) < 1.0
    y = (rng.integers(-245, 146, n) / np.float32(28.0)).astype(np.float32)
    assert ulp_err(run(0, np.full(n, 10.0, np.float32), y), np.power(10.0, y.astype(np.float64))) < 1.0
    x = (np.float32(np.pi / 17.0) * np.arange(-8, 9, dtype=np.float32)).astype(np.float32)
    assert np.array_equal(run(5, x), np.sin(x.astype(np.float64)).astype(np.float32))   # sinf: f64 kernels, correctly rounded here


def test_mixed_layout_is_bucket_ordered(L):
    """lc3b_mixed_decoder_layout (host only): buckets sorted by (sampling frequency, duration), streams of a bucket in
    original order, dense host PCM offsets; bad enums rejected."""
    import ctypes as C

    import numpy as np
    from lc3_codec_b200 import native
    rng = np.random.default_rng(2)
    n = 1000
    sf = rng.integers(0, 6, n).astype(np.int32)
    fd = rng.integers(0, 2, n).astype(np.int32)
    order = np.zeros(n, np.int32)
    buckets = (native.MixedBucket * 12)()
    nb, elems = C.c_int32(0), C.c_uint64(0)
    assert L.lib().lc3b_mixed_decoder_layout(n, sf.ctypes.data, fd.ctypes.data, order.ctypes.data, buckets, C.byref(nb), C.byref(elems)) == 0
    assert sorted(order.tolist()) == list(range(n))
    keys = [(int(sf[s]), int(fd[s]), int(s)) for s in order]
    assert keys == sorted(keys)
    row, off = 0, 0
    for b in buckets[:nb.value]:
        assert b.first_row == row and b.host_pcm_offset == off
        cfg = native.config(b.sampling_frequency, b.frame_duration)
        assert b.nf == cfg.nf and all(sf[s] == b.sampling_frequency and fd[s] == b.frame_duration for s in order[row:row + b.n_rows])
        row += b.n_rows
        off += b.n_rows * b.nf
    assert row == n and off == elems.value
    sf[5] = 6
    assert L.lib().lc3b_mixed_decoder_layout(n, sf.ctypes.data, fd.ctypes.data, None, None, None, None) == 2
    size = C.c_size_t(0)
    sf[5] = 0
    assert L.lib().lc3b_mixed_decoder_workspace_bytes(n, sf.ctypes.data, fd.ctypes.data, 120, C.byref(size)) == 0 and size.value > 0
    assert L.lib().lc3b_mixed_decoder_workspace_bytes(n, sf.ctypes.data, fd.ctypes.data, 401, C.byref(size)) == 2


def test_new_entry_points_reject_null_handles(L):
    lib = L.lib()
    assert lib.lc3b_decoder_set_graph_mode(None, 1) == 2
    assert lib.lc3b_decoder_set_split(None, 2) == 2
    assert lib.lc3b_encoder_set_graph_mode(None, 1) == 2
    assert lib.lc3b_mixed_decode_frames(None, 16, None, None, 10, 10, None, 480, None, None) == 2
    assert lib.lc3b_mixed_decoder_host_fence(None, None) == 2
    assert lib.lc3b_sharded_decoder_wait(None) == 2 and lib.lc3b_sharded_encoder_wait(None) == 2
    assert lib.lc3b_sharded_decoder_n_shards(None) == 0
    assert lib.lc3b_sharded_decode_frames_host(None, 16, None, None, 10, 10, None, 480, None) == 2
    lib.lc3b_mixed_decoder_destroy(None)
    lib.lc3b_sharded_decoder_destroy(None)
    lib.lc3b_sharded_encoder_destroy(None)
    lib.lc3b_host_free(None)


def test_frame_duration_from_ms_rejects_other_values(L):
    assert L.FrameDuration.from_ms(7.5) == L.FrameDuration.SevenPointFiveMs and L.FrameDuration.from_ms(10) == L.FrameDuration.TenMs
    for bad in (5, 2.5, 0, 20):
        with pytest.raises(ValueError):
            L.FrameDuration.from_ms(bad)
