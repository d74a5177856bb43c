"""Pins the CPU oracle to EVERY golden vector the reference's own unit tests hold.

Each test replays one `#[test] fn` of /root/reference/src (cited per test) against the C++
restatement in oracle/, with the reference's own exact-equality bar (`assert_eq!` on f32 is `==`).
The literal arrays were extracted mechanically by tools/extract_golden.py into tests/golden/;
scalar literals are restated inline.  CPU only - no GPU, no /root/reference at run time.
"""
import ctypes as C

import numpy as np

from conftest import f32, load_golden

SF48, FD10 = 5, 1


def i32(a):
    return np.ascontiguousarray(a, np.int32)


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    bad = np.nonzero(~(a == b))[0]
    assert bad.size == 0, f"{bad.size} mismatches, first at {bad[:5]}: {a[bad[:5]]} vs {b[bad[:5]]}"


# ------------------------------------------------------------------------------------ common/
def test_simple_config(oracle):                      # common/config.rs:109
    c = oracle.config(48000, 10)
    assert c == dict(fs=48000, fs_ind=4, z=180, nf=480, nb=64, ne=400)


def test_config_all_rates(oracle):                   # common/config.rs:42-100 (table restated by hand)
    exp = {(8000, 7.5): (60, 60, 60, 14), (16000, 7.5): (120, 120, 64, 28), (24000, 7.5): (180, 180, 64, 42),
           (32000, 7.5): (240, 240, 64, 56), (44100, 7.5): (360, 300, 64, 84), (48000, 7.5): (360, 300, 64, 84),
           (8000, 10): (80, 80, 64, 30), (16000, 10): (160, 160, 64, 60), (24000, 10): (240, 240, 64, 90),
           (32000, 10): (320, 320, 64, 120), (44100, 10): (480, 400, 64, 180), (48000, 10): (480, 400, 64, 180)}
    for (fs, ms), (nf, ne, nb, z) in exp.items():
        c = oracle.config(fs, ms)
        assert (c["nf"], c["ne"], c["nb"], c["z"]) == (nf, ne, nb, z)
    assert oracle.config(44100, 10)["fs_ind"] == 4 and oracle.config(44100, 10)["fs"] == 44100


def test_kissfft_non_inverse(oracle):                # common/kissfft.rs:298
    i_in, r_in, i_exp, r_exp = load_golden("common__kissfft__kissfft_non_inverse")
    out_r, out_i = np.zeros(240, np.float32), np.zeros(240, np.float32)
    oracle.lib().lc3o_kissfft(240, oracle.p(r_in), oracle.p(i_in), oracle.p(out_r), oracle.p(out_i))
    eq(out_i, i_exp)
    eq(out_r, r_exp)


def test_kissfft_factors(oracle):                    # common/kissfft.rs:47 kf_factor, worked by hand in DESIGN.md
    exp = {30: [2, 15, 3, 5, 5, 1], 40: [4, 10, 2, 5, 5, 1], 60: [4, 15, 3, 5, 5, 1], 80: [4, 20, 4, 5, 5, 1],
           90: [2, 45, 3, 15, 3, 5, 5, 1], 120: [4, 30, 2, 15, 3, 5, 5, 1], 160: [4, 40, 4, 10, 2, 5, 5, 1],
           180: [4, 45, 3, 15, 3, 5, 5, 1], 240: [4, 60, 4, 15, 3, 5, 5, 1]}
    for n, f in exp.items():
        out = np.zeros(64, np.int32)
        oracle.lib().lc3o_kissfft_factors(n, oracle.p(out))
        assert out[:len(f)].tolist() == f and not out[len(f):].any(), (n, out[:10])


def test_mdct_iv_run(oracle):                        # common/dct_iv.rs:80
    buf, exp = load_golden("common__dct_iv__mdct_iv_run")
    buf = buf.copy()
    oracle.lib().lc3o_dct_iv(480, oracle.p(buf))
    eq(buf, exp)


# ------------------------------------------------------------------------------------ decoder/
def _tail_usize(oracle, buf, cursor, nbits, head=0):
    b = np.array(buf, np.uint8)
    cur, out = C.c_int(cursor), C.c_uint64(0)
    ok = oracle.lib().lc3o_read_tail_usize(oracle.p(b), len(b), head, C.byref(cur), nbits, C.byref(out))
    return ok, out.value, cur.value


def test_buffer_reader(oracle):                      # decoder/buffer_reader.rs:123,132,147
    (buf,) = load_golden("decoder__buffer_reader__read_5_bits_over_byte_boundary_unto_usize")
    assert _tail_usize(oracle, buf, 23, 5) == (1, 8, 28)
    (buf,) = load_golden("decoder__buffer_reader__read_multiple_values_from_bigendian_bitstream")
    ok, v1, cur = _tail_usize(oracle, buf, 0, 3)
    ok2, v2, cur = _tail_usize(oracle, buf, cur, 8)
    assert (ok, v1, ok2, v2) == (1, 4, 1, 97)
    b = np.array([0b0100_1000], np.uint8)            # read_bool_from_bigendian_bitstream
    cur, got = C.c_int(0), []
    for _ in range(8):
        o = C.c_int(0)
        assert oracle.lib().lc3o_read_tail_bool(oracle.p(b), 1, 0, C.byref(cur), C.byref(o))
        got.append(o.value)
    assert got == [0, 0, 0, 1, 0, 0, 1, 0]


def test_buffer_reader_bounds(oracle):               # decoder/buffer_reader.rs:70-76,102-104 (error paths)
    buf = [1, 2, 3, 4]
    assert _tail_usize(oracle, buf, 0, 8, head=4)[0] == 0       # tail may not cross the head cursor
    # an aligned 8-bit read still loads num_bits/8 + 1 = 2 bytes (buffer_reader.rs:67-69), so it needs 2 free bytes
    assert _tail_usize(oracle, buf, 0, 8, head=3)[0] == 0
    assert _tail_usize(oracle, buf, 0, 8, head=2) == (1, 4, 8)
    assert _tail_usize(oracle, buf, 28, 8)[0] == 0              # needs 2 bytes, only 1 left
    b = np.array(buf, np.uint8)
    cur, o = C.c_int(8), C.c_int(0)                              # bool check is looser: len - head - byte + 2 >= 0
    assert oracle.lib().lc3o_read_tail_bool(oracle.p(b), 4, 4, C.byref(cur), C.byref(o)) == 1


def test_read_side_info(oracle):                     # decoder/side_info_reader.rs:208
    buf, _ = load_golden("decoder__side_info_reader__read_side_info_test")
    b = np.array(buf, np.uint8)
    out, cur = np.zeros(20, np.int32), C.c_int(0)
    assert oracle.lib().lc3o_read_side_info(oracle.p(b), 8, 4, 400, oracle.p(out), C.byref(cur))
    #      bw  lastnz lsb gg  ntns rc_in   lf hf lsa lsb idx_a  idx_b sub_lsb sub_msb g  pp act pidx nf
    exp = [4, 398, 0, 184, 2, 1, 1, 25, 1, 0, 0, 307189, 0, 1, 0, 0, 0, 0, 0, 6]
    assert out.tolist() == exp


def test_arithmetic_decode(oracle):                  # decoder/arithmetic_codec.rs:415
    buf, _, rc_i_exp, res_exp, order_exp = load_golden("decoder__arithmetic_codec__arithmetic_decode")
    b = np.array(buf, np.uint8)
    si = i32([4, 400, 0, 204, 2, 1, 0, 13, 4, 1, 0, 1718290, 2, 0, 0, 0, 0, 0, 0, 3])
    x, rc_i, res, misc = np.zeros(400, np.int32), np.zeros(16, np.int32), np.zeros(480, np.uint8), np.zeros(6, np.int32)
    ok = oracle.lib().lc3o_arithmetic_decode(oracle.p(b), 150, 0, 64, 4, 400, oracle.p(si), FD10, oracle.p(x),
                                             oracle.p(rc_i), oracle.p(res), oracle.p(misc))
    assert ok
    assert misc[4] == 0 and misc[5] == 1200 and misc[3] == 56909
    eq(rc_i, rc_i_exp)
    assert misc[2] == len(res_exp)
    eq(res[:misc[2]], res_exp)
    assert misc[:2].tolist() == order_exp.tolist() == [8, 0]
    # the decoded integer spectrum is pinned transitively: decode_noise_filling's spec_lines_int is this frame's x
    _, xi, _ = load_golden("decoder__noise_filling__decode_noise_filling")
    eq(x, xi)


def test_residual_spectrum_decode(oracle):           # decoder/residual_spectrum.rs:47
    bits, x, exp = load_golden("decoder__residual_spectrum__residual_spectrum_decode")
    x = x.copy()
    oracle.lib().lc3o_residual_spectrum_decode(0, oracle.p(bits), len(bits), oracle.p(x), len(x))
    eq(x, exp)


def test_decode_noise_filling(oracle):               # decoder/noise_filling.rs:65
    xf, xi, exp = load_golden("decoder__noise_filling__decode_noise_filling")
    xf = xf.copy()
    oracle.lib().lc3o_noise_filling(0, 56909, 4, FD10, 3, oracle.p(i32(xi)), oracle.p(xf), 400)
    eq(xf, exp)


def test_global_gain_decode(oracle):                 # decoder/global_gain.rs:33
    x, exp = load_golden("decoder__global_gain__global_gain_decode")
    x = x.copy()
    oracle.lib().lc3o_global_gain(1200, 4, 204, oracle.p(x), 3)
    eq(x, exp)
    assert exp[0] == f32(61.0540199)


def test_tns_decode(oracle):                         # decoder/temporal_noise_shaping.rs:147
    order, rc_i, x, exp = load_golden("decoder__temporal_noise_shaping__decode_test")
    x = x.copy()
    oracle.lib().lc3o_tns_decode(FD10, 4, 2, oracle.p(i32(order)), oracle.p(i32(rc_i)), len(rc_i), oracle.p(x))
    eq(x, exp)


def test_sns_decode(oracle):                         # decoder/spectral_noise_shaping.rs:244 (pins fast_math::exp2_raw)
    x, exp = load_golden("decoder__spectral_noise_shaping__spectral_noise_shaping_decode")
    x = x.copy()
    sns = i32([13, 4, 1, 0, 1718290, 2, 0, 0, 0])
    oracle.lib().lc3o_sns_decode(SF48, FD10, oracle.p(sns), oracle.p(x))
    eq(x, exp)


def test_mpvq_deenum(oracle):                        # decoder/spectral_noise_shaping.rs:353,362
    (e1,) = load_golden("decoder__spectral_noise_shaping__mpvq_deenum_test1")
    (e2,) = load_golden("decoder__spectral_noise_shaping__mpvq_deenum_test2")
    out = np.zeros(16, np.int32)
    oracle.lib().lc3o_mpvq_deenum(10, 10, 1, 1718290, oracle.p(out))
    eq(out, e1)
    out[:] = 0
    oracle.lib().lc3o_mpvq_deenum(6, 1, 0, 2, oracle.p(out))
    eq(out, e2)


def test_plc_save_and_load(oracle):                  # decoder/packet_loss_concealment.rs:94
    x, exp = load_golden("decoder__packet_loss_concealment__save_and_load")
    x = x.copy()
    h = C.c_void_p(oracle.lib().lc3o_plc_new(4))
    oracle.lib().lc3o_plc_save(h, oracle.p(x))
    for _ in range(3):
        oracle.lib().lc3o_plc_load(h, oracle.p(x), 4)
    oracle.lib().lc3o_plc_free(h)
    eq(x, exp)


def test_modified_dct_decode(oracle):                # decoder/modified_dct.rs:174 (two frames: overlap-add state)
    x1, x2, exp = load_golden("decoder__modified_dct__modified_dct_decode")
    h = C.c_void_p(oracle.lib().lc3o_decmdct_new(SF48, FD10))
    out = np.zeros(480, np.float32)
    oracle.lib().lc3o_decmdct_run(h, oracle.p(x1), oracle.p(out))
    oracle.lib().lc3o_decmdct_run(h, oracle.p(x2), oracle.p(out))
    oracle.lib().lc3o_decmdct_free(h)
    eq(out, exp)


def test_ltpf_full_cycle(oracle):                    # decoder/long_term_post_filter.rs:504 (6 frames, 5 transitions)
    arrs = load_golden("decoder__long_term_post_filter__long_term_post_filter_full_cycle")
    infos = [(0, 1, 134), (0, 1, 132), (1, 1, 134), (1, 1, 136), (1, 1, 136), (0, 1, 132)]   # (active, present, index)
    h = C.c_void_p(oracle.lib().lc3o_decltpf_new(SF48, FD10))
    for k, (act, pres, idx) in enumerate(infos):
        x, exp = arrs[2 * k].copy(), arrs[2 * k + 1]
        oracle.lib().lc3o_decltpf_run(h, act, pres, idx, 320, oracle.p(x))
        eq(x, exp)
    oracle.lib().lc3o_decltpf_free(h)


def test_ltpf_activated_runs(oracle):                # decoder/long_term_post_filter.rs:434 (no assertion in the reference)
    (x,) = load_golden("decoder__long_term_post_filter__long_term_post_filter_activated")
    x = x.copy()
    h = C.c_void_p(oracle.lib().lc3o_decltpf_new(SF48, FD10))
    oracle.lib().lc3o_decltpf_run(h, 1, 1, 473, 600, oracle.p(x))
    oracle.lib().lc3o_decltpf_free(h)
    assert np.isfinite(x).all()


def test_scale_and_round(oracle):                    # decoder/output_scaling.rs:34
    x, exp = load_golden("decoder__output_scaling__scale_and_round_test")
    out = np.zeros(9, np.int16)
    oracle.lib().lc3o_scale_and_round(oracle.p(x), 9, oracle.p(out))
    eq(out, exp)


def test_lc3_decode_channel(oracle):                 # decoder/lc3_decoder.rs:374 (end to end, 150 B -> 480 i16)
    buf, exp = load_golden("decoder__lc3_decoder__lc3_decode_channel")
    b = np.array(buf, np.uint8)
    h = C.c_void_p(oracle.lib().lc3o_decoder_new(SF48, FD10))
    out = np.zeros(480, np.int16)
    assert oracle.lib().lc3o_decoder_decode(h, 16, oracle.p(b), 150, oracle.p(out), 480, None, None) == 0
    eq(out, exp)
    # the one error the reference returns (lc3_decoder.rs:80)
    assert oracle.lib().lc3o_decoder_decode(h, 24, oracle.p(b), 150, oracle.p(out), 480, None, None) == 1
    oracle.lib().lc3o_decoder_free(h)


# ------------------------------------------------------------------------------------ encoder/
def test_modified_dct_encode(oracle):                # encoder/modified_dct.rs:191
    s1, s2, out_exp, eb_exp = load_golden("encoder__modified_dct__modified_dct_encode")
    h = C.c_void_p(oracle.lib().lc3o_encmdct_new(SF48, FD10))
    out, eb = np.zeros(480, np.float32), np.zeros(64, np.float32)
    oracle.lib().lc3o_encmdct_run(h, oracle.p(s1.astype(np.int16)), oracle.p(out), oracle.p(eb))
    nn = oracle.lib().lc3o_encmdct_run(h, oracle.p(s2.astype(np.int16)), oracle.p(out), oracle.p(eb))
    oracle.lib().lc3o_encmdct_free(h)
    eq(out, out_exp)
    eq(eb, eb_exp)
    assert nn == 0


def test_bandwidth_detector_run(oracle):             # encoder/bandwidth_detector.rs:137
    (eb,) = load_golden("encoder__bandwidth_detector__bandwidth_detector_run")
    out = np.zeros(2, np.int32)
    oracle.lib().lc3o_bandwidth_detect(SF48, FD10, oracle.p(eb), oracle.p(out))
    assert out.tolist() == [4, 3]


def test_attack_detector_run(oracle):                # encoder/attack_detector.rs:138
    (x,) = load_golden("encoder__attack_detector__attack_detector_run")
    h = C.c_void_p(oracle.lib().lc3o_attack_new(SF48, FD10))
    det = oracle.lib().lc3o_attack_run(h, oracle.p(x.astype(np.int16)), 150)
    f2, i4 = np.zeros(2, np.float32), np.zeros(4, np.int32)
    oracle.lib().lc3o_attack_state(h, oracle.p(f2), oracle.p(i4))
    oracle.lib().lc3o_attack_free(h)
    assert f2[0] == f32(905588.875) and f2[1] == f32(549861.5)
    assert i4.tolist() == [160, 0, 4846, 5210]
    assert det == 1


def test_sns_run(oracle):                            # encoder/spectral_noise_shaping.rs:658 (pins powf/log2/exp2/powi)
    ifs, x, eb, exp = load_golden("encoder__spectral_noise_shaping__sns_run")
    x = x.copy()
    out = np.zeros(7, np.int64)
    oracle.lib().lc3o_sns_encode(SF48, FD10, oracle.p(x), oracle.p(eb), 1, oracle.p(out))
    eq(x[:400], exp)


def test_sns_quant_run(oracle):                      # encoder/spectral_noise_shaping.rs:780
    scf, exp = load_golden("encoder__spectral_noise_shaping__sns_quant_run")
    scfq, out = np.zeros(16, np.float32), np.zeros(7, np.int64)
    oracle.lib().lc3o_sns_quant(oracle.p(scf), oracle.p(scfq), oracle.p(out))
    eq(scfq, exp)
    #                     ind_lf ind_hf shape gind ls_a ls_b joint
    assert out.tolist() == [8, 17, 3, 0, 0, 0, 15253432]


def test_temporal_noise_shaping_run(oracle):         # encoder/temporal_noise_shaping.rs:359 (pins asin/sin)
    x, exp, rc_i_exp, rc_q_exp, order_exp = load_golden("encoder__temporal_noise_shaping__temporal_noise_shaping_run")
    x = x.copy()
    oi, rq = np.zeros(21, np.int32), np.zeros(16, np.float32)
    oracle.lib().lc3o_tns_encode(SF48, FD10, oracle.p(x), 4, 1200, 0, oracle.p(oi), oracle.p(rq))
    eq(x, exp)
    eq(oi[5:], rc_i_exp)
    eq(rq, rc_q_exp)
    assert oi[:5].tolist() == [42, 0, 2, 8, 6] and order_exp.tolist() == [8, 6]


def test_long_term_post_filter_run(oracle):          # encoder/long_term_post_filter.rs:479
    (x,) = load_golden("encoder__long_term_post_filter__long_term_post_filter_run")
    h = C.c_void_p(oracle.lib().lc3o_encltpf_new(SF48, FD10))
    out = np.zeros(4, np.int32)
    oracle.lib().lc3o_encltpf_run(h, oracle.p(x.astype(np.int16)), 0, 1200, oracle.p(out))
    oracle.lib().lc3o_encltpf_free(h)
    assert out.tolist() == [0, 1, 0, 11]             # pitch_index, pitch_present, ltpf_active, nbits_ltpf


def test_long_term_post_filter_active(oracle):       # encoder/long_term_post_filter.rs:523 (8 frames)
    frames = load_golden("encoder__long_term_post_filter__long_term_post_filter_active")
    exp = [(0, 0, 0, 1), (0, 0, 0, 1), (180, 1, 0, 11), (184, 1, 0, 11), (477, 1, 0, 11), (478, 1, 0, 11),
           (478, 1, 1, 11), (478, 1, 1, 11)]
    h = C.c_void_p(oracle.lib().lc3o_encltpf_new(SF48, FD10))
    for x, e in zip(frames, exp):
        out = np.zeros(4, np.int32)
        oracle.lib().lc3o_encltpf_run(h, oracle.p(x.astype(np.int16)), 0, 400, oracle.p(out))
        assert tuple(out.tolist()) == e
    oracle.lib().lc3o_encltpf_free(h)


def test_spectral_quantization_run(oracle):          # encoder/spectral_quantization.rs:404
    xf, xq_exp = load_golden("encoder__spectral_quantization__spectral_quantization_run")
    h = C.c_void_p(oracle.lib().lc3o_quant_new(400, 4))
    xq, oi, gg = np.zeros(400, np.int16), np.zeros(7, np.int32), C.c_float(0)
    oracle.lib().lc3o_quant_run(h, oracle.p(xf), oracle.p(xq), 1200, 3, 42, 11, oracle.p(oi), C.byref(gg))
    oracle.lib().lc3o_quant_free(h)
    eq(xq, xq_exp)
    assert np.float32(gg.value) == f32(24.7091141)   # msun powf: -0.58 ulp from the true value
    # gg_ind, nbits_spec, nbits_lsb, nbits_trunc, lsb_mode, rate_flag, lastnz_trunc
    assert (oi[0], oi[2], oi[4], oi[5], oi[6]) == (193, 107, 0, 512, 350)


def test_noise_level_estimation_run(oracle):         # encoder/noise_level_estimation.rs:65
    xf, xq = load_golden("encoder__noise_level_estimation__noise_level_estimation_run")
    oracle.lib().lc3o_noise_factor.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_float]
    assert oracle.lib().lc3o_noise_factor(SF48, FD10, oracle.p(xf), oracle.p(xq.astype(np.int16)), 4,
                                          C.c_float(24.709114)) == 6


def test_bitstream_encoding_run(oracle):             # encoder/bitstream_encoding.rs:457
    rc_order, rc_i, xq, res, exp = load_golden("encoder__bitstream_encoding__bitstream_encoding_run")
    out = np.zeros(150, np.uint8)
    sns = np.array([8, 17, 3, 0, 0, 0, 15253432], np.int64)
    oracle.lib().lc3o_bitstream_encode(SF48, FD10, 4, 3, oracle.p(sns), 0, 2, oracle.p(i32(rc_order)),
                                       oracle.p(i32(rc_i)), 1, 0, 0, 193, 0, 512, 350, 107, oracle.p(res), len(res), 6,
                                       oracle.p(xq.astype(np.int16)), oracle.p(out), 150)
    eq(out, exp)


def test_buffer_writer(oracle):                      # encoder/buffer_writer.rs:75,90
    (exp,) = load_golden("encoder__buffer_writer__buffer_writer_forward_and_backwards")
    buf = np.zeros(10, np.uint8)
    ops = i32([[0, 1, 0], [1, 123, 0], [2, 22, 6], [0, 0, 0]])
    oracle.lib().lc3o_writer_script(oracle.p(buf), 10, oracle.p(ops), 4)
    eq(buf, exp)
    assert oracle.lib().lc3o_writer_nbits_side_written(140, 4, 1200) == 74


def test_lc3_encode_channel(oracle):                 # encoder/lc3_encoder.rs:314 (end to end, 480 i16 -> 150 B)
    x, exp = load_golden("encoder__lc3_encoder__lc3_encode_channel")
    h = C.c_void_p(oracle.lib().lc3o_encoder_new(SF48, FD10))
    out = np.zeros(150, np.uint8)
    oracle.lib().lc3o_encoder_encode(h, oracle.p(x.astype(np.int16)), oracle.p(out), 150)
    oracle.lib().lc3o_encoder_free(h)
    eq(out, exp)
    # bitstream_encoding_run and lc3_encode_channel assert the same 150 bytes
    eq(exp, load_golden("encoder__bitstream_encoding__bitstream_encoding_run")[4])
