// lc3b engine, encoder kernel 2 of 2: spectrum + analysis results -> bitstream, one THREAD per frame.
// Compiled with -fmad=false: every expression below rounds exactly like the reference's f32 code.
//
// Replaces, per stream, the second half of EncoderChannel::encode (src/encoder/lc3_encoder.rs:74-110):
//   BandwidthDetector::run          src/encoder/bandwidth_detector.rs:64
//   SpectralNoiseShaping::run       src/encoder/spectral_noise_shaping.rs:203 (two-stage VQ :318/:363, MPVQ enumeration :585)
//   TemporalNoiseShaping::run       src/encoder/temporal_noise_shaping.rs:40 (autocorrelation, Levinson-Durbin, lattice)
//   SpectralQuantization::run       src/encoder/spectral_quantization.rs:75 (gain bisection, bit consumption, adjustment)
//   ResidualBitsEncoder::encode     src/encoder/residual_spectrum.rs:33
//   NoiseLevelEstimation            src/encoder/noise_level_estimation.rs:21
//   BitstreamEncoding::encode       src/encoder/bitstream_encoding.rs:77 with BufferWriter (buffer_writer.rs:5)
// Everything here is decision logic and serial recurrences (codebook argmins, Levinson, the rate loop's context walk,
// the range coder), so the parallel axis is frames.  The frame's bytes are assembled in a shared-memory row and
// copied out coalesced.
#include <stdio.h>
#include "lc3b_enc_common.cuh"
#include "lc3b_math.cuh"
#include "lc3_tables.h"

namespace lc3b {

struct QuantParams {
    const EncConfig* cfg;
    int n_streams, nbytes;
    size_t frame_stride;
    float* xf;
    const float* e_b;
    const int32_t* ehand;
    int32_t* estate;
    int16_t* xq;
    float* scratch_e;
    uint8_t* lsbs;
    uint8_t* frames_out;
    int row_pitch;
    int debug;
};

constexpr int QNT_THREADS = 32;
// The frame's spectrum (f32), quantised spectrum (i16) and 4-line energies live in shared memory as per-thread
// columns: element k of lane l at [k * XS + l].  XS = 33 keeps both the thread-private walks (fixed lane, any k) and
// the transposed load of stream-major rows (fixed row, consecutive k) free of bank conflicts.
constexpr int XS = 33;

struct BwRes { int bw, nbits; };
struct SnsRes { int ind_lf, ind_hf, shape_j, gind, ls_inda, ls_indb; uint64_t joint; };
struct TnsRes { int nbits_tns, lpc_weighting, num_filters; int rc_order[2]; int rc_i[16]; float rc_q[16]; };
struct QRes { int gg_ind, nbits_spec, nbits_lsb, nbits_trunc, lsb_mode, rate_flag, lastnz_trunc; float gg; };

// Constant tables live at namespace scope: a function-local 64-bit table (msun's exp2f) was observed to be laid
// over live stack data by nvcc 12.9 at -O3, so nothing table-like is left on the stack in this file.
__device__ const int START10[4][4] = {{53, 0, 0, 0}, {47, 59, 0, 0}, {44, 54, 60, 0}, {41, 51, 57, 61}};   // bandwidth_detector.rs:5-18
__device__ const int STOP10[4][4] = {{63, 0, 0, 0}, {56, 63, 0, 0}, {52, 59, 63, 0}, {49, 55, 60, 63}};
__device__ const int START75[4][4] = {{51, 0, 0, 0}, {45, 58, 0, 0}, {42, 53, 60, 0}, {40, 51, 57, 61}};
__device__ const int STOP75[4][4] = {{63, 0, 0, 0}, {55, 63, 0, 0}, {51, 58, 63, 0}, {48, 55, 60, 63}};
__device__ const int NBITS_BW[5] = {0, 1, 2, 2, 3};
__device__ const int QUIET[4] = {20, 10, 10, 10}, CUTOFF[4] = {15, 23, 20, 20};
__device__ const int L10[4] = {4, 4, 3, 1}, L75[4] = {4, 4, 3, 2};
__device__ const float SNS_W[6] = {1.0f / 12.0f, 2.0f / 12.0f, 3.0f / 12.0f, 3.0f / 12.0f, 2.0f / 12.0f, 1.0f / 12.0f};   // spectral_noise_shaping.rs:59
__device__ const float TNS_LAG[9] = {1.0f, 0.9980280260203829f, 0.9921354055113971f, 0.9823915844707989f, 0.9689107911912967f,
                                     0.9518498073692735f, 0.9314049334023056f, 0.9078082299969592f, 0.8813231366694713f};   // temporal_noise_shaping.rs:81-84
__device__ const int GGA_T1[5] = {80, 230, 380, 530, 680}, GGA_T2[5] = {500, 1025, 1550, 2075, 2600},
                     GGA_T3[5] = {850, 1700, 2550, 3400, 4250};   // spectral_quantization.rs:351-353

// ---------------------------------------------------------------- bandwidth_detector.rs:64-127
__device__ BwRes bandwidth_detect(const EncConfig& c, const float* e_b) {
    const int n_bw = c.fs_ind, nbits = NBITS_BW[n_bw];
    if (n_bw == 0) return {0, nbits};
    const bool d10 = c.n_ms == LC3B_10MS;
    const int* start = (d10 ? START10 : START75)[n_bw - 1];
    const int* stop = (d10 ? STOP10 : STOP75)[n_bw - 1];
    const int* l = d10 ? L10 : L75;
    int bw = 0;
    for (int k = n_bw - 1; k >= 0; k--) {
        const float width = (float)(stop[k] + 1 - start[k]);
        float quiet = 0.0f;
        for (int n = start[k]; n <= stop[k]; n++) quiet += e_b[n] / width;
        if (quiet >= (float)QUIET[k]) { bw = k + 1; break; }
    }
    if (n_bw == bw) return {bw, nbits};
    float cutoff_max = 0.0f;
    const int l_bw = l[bw];
    const int from = start[bw] + 1 - l_bw, to = start[bw];
    for (int n = from; n < to; n++) {
        const float cutoff = e_b[n - l_bw] / e_b[n];
        cutoff_max = maxf_rs(cutoff, cutoff_max);
    }
    if (cutoff_max > (float)CUTOFF[bw]) return {bw, nbits};
    return {n_bw, nbits};
}

// ---------------------------------------------------------------- spectral_noise_shaping.rs
__device__ void add_unit_pulse(const float* abs_x, int n_max, int k, int k_max, int* cand, float* corr_xy, float* energy_y) {   // :285-316
    float corr_last = *corr_xy, en_last = *energy_y;
    for (int it = k; it < k_max; it++) {
        int n_best = 0;
        *corr_xy = corr_last + abs_x[0];
        float best_corr_sq = *corr_xy * *corr_xy;
        float best_en = en_last + 2.0f * (float)cand[0] + 1.0f;
        for (int n_c = 1; n_c < n_max; n_c++) {
            *corr_xy = corr_last + abs_x[n_c];
            *energy_y = en_last + 2.0f * (float)cand[n_c] + 1.0f;
            if (*corr_xy * *corr_xy * best_en > best_corr_sq * *energy_y) {
                n_best = n_c;
                best_corr_sq = *corr_xy * *corr_xy;
                best_en = *energy_y;
            }
        }
        corr_last += abs_x[n_best];
        en_last += 2.0f * (float)cand[n_best] + 1.0f;
        cand[n_best] += 1;
    }
}

__device__ void normalize_candidate(const int* y, float* xq, int n_max) {   // :629-648
    float norm = 0.0f;
    for (int n = 0; n < n_max; n++) if (y[n] != 0) norm += (float)y[n] * (float)y[n];
    norm = sqrtf(norm);
    for (int n = 0; n < n_max; n++) {
        xq[n] = (float)y[n];
        if (y[n] != 0) xq[n] /= norm;
    }
    for (int n = n_max; n < 16; n++) xq[n] = 0.0f;
}

__device__ void mvpq_enum(uint64_t* index, int* lead_sign_ind, int dim_in, const int* vec_in) {   // :585-627
    int next_sign_ind = INT32_MIN;
    int8_t k_val_acc = 0;
    *index = 0;
    int n = 0;
    uint64_t tmp_h_row = LC3T_MPVQ_OFFSETS[n][0];
    for (int pos = dim_in - 1; pos >= 0; pos--) {
        const int8_t tmp_val = (int8_t)vec_in[pos];
        if (((uint32_t)next_sign_ind & 0x80000000u) == 0 && tmp_val != 0) *index = 2 * *index + (uint64_t)next_sign_ind;
        if (tmp_val < 0) next_sign_ind = 1;
        else if (tmp_val > 0) next_sign_ind = 0;
        *index += tmp_h_row;
        k_val_acc = (int8_t)(k_val_acc + (tmp_val < 0 ? -tmp_val : tmp_val));
        if (pos != 0) n += 1;
        tmp_h_row = (k_val_acc >= 11) ? LC3T_MPVQ_OFFSETS[n + 1][k_val_acc % 11] : LC3T_MPVQ_OFFSETS[n][k_val_acc];
    }
    *lead_sign_ind = next_sign_ind;
}

__device__ void sns_run_quant(const float* scf, float* scfq, SnsRes* res) {   // :318-582
    float st1[16], r1[16];
    float dlf_min = INFINITY, dhf_min = INFINITY;
    int ind_lf = 0, ind_hf = 0;
    for (int i = 0; i < 32; i++) {
        float dlf = 0.0f, dhf = 0.0f;
        for (int n = 0; n < 8; n++) {
            dlf += (scf[n] - LC3T_LFCB[i][n]) * (scf[n] - LC3T_LFCB[i][n]);
            dhf += (scf[8 + n] - LC3T_HFCB[i][n]) * (scf[8 + n] - LC3T_HFCB[i][n]);
        }
        if (dlf < dlf_min) { ind_lf = i; dlf_min = dlf; }
        if (dhf < dhf_min) { ind_hf = i; dhf_min = dhf; }
    }
    for (int n = 0; n < 8; n++) { st1[n] = LC3T_LFCB[ind_lf][n]; st1[8 + n] = LC3T_HFCB[ind_hf][n]; }
    for (int n = 0; n < 16; n++) r1[n] = scf[n] - st1[n];

    float t2rot[16];
    int y0[16], y1[16], y2[16], y3[16];
    float xq0[16], xq1[16], xq2[16], xq3[16];
    for (int n = 0; n < 16; n++) { t2rot[n] = 0.0f; y0[n] = y1[n] = y2[n] = y3[n] = 0; }
    for (int row = 0; row < 16; row++)
        for (int n = 0; n < 16; n++) t2rot[n] += r1[row] * LC3T_D[row][n];
    int k = 0;
    float abs_sum = 0.0f, abs_x[16];
    for (int n = 0; n < 16; n++) { abs_x[n] = fabsf(t2rot[n]); abs_sum += abs_x[n]; }
    const float proj = (6.0f - 1.0f) / abs_sum;
    float corr_xy = 0.0f, energy_y = 0.0f;
    for (int n = 0; n < 16; n++) {
        y3[n] = cast_i32(floorf(abs_x[n] * proj));
        if (y3[n] != 0) {
            k += y3[n];
            corr_xy += (float)y3[n] * abs_x[n];
            energy_y += (float)y3[n] * (float)y3[n];
        }
    }
    add_unit_pulse(abs_x, 16, k, 6, y3, &corr_xy, &energy_y);
    for (int n = 0; n < 16; n++) y2[n] = y3[n];
    add_unit_pulse(abs_x, 16, 6, 8, y2, &corr_xy, &energy_y);
    for (int n = 0; n < 10; n++) y1[n] = y2[n];
    int k1 = 8;
    for (int n = 10; n < 16; n++) {
        if (y2[n] != 0) {
            k1 -= y2[n];
            corr_xy -= (float)y2[n] * abs_x[n];
            energy_y -= (float)y2[n] * (float)y2[n];
        }
    }
    add_unit_pulse(abs_x, 10, k1, 10, y1, &corr_xy, &energy_y);
    for (int n = 0; n < 10; n++) y0[n] = y1[n];
    float max_abs_x = 0.0f;
    int n_best = 0;
    for (int n_c = 10; n_c < 16; n_c++) {
        y0[n_c] = 0;
        if (abs_x[n_c] > max_abs_x) { max_abs_x = abs_x[n_c]; n_best = n_c; }
    }
    y0[n_best] = 1;
    for (int n = 0; n < 10; n++)
        if (t2rot[n] < 0.0f) { y0[n] *= -1; y1[n] *= -1; y2[n] *= -1; y3[n] *= -1; }
    for (int n = 10; n < 16; n++)
        if (t2rot[n] < 0.0f) { y0[n] *= -1; y2[n] *= -1; y3[n] *= -1; }
    normalize_candidate(y0, xq0, 16);
    normalize_candidate(y1, xq1, 10);
    normalize_candidate(y2, xq2, 16);
    normalize_candidate(y3, xq3, 16);

    int shape_j = 0, gind = 0;
    float g_sel = 0.0f;
    const float* xq_sel = xq0;
    float d_min = INFINITY;
    for (int j = 0; j < 4; j++) {
        int g_max;
        const float *gains, *xq;
        switch (j) {
            case 0: g_max = 1; gains = LC3T_SNS_VQ_REG_ADJ_GAINS; xq = xq0; break;
            case 1: g_max = 3; gains = LC3T_SNS_VQ_REG_LF_ADJ_GAINS; xq = xq1; break;
            case 2: g_max = 3; gains = LC3T_SNS_VQ_NEAR_ADJ_GAINS; xq = xq2; break;
            default: g_max = 7; gains = LC3T_SNS_VQ_FAR_ADJ_GAINS; xq = xq3; break;
        }
        for (int i = 0; i < g_max; i++) {
            float d = 0.0f;
            for (int n = 0; n < 16; n++) {
                const float diff = t2rot[n] - gains[i] * xq[n];
                d += diff * diff;
            }
            if (d < d_min) { shape_j = j; gind = i; d_min = d; g_sel = gains[i]; xq_sel = xq; }
        }
    }
    const int lsb_gain = gind & 1;
    uint64_t idxa = 0, idxb = 0;
    int ls_inda = 0, ls_indb = 0;
    uint64_t joint;
    switch (shape_j) {
        case 0:
            mvpq_enum(&idxa, &ls_inda, 10, y0);
            mvpq_enum(&idxb, &ls_indb, 6, y0 + 10);
            joint = (2 * idxb + (uint64_t)(int64_t)ls_indb + 2) * 2390004ull + idxa;
            break;
        case 1:
            mvpq_enum(&idxa, &ls_inda, 10, y1);
            joint = (uint64_t)lsb_gain * 2390004ull + idxa;
            break;
        case 2:
            mvpq_enum(&idxa, &ls_inda, 16, y2);
            joint = idxa;
            break;
        default:
            mvpq_enum(&idxa, &ls_inda, 16, y3);
            joint = 15158272ull + (uint64_t)lsb_gain + 2 * idxa;
            break;
    }
    for (int n = 0; n < 16; n++) {
        float factor = 0.0f;
        for (int col = 0; col < 16; col++) factor += xq_sel[col] * LC3T_D[n][col];
        scfq[n] = st1[n] + g_sel * factor;
    }
    res->ind_lf = ind_lf; res->ind_hf = ind_hf; res->shape_j = shape_j; res->gind = gind;
    res->ls_inda = ls_inda; res->ls_indb = ls_indb; res->joint = joint;
}

__device__ SnsRes sns_encode(const EncConfig& c, float* x, const float* e_b, bool attack) {   // :203-282
    const float* W = SNS_W;
    float padded[64], e[64];
    const int nb = c.nb, diff = 64 - nb;
    if (diff > 0) {
        for (int i = 0; i < 64; i++) padded[i] = 0.0f;
        for (int i = 0; i < diff; i++) { padded[2 * i] = e_b[i]; padded[2 * i + 1] = e_b[i]; }
        for (int i = 0; i < nb - diff; i++) padded[2 * diff + i] = e_b[diff + i];
    } else {
        for (int i = 0; i < 64; i++) padded[i] = e_b[i];
    }
    e[0] = 0.75f * padded[0] + 0.25f * padded[1];
    for (int b = 1; b < 63; b++) e[b] = 0.25f * padded[b - 1] + 0.5f * padded[b] + 0.25f * padded[b + 1];
    e[63] = 0.25f * padded[62] + 0.75f * padded[63];
    for (int b = 0; b < 64; b++) e[b] *= c.pre_emph[b];
    float total = 0.0f;
    for (int b = 0; b < 64; b++) total += e[b];
    total = (total / 64.0f) * powi_nt(10.0f, -4);
    const float floor_ = maxf_rs(powi_nt(2.0f, -32), total);
    for (int b = 0; b < 64; b++) e[b] = maxf_rs(e[b], floor_);
    for (int b = 0; b < 64; b++) e[b] = log2f_msun(1.1920929e-07f + e[b]) / 2.0f;
    float ds[16];
    ds[0] = W[0] * e[0];
    for (int k = 1; k < 6; k++) ds[0] += W[k] * e[k - 1];
    for (int b2 = 1; b2 < 15; b2++) {
        float v = 0.0f;
        const int from = 4 * b2 - 1;
        for (int k = 0; k < 6; k++) v += W[k] * e[from + k];
        ds[b2] = v;
    }
    ds[15] = W[5] * e[63];
    for (int k = 0; k < 5; k++) ds[15] += W[k] * e[60 + k - 1];
    float tot = 0.0f;
    for (int n = 0; n < 16; n++) tot += ds[n];
    const float avg = tot / 16.0f;
    for (int n = 0; n < 16; n++) ds[n] = 0.85f * (ds[n] - avg);
    float scf[16];
    if (attack) {
        scf[0] = (ds[0] + ds[1] + ds[2]) / 3.0f;
        scf[1] = (ds[0] + ds[1] + ds[2] + ds[3]) / 4.0f;
        for (int n = 2; n < 14; n++) {
            float s = 0.0f;
            for (int j = n - 2; j < n + 3; j++) s += ds[j];
            scf[n] = s / 5.0f;
        }
        scf[14] = (ds[12] + ds[13] + ds[14] + ds[15]) / 4.0f;
        scf[15] = (ds[13] + ds[14] + ds[15]) / 3.0f;
        float st = 0.0f;
        for (int n = 0; n < 16; n++) st += scf[n];
        const float sa = st / 16.0f;
        const float att = c.n_ms == LC3B_10MS ? 0.5f : 0.3f;
        for (int n = 0; n < 16; n++) scf[n] = att * (scf[n] - sa);
    } else {
        for (int n = 0; n < 16; n++) scf[n] = ds[n];
    }
    float scfq[16];
    SnsRes res;
    sns_run_quant(scf, scfq, &res);
#ifdef LC3B_DEBUG_PRINT
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        printf("e_b0..3 %g %g %g %g  e[0..3] %g %g %g %g pre %g %g\n", e_b[0], e_b[1], e_b[2], e_b[3], e[0], e[1], e[2], e[3], c.pre_emph[0], c.pre_emph[1]);
        printf("scf "); for (int n = 0; n < 16; n++) printf("%g ", scf[n]); printf("\nscfq "); for (int n = 0; n < 16; n++) printf("%g ", scfq[n]);
        printf("\nres %d %d %d %d\n", res.ind_lf, res.ind_hf, res.shape_j, res.gind);
    }
#endif
    float it[64];
    it[0] = scfq[0];
    it[1] = scfq[0];
    for (int n = 0; n < 15; n++) {
        const float d = scfq[n + 1] - scfq[n];
        it[4 * n + 2] = scfq[n] + (0.125f * d);
        it[4 * n + 3] = scfq[n] + (0.375f * d);
        it[4 * n + 4] = scfq[n] + (0.625f * d);
        it[4 * n + 5] = scfq[n] + (0.875f * d);
    }
    it[62] = scfq[15] + (0.125f * (scfq[15] - scfq[14]));
    it[63] = scfq[15] + (0.375f * (scfq[15] - scfq[14]));
    if (diff > 0) {
        for (int i = 0; i < diff; i++) it[i] = (it[2 * i] + it[2 * i + 1]) / 2.0f;
        for (int i = diff; i < nb; i++) it[i] = it[diff + 1];
    }
    for (int b = 0; b < nb; b++) {
        const float g = exp2f_msun(-it[b]);
        for (int k = c.band_idx[b]; k < c.band_idx[b + 1]; k++) x[k * XS] *= g;
    }
    return res;
}

// ---------------------------------------------------------------- temporal_noise_shaping.rs
__device__ int8_t tns_to_int(float x) {
    auto cast_i8 = [](float v) -> int8_t {
        if (v != v) return 0;
        if (v >= 127.0f) return 127;
        if (v <= -128.0f) return -128;
        return (int8_t)v;
    };
    if (x >= 0.0f) return cast_i8(x + 0.5f);
    return cast_i8(-(-x + 0.5f));
}

struct TnsP { int nf; int start[2], stop[2], ss[2][3], se[2][3]; };
__device__ const TnsP TNS_T10[5] = {
    {1, {12, 160}, {80, 0}, {{12, 34, 57}, {0, 0, 0}}, {{34, 57, 80}, {0, 0, 0}}},
    {1, {12, 160}, {160, 0}, {{12, 61, 110}, {0, 0, 0}}, {{61, 110, 160}, {0, 0, 0}}},
    {1, {12, 160}, {200, 0}, {{12, 88, 164}, {0, 0, 0}}, {{88, 164, 240}, {0, 0, 0}}},
    {2, {12, 160}, {160, 320}, {{12, 61, 110}, {160, 213, 266}}, {{61, 110, 160}, {213, 266, 320}}},
    {2, {12, 200}, {200, 400}, {{12, 74, 137}, {200, 266, 333}}, {{74, 137, 200}, {266, 333, 400}}},
};
__device__ const TnsP TNS_T75[5] = {
    {1, {9, 120}, {60, 0}, {{9, 26, 43}, {0, 0, 0}}, {{26, 43, 60}, {0, 0, 0}}},
    {1, {9, 120}, {120, 0}, {{9, 46, 83}, {0, 0, 0}}, {{46, 83, 120}, {0, 0, 0}}},
    {1, {9, 120}, {180, 0}, {{9, 66, 123}, {0, 0, 0}}, {{66, 123, 180}, {0, 0, 0}}},
    {2, {9, 120}, {120, 240}, {{9, 46, 82}, {120, 159, 200}}, {{46, 82, 120}, {159, 200, 240}}},
    {2, {9, 150}, {150, 300}, {{9, 56, 103}, {150, 200, 250}}, {{56, 103, 150}, {200, 250, 300}}},
};

__device__ void tns_encode(const EncConfig& c, float* x, int p_bw, int nbits, bool near_nyquist, TnsRes& r) {   // :40-78
    const TnsP& tp = (c.n_ms == LC3B_10MS ? TNS_T10 : TNS_T75)[p_bw];
    for (int i = 0; i < 16; i++) { r.rc_i[i] = 0; r.rc_q[i] = 0.0f; }
    r.num_filters = tp.nf;
    r.lpc_weighting = (c.n_ms == LC3B_10MS ? nbits < 480 : nbits < 360) ? 1 : 0;
    const int ne = c.ne;
    const float* LAG = TNS_LAG;
    for (int f = 0; f < tp.nf; f++) {
        // compute_normalized_autocorrelation :80-115.  The reference recomputes the sub-block energy for every lag and
        // walks the sub-block once per lag; the sums below are the same sums (same operands, same order), gathered in
        // ONE walk per sub-block with the last eight lines held in registers.
        float es_s[3], ac_s[3][9];
        for (int sb = 0; sb < 3; sb++) {
            const int start = tp.ss[f][sb], stop = tp.se[f][sb];
            float acc[9], w8[8];
#pragma unroll
            for (int k = 0; k < 9; k++) acc[k] = 0.0f;
#pragma unroll
            for (int k = 0; k < 8; k++) w8[k] = 0.0f;
            int cnt = 0;
            for (int n = start; n < stop; n++) {
                const float xn = x[n * XS];
                acc[0] += xn * xn;
#pragma unroll
                for (int k = 1; k < 9; k++) if (cnt >= k) acc[k] += w8[k - 1] * xn;
#pragma unroll
                for (int k = 7; k > 0; k--) w8[k] = w8[k - 1];
                w8[0] = xn;
                cnt++;
            }
            es_s[sb] = acc[0];
#pragma unroll
            for (int k = 0; k < 9; k++) ac_s[sb][k] = acc[k];
        }
        float rr[9];
        for (int k = 0; k < 9; k++) {
            const float r0 = k == 0 ? 3.0f : 0.0f;
            float rk = 0.0f, e_prod = 1.0f;
            for (int sb = 0; sb < 3; sb++) {
                e_prod *= es_s[sb];
                rk += ac_s[sb][k] / es_s[sb];
            }
            rr[k] = (e_prod == 0.0f ? r0 : rk) * LAG[k];
        }
        float mem[2][9];
        for (int i = 0; i < 9; i++) mem[0][i] = mem[1][i] = 0.0f;
        float* a = mem[0];
        float* a_last = mem[1];
        float e = rr[0];
        a[0] = 1.0f;
        for (int k = 1; k < 9; k++) {
            float* tmp = a_last; a_last = a; a = tmp;
            float rc = 0.0f;
            for (int n = 0; n < k; n++) rc -= a_last[n] * rr[k - n];
            if (e != 0.0f) rc /= e;
            a[0] = 1.0f;
            for (int n = 1; n < k; n++) a[n] = a_last[n] + rc * a_last[k - n];
            a[k] = rc;
            e *= 1.0f - rc * rc;
        }
        const float pred_gain = e == 0.0f ? rr[0] : rr[0] / e;
        float* rcq = r.rc_q + f * 8;
        if (pred_gain > 1.5f && !near_nyquist) {
            float gamma = 1.0f;
            if (r.lpc_weighting > 0 && pred_gain < 2.0f) gamma -= (1.0f - 0.85f) * (2.0f - pred_gain) / (2.0f - 1.5f);
            for (int k = 0; k < 9; k++) a[k] *= powi_nt(gamma, k);
            float* a_k = a;
            float* a_km1 = a_last;
            for (int k = 8; k >= 1; k--) {
                rcq[k - 1] = a_k[k];
                const float ee = 1.0f - rcq[k - 1] * rcq[k - 1];
                for (int n = 1; n < k; n++) {
                    a_km1[n] = a_k[n] - rcq[k - 1] * a_k[k - n];
                    a_km1[n] /= ee;
                }
                float* tmp = a_k; a_k = a_km1; a_km1 = tmp;
            }
        } else {
            for (int k = 0; k < 8; k++) rcq[k] = 0.0f;
        }
    }
    const float step = (float)M_PI / 17.0f;
    for (int f = 0; f < tp.nf; f++) {
        for (int k = 0; k < 8; k++) {
            const int idx = f * 8 + k;
            r.rc_i[idx] = (int)(tns_to_int(asinf_msun(r.rc_q[idx]) / step) + 8);
            r.rc_q[idx] = c.tns_sin[r.rc_i[idx]];        // sin(step * (rc_i - 8)), tabulated (17 arguments)
        }
        int k = 7;
        while (k >= 0 && r.rc_i[f * 8 + k] == 8) k--;
        r.rc_order[f] = k + 1;
    }
    for (int f = tp.nf; f < 2; f++) {
        for (int k = 0; k < 8; k++) { r.rc_i[f * 8 + k] = 8; r.rc_q[f * 8 + k] = 0.0f; }
        r.rc_order[f] = 0;
    }
    int nbits_tns = 0;
    for (int f = 0; f < tp.nf; f++) {
        const int ob = r.rc_order[f] != 0 ? LC3T_AC_TNS_ORDER_BITS[r.lpc_weighting][r.rc_order[f] - 1] : 0;
        int cb = 0;
        for (int k = 0; k < r.rc_order[f]; k++) cb += LC3T_AC_TNS_COEF_BITS[k][r.rc_i[f * 8 + k]];
        nbits_tns += (int)ceilf((2048.0f + (float)ob + (float)cb) / 2048.0f);
    }
    r.nbits_tns = nbits_tns;
    float st[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int f = 0; f < tp.nf; f++) {
        if (r.rc_order[f] == 0) continue;
        for (int n = tp.start[f]; n < tp.stop[f]; n++) {
            float t = x[n * XS], st_save = t;
            const int po = r.rc_order[f] - 1;
            for (int k = 0; k < po; k++) {
                const float rq = r.rc_q[f * 8 + k];
                const float st_tmp = rq * t + st[k];
                t += rq * st[k];
                st[k] = st_save;
                st_save = st_tmp;
            }
            t += r.rc_q[f * 8 + po] * st[po];
            st[po] = st_save;
            x[n * XS] = t;
        }
    }
}

// ---------------------------------------------------------------- spectral_quantization.rs
struct BitCons { int rate_flag, lastnz, nbits_lsb, lastnz_trunc, nbits_est, nbits_trunc; bool mode_flag; };

__device__ BitCons compute_bit_consumption(int ne, int fs_ind, const int16_t* xq, int nbits, int nbits_spec) {   // :265-348
    BitCons bc;
    bc.rate_flag = nbits > (160 + fs_ind * 160) ? 512 : 0;
    bc.mode_flag = nbits >= (480 + fs_ind * 160);
    int lastnz = ne;
    while (lastnz > 2 && xq[(lastnz - 1) * XS] == 0 && xq[(lastnz - 2) * XS] == 0) lastnz -= 2;
    uint32_t est = 0, trunc = 0;
    int nbits_lsb = 0, lastnz_trunc = 2, c = 0;
    for (int n = 0; n < lastnz; n += 2) {
        int t = c + bc.rate_flag;
        if (n > ne / 2) t += 256;
        const int q0 = xq[n * XS], q1 = xq[(n + 1) * XS];
        uint32_t a = (uint32_t)(q0 < 0 ? -q0 : q0) & 0xffffu, a_lsb = a;
        uint32_t b = (uint32_t)(q1 < 0 ? -q1 : q1) & 0xffffu, b_lsb = b;
        int lev = 0;
        while ((a > b ? a : b) >= 4) {
            const int pki = LC3T_AC_SPEC_LOOKUP[t + lev * 1024];
            est += LC3T_AC_SPEC_BITS[pki][16];
            if (lev == 0 && bc.mode_flag) nbits_lsb += 2; else est += 2 * 2048;
            a >>= 1;
            b >>= 1;
            lev = lev + 1 < 3 ? lev + 1 : 3;
        }
        const int pki = LC3T_AC_SPEC_LOOKUP[t + lev * 1024];
        const int sym = (int)(a + 4 * b);
        est += LC3T_AC_SPEC_BITS[pki][sym];
        if (a_lsb > 0) est += 2048;
        if (b_lsb > 0) est += 2048;
        if (lev > 0 && bc.mode_flag) {
            a_lsb >>= 1;
            b_lsb >>= 1;
            if (a_lsb == 0 && q0 != 0) nbits_lsb += 1;
            if (b_lsb == 0 && q1 != 0) nbits_lsb += 1;
        }
        if ((q0 != 0 || q1 != 0) && (int)ceilf((float)est / 2048.0f) <= nbits_spec) {
            lastnz_trunc = n + 2;
            trunc = est;
        }
        t = lev <= 1 ? 1 + (int)(a + b) * (lev + 1) : 12 + lev;
        c = (c & 15) * 16 + t;
    }
    bc.lastnz = lastnz;
    bc.lastnz_trunc = lastnz_trunc;
    bc.nbits_lsb = nbits_lsb;
    bc.nbits_est = (int)ceilf((float)est / 2048.0f) + nbits_lsb;
    bc.nbits_trunc = (int)ceilf((float)trunc / 2048.0f);
    return bc;
}

__device__ float gain_of(const EncConfig& c, int gg_ind, int gg_off) {   // 10^((gg_ind + gg_off) / 28), :239
    const int idx = gg_ind + gg_off + 245;
    if (idx >= 0 && idx < 400) return c.gg_table[idx];
    return powf_msun(10.0f, ((float)gg_ind + (float)gg_off) / 28.0f);
}

__device__ BitCons quantize_spectrum(const EncConfig& c, const float* xf, int16_t* xq, int nbits, int gg_off, int gg_ind,
                                     int nbits_spec, float* gg_out, bool* lsb_mode) {   // :230-263
    const int ne = c.ne;
    const float gg = gain_of(c, gg_ind, gg_off);
    for (int k = 0; k < ne; k++) {
        const float v = xf[k * XS];
        xq[k * XS] = v >= 0.0f ? cast_i16(v / gg + 0.375f) : cast_i16(v / gg - 0.375f);
    }
    BitCons bc = compute_bit_consumption(ne, c.fs_ind, xq, nbits, nbits_spec);
    for (int k = bc.lastnz_trunc; k < bc.lastnz; k++) xq[k * XS] = 0;
    *gg_out = gg;
    *lsb_mode = bc.mode_flag && bc.nbits_est > nbits_spec;
    return bc;
}

__device__ QRes spectral_quantization(const EncConfig& c, int32_t* es, const float* xf, int16_t* xq, float* e, int nbits,
                                      int nbits_bw, int nbits_tns, int nbits_ltpf) {   // :75-120
    const int ne = c.ne, fs_ind = c.fs_ind;
    int lg = 0;
    while ((1 << lg) < ne / 2) lg++;
    const int nbits_ari = lg + (nbits <= 1280 ? 3 : nbits <= 2560 ? 4 : 5);
    const int nbits_spec = nbits - (nbits_bw + nbits_tns + nbits_ltpf + 38 + 8 + 3 + nbits_ari);
    const bool reset_offset_old = es[ES_Q_RESET_OFFSET_OLD] != 0;
    const float nbits_offset_old = __uint_as_float((uint32_t)es[ES_Q_NBITS_OFFSET_OLD]);
    const int nbits_est_old = es[ES_Q_NBITS_EST_OLD];
    float nbits_offset;
    if (reset_offset_old) nbits_offset = 0.0f;
    else {
        const float prev = nbits_offset_old + (float)0 - (float)nbits_est_old;   // nbits_spec_old is never updated (:59)
        nbits_offset = 0.8f * nbits_offset_old + 0.2f * minf_rs(40.0f, maxf_rs(-40.0f, prev));
    }
    int nbits_spec_adj;
    {
        const float v = (float)nbits_spec + nbits_offset + 0.5f;       // `as u16`: saturating
        nbits_spec_adj = v != v ? 0 : v >= 65535.0f ? 65535 : v <= 0.0f ? 0 : (int)v;
    }
    const int q = (int16_t)nbits / (int16_t)(10 * (fs_ind + 1));
    const int gg_off = -(115 < q ? 115 : q) - 105 - 5 * (fs_ind + 1);
    const int ne4 = ne / 4;
    for (int i = 0; i < ne4; i++) {                                    // compute_spectral_energy :390-395
        const float* pp = xf + 4 * i * XS;
        const float total = pp[0] * pp[0] + pp[XS] * pp[XS] + pp[2 * XS] * pp[2 * XS] + pp[3 * XS] * pp[3 * XS];
        e[i * XS] = 10.0f * log10f_msun(1.1920929e-07f + total);
    }
    int fac = 256, gg_ind = 255;                                       // global_gain_estimation :174-210
    for (int it = 0; it < 8; it++) {
        fac >>= 1;
        gg_ind -= fac;
        float tmp = 0.0f;
        bool is_zero = true;
        const float g = (float)gg_ind + (float)gg_off;
        for (int i = ne4 - 1; i >= 0; i--) {
            const float ei = e[i * XS];
            if (ei * 28.0f / 20.0f < g) {
                if (!is_zero) tmp += 2.7f * 28.0f / 20.0f;
            } else {
                if (g < (ei * 28.0f / 20.0f - 43.0f * 28.0f / 20.0f))
                    tmp += 2.0f * ei * 28.0f / 20.0f - 2.0f * g - 36.0f * 28.0f / 20.0f;
                else
                    tmp += ei * 28.0f / 20.0f - g + 7.0f * 28.0f / 20.0f;
                is_zero = false;
            }
        }
        if ((tmp > (float)nbits_spec_adj * 1.4f * 28.0f / 20.0f) && !is_zero) gg_ind += fac;
    }
    float xmax = 0.0f;                                                 // global_gain_limitation :212-228
    for (int k = 0; k < ne; k++) xmax = maxf_rs(xmax, fabsf(xf[k * XS]));
    int gg_min = 0;
    if (xmax > 0.0f) gg_min = (int)(int16_t)(cast_i16(ceilf(28.0f * log10f_msun(xmax / (32768.0f - 0.375f)))) - (int16_t)gg_off);
    bool reset_offset;
    if (gg_ind < gg_min || xmax == 0.0f) { reset_offset = true; gg_ind = gg_min; } else reset_offset = false;

    float gg;
    bool lsb_mode;
    BitCons bc = quantize_spectrum(c, xf, xq, nbits, gg_off, gg_ind, nbits_spec, &gg, &lsb_mode);
    es[ES_Q_NBITS_OFFSET_OLD] = (int32_t)__float_as_uint(nbits_offset);   // state saved BEFORE the adjustment (:96-100)
    es[ES_Q_NBITS_EST_OLD] = bc.nbits_est;
    es[ES_Q_RESET_OFFSET_OLD] = reset_offset;
    const int t1 = GGA_T1[fs_ind], t2 = GGA_T2[fs_ind], t3 = GGA_T3[fs_ind];
    const int nbits_est = bc.nbits_est;
    float delta;
    if (nbits_est < t1) delta = ((float)nbits_est + 48.0f) / 16.0f;
    else if (nbits_est < t2) {
        const float tmp1 = (float)t1 / 16.0f + 3.0f, tmp2 = (float)t2 / 48.0f;
        delta = ((float)nbits_est - (float)t1) * (tmp2 - tmp1) / ((float)t2 - (float)t1) + tmp1;
    } else if (nbits_est < t3) delta = (float)nbits_est / 48.0f;
    else delta = (float)t3 / 48.0f;
    delta = floorf(delta + 0.5f);
    const float delta2 = delta + 2.0f;
    const int origin = gg_ind;
    if ((gg_ind < 255 && nbits_est > nbits_spec) || (gg_ind > 0 && (float)nbits_est < ((float)nbits_spec - delta2))) {
        if ((float)nbits_est < ((float)nbits_spec - delta2)) gg_ind -= 1;
        else if (gg_ind == 254 || (float)nbits_est < ((float)nbits_spec + delta)) gg_ind += 1;
        else gg_ind += 2;
        gg_ind = gg_ind > gg_min ? gg_ind : gg_min;
    }
    if (origin != gg_ind) bc = quantize_spectrum(c, xf, xq, nbits, gg_off, gg_ind, nbits_spec, &gg, &lsb_mode);
    QRes r;
    r.gg_ind = gg_ind; r.nbits_spec = nbits_spec; r.nbits_lsb = bc.nbits_lsb; r.lsb_mode = lsb_mode;
    r.nbits_trunc = bc.nbits_trunc; r.rate_flag = bc.rate_flag; r.lastnz_trunc = bc.lastnz_trunc; r.gg = gg;
    return r;
}

// ---------------------------------------------------------------- noise_level_estimation.rs:21-55
__device__ int noise_factor(const EncConfig& c, const float* xf, const int16_t* xq, int bw_ind, float gg) {
    const bool d10 = c.n_ms == LC3B_10MS;
    const int bw_stop = d10 ? 80 * (bw_ind + 1) : 60 * (bw_ind + 1);
    const int nf_start = d10 ? 24 : 18, nf_width = d10 ? 3 : 2;
    float sum = 0.0f;
    int count = 0;
    const int nf_stop = c.ne < bw_stop ? c.ne : bw_stop;
    // a line is relevant when xq is zero on [k - w, min(bw_stop - 1, k + w)]; track the last non-zero index seen by
    // a scan pointer that runs w lines ahead instead of re-reading the window for every line
    int last_nz = -1000, scan = nf_start - nf_width;
    for (int k = nf_start; k < nf_stop; k++) {
        const int hi = bw_stop - 1 < k + nf_width ? bw_stop - 1 : k + nf_width;
        while (scan <= hi) { if (xq[scan * XS] != 0) last_nz = scan; scan++; }
        if (last_nz < k - nf_width) { sum += fabsf(xf[k * XS]) / gg; count++; }
    }
    const float level = count > 0 ? sum / (float)count : 0.0f;
    const float diff = 8.0f - 16.0f * level;
    if (diff >= 0.0f) {
        const int v = cast_i32(diff + 0.5f);
        return v < 7 ? v : 7;
    }
    return 0;
}

// ---------------------------------------------------------------- bitstream_encoding.rs + buffer_writer.rs
struct Writer {
    uint8_t* buf;
    int bp;
    int bp_side;
    uint32_t mask_side;
    __device__ __forceinline__ void bool_backward(bool bit) {                 // buffer_writer.rs:27-40
        if (!bit) buf[bp_side] &= (uint8_t)~mask_side; else buf[bp_side] |= (uint8_t)mask_side;
        if (mask_side == 0x80) { mask_side = 1; bp_side -= 1; } else mask_side <<= 1;
    }
    __device__ __forceinline__ void uint_backward(uint64_t val, int nbits) {   // :19-25
        for (int i = 0; i < nbits; i++) { bool_backward((val & 1) != 0); val >>= 1; }
    }
    __device__ __forceinline__ void uint_forward(uint32_t val, int nbits) {    // :42-53 (QUIRK: bp is not advanced)
        uint32_t mask = 0x80;
        for (int i = 0; i < nbits; i++) {
            if (((val & 0xff) & mask) == 0) buf[bp] &= (uint8_t)~mask; else buf[bp] |= (uint8_t)mask;
            mask >>= 1;
        }
    }
    __device__ __forceinline__ void byte_forward(uint32_t v) { buf[bp++] = (uint8_t)v; }
    __device__ __forceinline__ int nbits_side_written(int nbits) const { return nbits - (8 * bp_side + 8 - (31 - __clz(mask_side))); }
};
struct AcEnc { uint32_t low, range; int cache, carry, carry_count; };

__device__ __forceinline__ void ac_shift(AcEnc& st, Writer& w) {              // bitstream_encoding.rs:397-415
    if (st.low < 0x00ff0000u || st.carry == 1) {
        if (st.cache >= 0) w.byte_forward((uint32_t)((st.cache + st.carry) & 0xff));
        while (st.carry_count > 0) {
            w.byte_forward((uint32_t)((st.carry + 0xff) & 0xff));
            st.carry_count -= 1;
        }
        st.cache = (int)(st.low >> 16);
        st.carry = 0;
    } else {
        st.carry_count += 1;
    }
    st.low <<= 8;
    st.low &= 0x00ffffffu;
}
__device__ __forceinline__ void ac_encode(AcEnc& st, Writer& w, int cum, int freq) {   // :417-429
    const uint32_t r = st.range >> 10;
    st.low += r * (uint32_t)cum;
    if (st.low >> 24 != 0) st.carry = 1;
    st.low &= 0x00ffffffu;
    st.range = r * (uint32_t)freq;
    while (st.range < 0x10000u) {
        st.range <<= 8;
        ac_shift(st, w);
    }
}

__device__ void bitstream_encode(const EncConfig& c, const BwRes& bw, const SnsRes& sns, const TnsRes& tns, int pitch_present,
                                 int ltpf_active, int pitch_index, const QRes& q, const uint32_t* res_bits, int n_res,
                                 int nf_factor, const int16_t* xq, uint8_t* lsbs, uint8_t* out, int nbytes) {   // :77-136
    const int ne = c.ne, nbits = nbytes * 8;
    for (int i = 0; i < nbytes; i++) out[i] = 0;
    Writer w;
    w.buf = out;
    w.bp = 0;
    w.bp_side = nbytes - 1;
    w.mask_side = 1;
    if (bw.nbits > 0) w.uint_backward((uint64_t)bw.bw, bw.nbits);
    {
        int lg = 0;
        while ((1 << lg) < ne / 2) lg++;
        w.uint_backward((uint64_t)((q.lastnz_trunc >> 1) - 1), lg);
    }
    w.bool_backward(q.lsb_mode != 0);
    w.uint_backward((uint64_t)(int64_t)q.gg_ind, 8);
    for (int f = 0; f < tns.num_filters; f++) w.bool_backward(tns.rc_order[f] != 0);
    w.bool_backward(pitch_present != 0);
    w.uint_backward((uint64_t)sns.ind_lf, 5);
    w.uint_backward((uint64_t)sns.ind_hf, 5);
    {
        const bool submode_msb = (sns.shape_j >> 1) != 0;
        w.bool_backward(submode_msb);
        const int gain_msbs = sns.gind >> LC3T_SNS_GAIN_LSB_BITS[sns.shape_j];
        w.uint_backward((uint64_t)gain_msbs, LC3T_SNS_GAIN_MSB_BITS[sns.shape_j]);
        w.bool_backward(sns.ls_inda != 0);
        if (!submode_msb) {
            w.uint_backward(sns.joint, 13);
            w.uint_backward(sns.joint >> 13, 12);
        } else {
            w.uint_backward(sns.joint, 12);
            w.uint_backward(sns.joint >> 12, 12);
        }
    }
    if (pitch_present) {
        w.bool_backward(ltpf_active != 0);
        w.uint_backward((uint64_t)pitch_index, 9);
    }
    w.uint_backward((uint64_t)nf_factor, 3);
    AcEnc st{0, 0x00ffffffu, -1, 0, 0};
    for (int f = 0; f < tns.num_filters; f++) {
        if (tns.rc_order[f] > 0) {
            ac_encode(st, w, LC3T_AC_TNS_ORDER_CUMFREQ[tns.lpc_weighting][tns.rc_order[f] - 1],
                      LC3T_AC_TNS_ORDER_FREQ[tns.lpc_weighting][tns.rc_order[f] - 1]);
            for (int k = 0; k < tns.rc_order[f]; k++)
                ac_encode(st, w, LC3T_AC_TNS_COEF_CUMFREQ[k][tns.rc_i[k + 8 * f]], LC3T_AC_TNS_COEF_FREQ[k][tns.rc_i[k + 8 * f]]);
        }
    }
    int nlsbs = 0;
    const int lsb_cap = 2 * ne;
    int cctx = 0;
    for (int k = 0; k < q.lastnz_trunc; k += 2) {
        int t = cctx + q.rate_flag + (k > ne / 2 ? 256 : 0);
        const int q0 = xq[k * XS], q1 = xq[(k + 1) * XS];
        uint32_t a = (uint32_t)(q0 < 0 ? -q0 : q0) & 0xffffu, a_lsb = a;
        uint32_t b = (uint32_t)(q1 < 0 ? -q1 : q1) & 0xffffu, b_lsb = b;
        int lev = 0;
        uint32_t lsb0 = 0, lsb1 = 0;
        while ((a > b ? a : b) >= 4) {
            const int pki = LC3T_AC_SPEC_LOOKUP[t + (lev < 3 ? lev : 3) * 1024];
            ac_encode(st, w, LC3T_AC_SPEC_CUMFREQ[pki][16], LC3T_AC_SPEC_FREQ[pki][16]);
            if (q.lsb_mode && lev == 0) { lsb0 = a & 1; lsb1 = b & 1; }
            else { w.bool_backward((a & 1) == 1); w.bool_backward((b & 1) == 1); }
            a >>= 1;
            b >>= 1;
            lev++;
        }
        const int pki = LC3T_AC_SPEC_LOOKUP[t + (lev < 3 ? lev : 3) * 1024];
        const int sym = (int)(a + 4 * b);
        ac_encode(st, w, LC3T_AC_SPEC_CUMFREQ[pki][sym], LC3T_AC_SPEC_FREQ[pki][sym]);
        if (q.lsb_mode && lev > 0) {
            a_lsb >>= 1;
            b_lsb >>= 1;
            if (nlsbs < lsb_cap) lsbs[nlsbs] = (uint8_t)lsb0;
            nlsbs++;
            if (a_lsb == 0 && q0 != 0) { if (nlsbs < lsb_cap) lsbs[nlsbs] = q0 > 0 ? 0 : 1; nlsbs++; }
            if (nlsbs < lsb_cap) lsbs[nlsbs] = (uint8_t)lsb1;
            nlsbs++;
            if (b_lsb == 0 && q1 != 0) { if (nlsbs < lsb_cap) lsbs[nlsbs] = q1 > 0 ? 0 : 1; nlsbs++; }
        }
        if (a_lsb > 0) w.bool_backward(q0 <= 0);
        if (b_lsb > 0) w.bool_backward(q1 <= 0);
        const int l = lev < 3 ? lev : 3;
        t = l <= 1 ? 1 + (int)(a + b) * (l + 1) : 12 + l;
        cctx = (cctx & 15) * 16 + t;
    }
    const int nbits_side = w.nbits_side_written(nbits);
    int nbits_ari = w.bp * 8;
    nbits_ari += 25 - (31 - __clz(st.range));
    nbits_ari += 8;                                   // QUIRK: `carry >= 0` is always true (:67)
    if (st.carry_count > 0) nbits_ari += st.carry_count * 8;
    int nres_enc = nbits - (nbits_side + nbits_ari);
    if (nres_enc < 0) nres_enc = 0;
    if (!q.lsb_mode) {
        for (int i = 0; i < n_res && i < nres_enc; i++) w.bool_backward((res_bits[i >> 5] >> (i & 31)) & 1u);
    } else {
        const int m = nres_enc < nlsbs ? nres_enc : nlsbs;
        for (int i = 0; i < m; i++) w.bool_backward(lsbs[i] == 1);
    }
    int bits = 1;                                     // ac_enc_finish :354-395
    while ((st.range >> (24 - bits)) == 0) bits++;
    uint32_t mask = 0x00ffffffu >> bits;
    uint32_t val = st.low + mask;
    const uint32_t over1 = val >> 24;
    const uint32_t high = st.low + st.range;
    const uint32_t over2 = high >> 24;
    val &= 0x00ffffffu & ~mask;
    if (over1 == over2) {
        if ((val + mask) >= high) {
            bits += 1;
            mask >>= 1;
            val = ((st.low + mask) & 0x00ffffffu) & ~mask;
        }
        if (val < st.low) st.carry = 1;
    }
    st.low = val;
    while (bits > 0) { ac_shift(st, w); bits -= 8; }
    bits += 8;
    if (st.carry_count > 0) {
        w.byte_forward((uint32_t)st.cache & 0xff);
        while (st.carry_count > 1) { w.byte_forward(0xff); st.carry_count -= 1; }
        w.uint_forward(0xffu >> (8 - bits), bits);
    } else {
        w.uint_forward((uint32_t)st.cache, bits);
    }
}

__global__ void __launch_bounds__(QNT_THREADS) enc_quant_kernel(QuantParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const EncConfig& c = *p.cfg;
    const int tid = threadIdx.x;
    const int stream0 = blockIdx.x * QNT_THREADS;
    const int stream = stream0 + tid;
    const int ne = c.ne;
    float* s_xf = (float*)smem;                                   // [ne][XS]
    float* s_e = s_xf + ne * XS;                                  // [100][XS]
    int16_t* s_xq = (int16_t*)(s_e + 100 * XS);                   // [ne][XS]
    uint8_t* s_rows = (uint8_t*)(s_xq + ne * XS + (ne & 1));      // [QNT_THREADS][row_pitch], 4-byte aligned
    uint8_t* row = s_rows + (size_t)tid * p.row_pitch;
    const int n_rows = min(QNT_THREADS, p.n_streams - stream0);
    // transposed load of the CTA's stream-major spectra: row r is read coalesced, lane = k
    for (int r = 0; r < n_rows; r++) {
        const float* src = p.xf + (size_t)(stream0 + r) * ne;
        for (int k = tid; k < ne; k += QNT_THREADS) s_xf[k * XS + r] = src[k];
    }
    __syncthreads();
    if (stream < p.n_streams) {
        const int nbytes = p.nbytes, nbits = nbytes * 8;
        float* xf = s_xf + tid;
        const float* e_b = p.e_b + (size_t)stream * 64;
        const int32_t* eh = p.ehand + (size_t)stream * EH_WORDS;
        int32_t* es = p.estate + (size_t)stream * ES_WORDS;
        int16_t* xq = s_xq + tid;
        float* e4 = s_e + tid;
        uint8_t* lsbs = p.lsbs + (size_t)stream * 2 * ne;

        const BwRes bw = bandwidth_detect(c, e_b);
        const SnsRes sns = sns_encode(c, xf, e_b, eh[EH_ATTACK] != 0);
        TnsRes tns;
        tns_encode(c, xf, bw.bw, nbits, eh[EH_NEAR_NYQUIST] != 0, tns);
        const QRes q = spectral_quantization(c, es, xf, xq, e4, nbits, bw.nbits, tns.nbits_tns, eh[EH_NBITS_LTPF]);
        // residual_spectrum.rs:33-62
        uint32_t res_bits[13];
        for (int i = 0; i < 13; i++) res_bits[i] = 0;
        int n_res = 0;
        {
            int mx = q.nbits_spec - q.nbits_trunc + 4;
            if (mx < 0) mx = 0;
            if (mx > 0) {
                for (int k = 0; k < ne; k++) {
                    if (n_res >= mx) break;
                    const int v = xq[k * XS];
                    if (v != 0) {
                        if (n_res >= 400) break;
                        if (xf[k * XS] >= (float)v * q.gg) res_bits[n_res >> 5] |= 1u << (n_res & 31);
                        n_res++;
                    }
                }
            }
        }
        const int nff = noise_factor(c, xf, xq, bw.bw, q.gg);
        bitstream_encode(c, bw, sns, tns, eh[EH_PITCH_PRESENT], eh[EH_LTPF_ACTIVE], eh[EH_PITCH_INDEX], q, res_bits, n_res,
                         nff, xq, lsbs, row, nbytes);
    }
    __syncthreads();
    for (int i = tid; i < n_rows * p.nbytes; i += QNT_THREADS) {
        const int r = i / p.nbytes, b = i - r * p.nbytes;
        p.frames_out[(size_t)(stream0 + r) * p.frame_stride + b] = s_rows[(size_t)r * p.row_pitch + b];
    }
    if (p.debug) {                                                // test hook: make the intermediates readable
        for (int r = 0; r < n_rows; r++) {
            for (int k = tid; k < ne; k += QNT_THREADS) {
                p.xf[(size_t)(stream0 + r) * ne + k] = s_xf[k * XS + r];
                p.xq[(size_t)(stream0 + r) * ne + k] = s_xq[k * XS + r];
            }
        }
    }
}

cudaError_t launch_enc_quant(const EncoderState& st, uint8_t* frames_out, int nbytes, size_t frame_stride, cudaStream_t stream) {
    const int ne = st.cfg.ne;
    QuantParams p;
    p.cfg = st.ecfg;
    p.n_streams = st.n_streams;
    p.nbytes = nbytes;
    p.frame_stride = frame_stride;
    p.xf = st.xf;
    p.e_b = st.e_b;
    p.ehand = st.ehand;
    p.estate = st.estate;
    p.xq = st.xq;
    p.scratch_e = st.scratch_e;
    p.lsbs = st.lsbs;
    p.frames_out = frames_out;
    int words = (nbytes + 3) / 4;
    if ((words & 1) == 0) words++;
    p.row_pitch = words * 4;
    p.debug = st.debug;
    const size_t smem = (size_t)ne * XS * 4 + 100 * XS * 4 + ((size_t)ne * XS + (ne & 1)) * 2 + (size_t)QNT_THREADS * p.row_pitch;
    cudaError_t e = cudaFuncSetAttribute(enc_quant_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    enc_quant_kernel<<<(st.n_streams + QNT_THREADS - 1) / QNT_THREADS, QNT_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lc3b
