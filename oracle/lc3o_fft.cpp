// CPU ORACLE - TEST INFRASTRUCTURE (see lc3o.h).
// Config + transform substrate: common/config.rs, common/kissfft.rs, common/dct_iv.rs.
#include "lc3o.h"

namespace lc3o {

// common/config.rs:42-100.  QUIRK: 44.1 kHz shares fs_ind = 4 and every size with 48 kHz.
Config make_config(SamplingFrequency sf, FrameDuration fd) {
    static const int FS_IND[6] = {0, 1, 2, 3, 4, 4};
    static const int FS[6] = {8000, 16000, 24000, 32000, 44100, 48000};
    static const int NF75[6] = {60, 120, 180, 240, 360, 360};
    static const int NF10[6] = {80, 160, 240, 320, 480, 480};
    Config c{};
    c.fs_ind = FS_IND[sf];
    c.fs = FS[sf];
    c.n_ms = fd;
    if (fd == SevenPointFiveMs) {
        c.nf = NF75[sf];
        c.ne = (c.nf == 360) ? 300 : c.nf;
        c.nb = (sf == Hz8000) ? 60 : 64;
        c.z = 7 * c.nf / 30;
    } else {
        c.nf = NF10[sf];
        c.ne = (c.nf == 480) ? 400 : c.nf;
        c.nb = 64;
        c.z = 3 * c.nf / 8;
    }
    return c;
}

// decoder/spectral_noise_shaping.rs:125-142, encoder/modified_dct.rs:39-58
const uint16_t* band_indices(const Config& c) {
    static const uint16_t* T75[5] = {LC3T_I_8000_7P5MS, LC3T_I_16000_7P5MS, LC3T_I_24000_7P5MS,
                                     LC3T_I_32000_7P5MS, LC3T_I_48000_7P5MS};
    static const uint16_t* T10[5] = {LC3T_I_8000_10MS, LC3T_I_16000_10MS, LC3T_I_24000_10MS,
                                     LC3T_I_32000_10MS, LC3T_I_48000_10MS};
    return (c.n_ms == SevenPointFiveMs ? T75 : T10)[c.fs_ind];
}

// decoder/modified_dct.rs:37-56
const float* mdct_window(const Config& c) {
    switch (c.nf) {
        case 60: return LC3T_W_N60_7P5MS;
        case 120: return LC3T_W_N120_7P5MS;
        case 180: return LC3T_W_N180_7P5MS;
        case 360: return LC3T_W_N360_7P5MS;
        case 80: return LC3T_W_N80_10MS;
        case 160: return LC3T_W_N160_10MS;
        case 320: return LC3T_W_N320_10MS;
        case 480: return LC3T_W_N480_10MS;
        case 240: return c.n_ms == SevenPointFiveMs ? LC3T_W_N240_7P5MS : LC3T_W_N240_10MS;
    }
    return nullptr;
}

// ------------------------------------------------------------------ kissfft.rs
// kissfft.rs:17-45 (twiddles in f64 then narrowed) and :47-76 (kf_factor).
void KissFft::init(int n, bool inv) {
    nfft = n;
    inverse = inv;
    twiddle.resize(n);
    for (int i = 0; i < n; i++) {
        double phase = -2.0 * M_PI * (double)i / (double)n;
        if (inv) phase *= -1.0;
        twiddle[i].r = (float)std::cos(phase);
        twiddle[i].i = (float)std::sin(phase);
    }
    std::memset(factors, 0, sizeof(factors));
    int p = 4, i = 0, rem = n;
    float floor_sqrt = std::floor(std::sqrt((float)n));
    for (;;) {
        while (rem % p != 0) {
            if (p == 4) p = 2;
            else if (p == 2) p = 3;
            else p += 2;
            if ((float)p > floor_sqrt) p = rem;
        }
        rem /= p;
        factors[i++] = p;
        factors[i++] = rem;
        if (rem <= 1) break;
    }
}

void KissFft::transform(const Cpx* fin, Cpx* fout) const { work(fout, fin, 1, 1, 0, 0, 0); }

// kissfft.rs:86-131 (recursive decimation in time)
void KissFft::work(Cpx* fout, const Cpx* fin, int fstride, int in_stride, int factor_idx, int fin_idx,
                   int fout_idx) const {
    int p = factors[factor_idx], m = factors[factor_idx + 1];
    factor_idx += 2;
    int begin = fout_idx, end = fout_idx + p * m;
    if (m == 1) {
        int src = fin_idx;
        for (int o = fout_idx; o < end; o++, src += fstride * in_stride) fout[o] = fin[src];
    } else {
        do {
            work(fout, fin, fstride * p, in_stride, factor_idx, fin_idx, fout_idx);
            fin_idx += fstride * in_stride;
            fout_idx += m;
        } while (fout_idx != end);
    }
    Cpx* f = fout + begin;
    switch (p) {
        case 2: bfly2(f, fstride, m); break;
        case 3: bfly3(f, fstride, m); break;
        case 4: bfly4(f, fstride, m); break;
        case 5: bfly5(f, fstride, m); break;
        default: bfly_generic(f, fstride, m, p); break;
    }
}

// kissfft.rs:133-141
void KissFft::bfly2(Cpx* f, int fstride, int m) const {
    Cpx* f2 = f + m;
    for (int i = 0; i < m; i++) {
        Cpx t = cmul(f2[i], twiddle[i * fstride]);
        f2[i] = csub(f[i], t);
        f[i] = cadd(f[i], t);
    }
}

// kissfft.rs:143-173
void KissFft::bfly4(Cpx* f, int fstride, int m) const {
    int m2 = 2 * m, m3 = 3 * m, tw1 = 0, tw2 = 0, tw3 = 0;
    for (int i = 0; i < m; i++) {
        Cpx s0 = cmul(f[i + m], twiddle[tw1]);
        Cpx s1 = cmul(f[i + m2], twiddle[tw2]);
        Cpx s2 = cmul(f[i + m3], twiddle[tw3]);
        Cpx s5 = csub(f[i], s1);
        f[i] = cadd(f[i], s1);
        Cpx s3 = cadd(s0, s2);
        Cpx s4 = csub(s0, s2);
        f[i + m2] = csub(f[i], s3);
        f[i] = cadd(f[i], s3);
        tw1 += fstride;
        tw2 += fstride * 2;
        tw3 += fstride * 3;
        if (inverse) {
            f[i + m] = {s5.r - s4.i, s5.i + s4.r};
            f[i + m3] = {s5.r + s4.i, s5.i - s4.r};
        } else {
            f[i + m] = {s5.r + s4.i, s5.i - s4.r};
            f[i + m3] = {s5.r - s4.i, s5.i + s4.r};
        }
    }
}

// kissfft.rs:175-203
void KissFft::bfly3(Cpx* f, int fstride, int m) const {
    int m2 = 2 * m, tw1 = 0, tw2 = 0;
    Cpx epi3 = twiddle[fstride * m];
    for (int i = 0; i < m; i++) {
        Cpx s1 = cmul(f[i + m], twiddle[tw1]);
        Cpx s2 = cmul(f[i + m2], twiddle[tw2]);
        Cpx s3 = cadd(s1, s2);
        Cpx s0 = csub(s1, s2);
        tw1 += fstride;
        tw2 += fstride * 2;
        Cpx fi = f[i];
        f[i + m] = {fi.r - (s3.r * 0.5f), fi.i - (s3.i * 0.5f)};
        s0.r *= epi3.i;
        s0.i *= epi3.i;
        f[i] = cadd(f[i], s3);
        Cpx fm = f[i + m];
        f[i + m2] = {fm.r + s0.i, fm.i - s0.r};
        f[i + m] = {fm.r - s0.i, fm.i + s0.r};
    }
}

// kissfft.rs:205-256
void KissFft::bfly5(Cpx* f, int fstride, int m) const {
    Cpx ya = twiddle[fstride * m], yb = twiddle[fstride * 2 * m];
    int m1 = m, m2 = 2 * m, m3 = 3 * m, m4 = 4 * m;
    for (int i = 0; i < m; i++) {
        Cpx s0 = f[i];
        Cpx s1 = cmul(f[i + m1], twiddle[i * fstride]);
        Cpx s2 = cmul(f[i + m2], twiddle[i * 2 * fstride]);
        Cpx s3 = cmul(f[i + m3], twiddle[i * 3 * fstride]);
        Cpx s4 = cmul(f[i + m4], twiddle[i * 4 * fstride]);
        Cpx s7 = cadd(s1, s4), s10 = csub(s1, s4), s8 = cadd(s2, s3), s9 = csub(s2, s3);
        f[i].r += s7.r + s8.r;
        f[i].i += s7.i + s8.i;
        Cpx s5 = {s0.r + (s7.r * ya.r) + (s8.r * yb.r), s0.i + (s7.i * ya.r) + (s8.i * yb.r)};
        Cpx s6 = {(s10.i * ya.i) + (s9.i * yb.i), -(s10.r * ya.i) - (s9.r * yb.i)};
        f[i + m1] = csub(s5, s6);
        f[i + m4] = cadd(s5, s6);
        Cpx s11 = {s0.r + (s7.r * yb.r) + (s8.r * ya.r), s0.i + (s7.i * yb.r) + (s8.i * ya.r)};
        Cpx s12 = {-(s10.i * yb.i) + (s9.i * ya.i), (s10.r * yb.i) - (s9.r * ya.i)};
        f[i + m2] = cadd(s11, s12);
        f[i + m3] = csub(s11, s12);
    }
}

// kissfft.rs:258-288.  Never reached for LC3 sizes (every nf/2 is 2^a 3^b 5); kept, as written,
// including its quirk of loading `m` (not `p`) scratch entries.
void KissFft::bfly_generic(Cpx* f, int fstride, int m, int p) const {
    Cpx scratch[480];
    for (int u = 0; u < m; u++) {
        int k = u;
        for (int q1 = 0; q1 < m; q1++) { scratch[q1] = f[k]; k += m; }
        k = u;
        for (int q1 = 0; q1 < p; q1++) {
            int twidx = 0;
            f[k] = scratch[0];
            for (int q = 1; q < p; q++) {
                twidx += fstride * k;
                if (twidx >= nfft) twidx -= nfft;
                Cpx t = cmul(scratch[q], twiddle[twidx]);
                f[k] = cadd(f[k], t);
            }
            k += m;
        }
    }
}

// ------------------------------------------------------------------ dct_iv.rs
// dct_iv.rs:22-47
void DctIv::init(int nf_) {
    nf = nf_;
    int count = nf / 2;
    fft.init(count, false);
    input.assign(count, {0, 0});
    output.assign(count, {0, 0});
    twiddle.resize(count);
    for (int i = 0; i < count; i++) {
        double temp = -M_PI * (double)(8 * i + 1) / (8.0 * (double)count * 2.0);
        twiddle[i] = {(float)std::cos(temp), (float)std::sin(temp)};
    }
}

// dct_iv.rs:49-67
void DctIv::run(float* buf) {
    int count = nf / 2;
    for (int n = 0; n < count; n++) {
        Cpx c = {buf[2 * n], buf[nf - 2 * n - 1]};
        input[n] = cmul(twiddle[n], c);
    }
    fft.transform(input.data(), output.data());
    for (int n = 0; n < count; n++) {
        Cpx c = cmul(twiddle[n], output[n]);
        buf[2 * n] = c.r * 2.0f;
        buf[nf - 2 * n - 1] = -c.i * 2.0f;
    }
}

}  // namespace lc3o
