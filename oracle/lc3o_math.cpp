// CPU ORACLE - TEST INFRASTRUCTURE (see lc3o.h).
//
// f32 transcendentals as the reference gets them: num-traits 0.2 with
// `features=["libm"]` (Cargo.toml:17) forwards `Real::{powf,log2,log10,exp2,asin,sin}`
// to the `libm` crate 0.2.x, whose f32 routines are ports of FreeBSD msun
// (e_powf.c, e_log2f.c, e_log10f.c, s_exp2f.c, e_asinf.c, s_sinf.c).  The `libm`
// crate's source is NOT under /root/reference (third-party dependency), so these
// are restatements of the published msun algorithms; every constant is given by
// its IEEE bit pattern where msun documents one.  Pinned by the reference's golden
// vectors: global_gain_decode (61.0540199), spectral_quantization_run
// (gg == 24.7091141, a value a correctly-rounded powf does NOT return), sns_run
// (powf/log2/exp2 over 64 bands) and temporal_noise_shaping_run (asin/sin).
#include "lc3o.h"

namespace lc3o {

static inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float bitsf(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline double bitsd(uint64_t u) { double f; std::memcpy(&f, &u, 8); return f; }

static float scalbnf_(float x, int n) {
    // msun s_scalbnf.c semantics for the finite range powf can reach
    double r = std::ldexp((double)x, n);
    return (float)r;
}

// ------------------------------------------------------------------ powf (e_powf.c)
float msun_powf(float x, float y) {
    static const float bp[2] = {1.0f, 1.5f};
    static const float dp_h[2] = {0.0f, bitsf(0x3f15c000)};   // 5.84960938e-01
    static const float dp_l[2] = {0.0f, bitsf(0x35d1cfdc)};   // 1.56322085e-06
    const float two24 = 16777216.0f, huge = 1.0e30f, tiny = 1.0e-30f;
    const float L1 = bitsf(0x3f19999a), L2 = bitsf(0x3edb6db7), L3 = bitsf(0x3eaaaaab),
                L4 = bitsf(0x3e8ba305), L5 = bitsf(0x3e6c3255), L6 = bitsf(0x3e53f142);
    const float P1 = bitsf(0x3e2aaaab), P2 = bitsf(0xbb360b61), P3 = bitsf(0x388ab355),
                P4 = bitsf(0xb5ddea0e), P5 = bitsf(0x3331bb4c);
    const float lg2 = bitsf(0x3f317218), lg2_h = bitsf(0x3f317200), lg2_l = bitsf(0x35bfbe8c);
    const float ovt = 4.2995665694e-08f;
    const float cp = bitsf(0x3f76384f), cp_h = bitsf(0x3f764000), cp_l = bitsf(0xb8f623c6);
    const float ivln2 = bitsf(0x3fb8aa3b), ivln2_h = bitsf(0x3fb8aa00), ivln2_l = bitsf(0x36eca570);

    float z, ax, z_h, z_l, p_h, p_l;
    float y1, t1, t2, r, s, sn, t, u, v, w;
    int32_t i, j, k, yisint, n;
    int32_t hx = (int32_t)fbits(x), hy = (int32_t)fbits(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    int32_t is;

    if (iy == 0) return 1.0f;                       // x**0 = 1
    if (hx == 0x3f800000) return 1.0f;              // 1**y = 1
    if (ix > 0x7f800000 || iy > 0x7f800000) return x + y;   // NaN

    yisint = 0;
    if (hx < 0) {
        if (iy >= 0x4b800000) yisint = 2;
        else if (iy >= 0x3f800000) {
            k = (iy >> 23) - 0x7f;
            j = iy >> (23 - k);
            if ((j << (23 - k)) == iy) yisint = 2 - (j & 1);
        }
    }
    if (iy == 0x7f800000) {
        if (ix == 0x3f800000) return 1.0f;
        else if (ix > 0x3f800000) return (hy >= 0) ? y : 0.0f;
        else return (hy >= 0) ? 0.0f : -y;
    }
    if (iy == 0x3f800000) return (hy >= 0) ? x : 1.0f / x;
    if (hy == 0x40000000) return x * x;
    if (hy == 0x3f000000 && hx >= 0) return std::sqrt(x);

    ax = std::fabs(x);
    if (ix == 0x7f800000 || ix == 0 || ix == 0x3f800000) {
        z = ax;
        if (hy < 0) z = 1.0f / z;
        if (hx < 0) {
            if (((ix - 0x3f800000) | yisint) == 0) z = (z - z) / (z - z);
            else if (yisint == 1) z = -z;
        }
        return z;
    }
    sn = 1.0f;
    if (hx < 0) {
        if (yisint == 0) return (x - x) / (x - x);
        if (yisint == 1) sn = -1.0f;
    }

    if (iy > 0x4d000000) {                          // |y| > 2**27
        if (ix < 0x3f7ffff8) return (hy < 0) ? sn * huge * huge : sn * tiny * tiny;
        if (ix > 0x3f800007) return (hy > 0) ? sn * huge * huge : sn * tiny * tiny;
        t = ax - 1.0f;
        w = (t * t) * (0.5f - t * (0.333333333333f - t * 0.25f));
        u = ivln2_h * t;
        v = t * ivln2_l - w * ivln2;
        t1 = u + v;
        is = (int32_t)fbits(t1);
        t1 = bitsf((uint32_t)is & 0xfffff000u);
        t2 = v - (t1 - u);
    } else {
        float s2, s_h, s_l, t_h, t_l;
        n = 0;
        if (ix < 0x00800000) { ax *= two24; n -= 24; ix = (int32_t)fbits(ax); }
        n += ((ix) >> 23) - 0x7f;
        j = ix & 0x007fffff;
        ix = j | 0x3f800000;
        if (j <= 0x1cc471) k = 0;
        else if (j < 0x5db3d7) k = 1;
        else { k = 0; n += 1; ix -= 0x00800000; }
        ax = bitsf((uint32_t)ix);

        u = ax - bp[k];
        v = 1.0f / (ax + bp[k]);
        s = u * v;
        s_h = s;
        is = (int32_t)fbits(s_h);
        s_h = bitsf((uint32_t)is & 0xfffff000u);
        is = (int32_t)((((uint32_t)ix >> 1) & 0xfffff000u) | 0x20000000u);
        t_h = bitsf((uint32_t)is + 0x00400000u + ((uint32_t)k << 21));
        t_l = ax - (t_h - bp[k]);
        s_l = v * ((u - s_h * t_h) - s_h * t_l);
        s2 = s * s;
        r = s2 * s2 * (L1 + s2 * (L2 + s2 * (L3 + s2 * (L4 + s2 * (L5 + s2 * L6)))));
        r += s_l * (s_h + s);
        s2 = s_h * s_h;
        t_h = 3.0f + s2 + r;
        is = (int32_t)fbits(t_h);
        t_h = bitsf((uint32_t)is & 0xfffff000u);
        t_l = r - ((t_h - 3.0f) - s2);
        u = s_h * t_h;
        v = s_l * t_h + t_l * s;
        p_h = u + v;
        is = (int32_t)fbits(p_h);
        p_h = bitsf((uint32_t)is & 0xfffff000u);
        p_l = v - (p_h - u);
        z_h = cp_h * p_h;
        z_l = cp_l * p_h + p_l * cp + dp_l[k];
        t = (float)n;
        t1 = (((z_h + z_l) + dp_h[k]) + t);
        is = (int32_t)fbits(t1);
        t1 = bitsf((uint32_t)is & 0xfffff000u);
        t2 = z_l - (((t1 - t) - dp_h[k]) - z_h);
    }

    is = (int32_t)fbits(y);
    y1 = bitsf((uint32_t)is & 0xfffff000u);
    p_l = (y - y1) * t1 + y * t2;
    p_h = y1 * t1;
    z = p_l + p_h;
    j = (int32_t)fbits(z);
    if (j > 0x43000000) return sn * huge * huge;
    else if (j == 0x43000000) { if (p_l + ovt > z - p_h) return sn * huge * huge; }
    else if ((j & 0x7fffffff) > 0x43160000) return sn * tiny * tiny;
    else if ((uint32_t)j == 0xc3160000u) { if (p_l <= z - p_h) return sn * tiny * tiny; }

    i = j & 0x7fffffff;
    k = (i >> 23) - 0x7f;
    n = 0;
    if (i > 0x3f000000) {
        n = j + (0x00800000 >> (k + 1));
        k = ((n & 0x7fffffff) >> 23) - 0x7f;
        t = bitsf((uint32_t)n & ~(0x007fffffu >> k));
        n = ((n & 0x007fffff) | 0x00800000) >> (23 - k);
        if (j < 0) n = -n;
        p_h -= t;
    }
    t = p_l + p_h;
    is = (int32_t)fbits(t);
    t = bitsf((uint32_t)is & 0xffff8000u);
    u = t * lg2_h;
    v = (p_l - (t - p_h)) * lg2 + t * lg2_l;
    z = u + v;
    w = v - (z - u);
    t = z * z;
    t1 = z - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    r = (z * t1) / (t1 - 2.0f) - (w + z * w);
    z = 1.0f - (r - z);
    j = (int32_t)fbits(z);
    j += n << 23;
    if ((j >> 23) <= 0) z = scalbnf_(z, n);
    else z = bitsf((uint32_t)j);
    return sn * z;
}

// ------------------------------------------------------------------ log2f / log10f (e_log2f.c, e_log10f.c)
static inline bool log_reduce(float& x, int32_t& k, float* early) {
    uint32_t ix = fbits(x);
    k = 0;
    if (ix < 0x00800000u || (ix >> 31)) {
        if ((ix << 1) == 0) { *early = -1.0f / (x * x); return true; }
        if (ix >> 31) { *early = (x - x) / 0.0f; return true; }
        k -= 25;
        x *= bitsf(0x4c000000);
        ix = fbits(x);
    } else if (ix >= 0x7f800000u) { *early = x; return true; }
    else if (ix == 0x3f800000u) { *early = 0.0f; return true; }
    ix += 0x3f800000u - 0x3f3504f3u;
    k += (int32_t)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    x = bitsf(ix);
    return false;
}

static inline void log_kernel(float x, float* hi_out, float* lo_out) {
    const float Lg1 = 0xaaaaaa.0p-24f, Lg2 = 0xccce13.0p-25f, Lg3 = 0x91e9ee.0p-25f, Lg4 = 0xf89e26.0p-26f;
    float f = x - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float w = z * z;
    float t1 = w * (Lg2 + w * Lg4);
    float t2 = z * (Lg1 + w * Lg3);
    float R = t2 + t1;
    float hfsq = 0.5f * f * f;
    float hi = f - hfsq;
    hi = bitsf(fbits(hi) & 0xfffff000u);
    float lo = (f - hi) - hfsq + s * (hfsq + R);
    *hi_out = hi;
    *lo_out = lo;
}

float msun_log2f(float x) {
    const float ivln2hi = bitsf(0x3fb8b000);   //  1.4428710938e+00
    const float ivln2lo = bitsf(0xb9389ad4);   // -1.7605285393e-04
    int32_t k;
    float early;
    if (log_reduce(x, k, &early)) return early;
    float hi, lo;
    log_kernel(x, &hi, &lo);
    return (lo + hi) * ivln2lo + lo * ivln2hi + hi * ivln2hi + (float)k;
}

float msun_log10f(float x) {
    const float ivln10hi = bitsf(0x3ede6000);   //  4.3432617188e-01
    const float ivln10lo = bitsf(0xb804ead9);   // -3.1689971365e-05
    const float log10_2hi = bitsf(0x3e9a2080);  //  3.0102920532e-01
    const float log10_2lo = bitsf(0x355427db);  //  7.9034151668e-07
    int32_t k;
    float early;
    if (log_reduce(x, k, &early)) return early;
    float hi, lo;
    log_kernel(x, &hi, &lo);
    float dk = (float)k;
    return dk * log10_2lo + (lo + hi) * ivln10lo + lo * ivln10hi + hi * ivln10hi + dk * log10_2hi;
}

// ------------------------------------------------------------------ exp2f (s_exp2f.c)
float msun_exp2f(float x) {
    static const uint64_t exp2ft[16] = {
        0x3fe6a09e667f3bcdull, 0x3fe7a11473eb0187ull, 0x3fe8ace5422aa0dbull, 0x3fe9c49182a3f090ull,
        0x3feae89f995ad3adull, 0x3fec199bdd85529cull, 0x3fed5818dcfba487ull, 0x3feea4afa2a490daull,
        0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
        0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
    };
    const float redux = 786432.0f;                  // 0x1.8p23f / 16
    const float P1 = 0x1.62e430p-1f, P2 = 0x1.ebfbe0p-3f, P3 = 0x1.c6b348p-5f, P4 = 0x1.3b2c9cp-7f;
    uint32_t ix = fbits(x) & 0x7fffffffu;
    if (ix > 0x42fc0000u) {                         // |x| > 126
        if (ix > 0x7f800000u) return x;             // NaN
        if (fbits(x) >= 0x43000000u && fbits(x) < 0x80000000u) return x * bitsf(0x7f000000);   // overflow
        if (fbits(x) >= 0x80000000u) {
            if (fbits(x) >= 0xc3160000u) return 0.0f;   // x <= -150: underflow
        }
    } else if (ix <= 0x33000000u) {                 // |x| <= 0x1p-25
        return 1.0f + x;
    }
    float uf = x + redux;
    uint32_t i0 = fbits(uf);
    i0 += 16 / 2;
    uint32_t k = i0 / 16;
    double uk = bitsd((uint64_t)(0x3ffu + k) << 52);
    i0 &= 16 - 1;
    uf -= redux;
    double z = (double)(x - uf);
    double r = bitsd(exp2ft[i0]);
    double t = r * z;
    r = r + t * ((double)P1 + z * (double)P2) + t * (z * z) * ((double)P3 + z * (double)P4);
    return (float)(r * uk);
}

// ------------------------------------------------------------------ asinf (e_asinf.c)
static inline float asin_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f,
                qS1 = -7.0662963390e-01f;
    float p = z * (pS0 + z * (pS1 + z * pS2));
    float q = 1.0f + z * qS1;
    return p / q;
}

float msun_asinf(float x) {
    const double pio2 = 1.570796326794896558e+00;
    uint32_t hx = fbits(x), ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) {
        if (ix == 0x3f800000u) return (float)((double)x * pio2 + 7.5231638452626401e-37 /*0x1p-120*/);
        return 0.0f / (x - x);
    }
    if (ix < 0x3f000000u) {
        if (ix < 0x39800000u && ix >= 0x00800000u) return x;
        return x + x * asin_R(x * x);
    }
    float z = (1.0f - std::fabs(x)) * 0.5f;
    double s = std::sqrt((double)z);
    x = (float)(pio2 - 2.0 * (s + s * (double)asin_R(z)));
    return (hx >> 31) ? -x : x;
}

// ------------------------------------------------------------------ sinf (s_sinf.c, k_sinf.c, k_cosf.c)
static inline float k_sindf(double x) {
    const double S1 = -0x15555554cbac77.0p-55, S2 = 0x111110896efbb2.0p-59,
                 S3 = -0x1a00f9e2cae774.0p-65, S4 = 0x16cd878c3b46a7.0p-71;
    double z = x * x, w = z * z, r = S3 + z * S4, s = z * x;
    return (float)((x + s * (S1 + z * S2)) + s * w * r);
}
static inline float k_cosdf(double x) {
    const double C0 = -0x1ffffffd0c5e81.0p-54, C1 = 0x155553e1053a42.0p-57,
                 C2 = -0x16c087e80f1e27.0p-62, C3 = 0x199342e0ee5069.0p-68;
    double z = x * x, w = z * z, r = C2 + z * C3;
    return (float)(((1.0 + z * C0) + w * C1) + (w * z) * r);
}

float msun_sinf(float x) {
    const double s1pio2 = 1 * M_PI_2, s2pio2 = 2 * M_PI_2, s3pio2 = 3 * M_PI_2, s4pio2 = 4 * M_PI_2;
    uint32_t hx = fbits(x), ix = hx & 0x7fffffffu;
    bool sign = hx >> 31;
    if (ix <= 0x3f490fdau) {                        // |x| ~<= pi/4
        if (ix < 0x39800000u) return x;             // |x| < 2**-12
        return k_sindf((double)x);
    }
    if (ix <= 0x407b53d1u) {                        // |x| ~<= 5*pi/4
        if (ix <= 0x4016cbe3u) {                    // |x| ~<= 3pi/4
            if (sign) return -k_cosdf((double)x + s1pio2);
            return k_cosdf((double)x - s1pio2);
        }
        return k_sindf(sign ? -((double)x + s2pio2) : -((double)x - s2pio2));
    }
    if (ix <= 0x40e231d5u) {                        // |x| ~<= 9*pi/4
        if (ix <= 0x40afeddfu) {                    // |x| ~<= 7*pi/4
            if (sign) return k_cosdf((double)x + s3pio2);
            return -k_cosdf((double)x - s3pio2);
        }
        return k_sindf(sign ? (double)x + s4pio2 : (double)x - s4pio2);
    }
    // the codec never leaves |x| <= 8*pi/17; general reduction is not needed for parity
    return (float)std::sin((double)x);
}

// ------------------------------------------------------------------ powi (num-traits, no_std)
// num_traits::float::FloatCore::powi default: negative exponent -> recip() first, then
// num_traits::pow::pow (square-and-multiply).  Call sites: encoder/spectral_noise_shaping.rs:223-224
// ((10.0).powi(-4), (2.0).powi(-32)) and encoder/temporal_noise_shaping.rs:244 (gamma.powi(k)).
float nt_powi(float base, int32_t exp) {
    if (exp < 0) { exp = -exp; base = 1.0f / base; }
    uint32_t e = (uint32_t)exp;
    if (e == 0) return 1.0f;
    while ((e & 1) == 0) { base = base * base; e >>= 1; }
    if (e == 1) return base;
    float acc = base;
    while (e > 1) {
        e >>= 1;
        base = base * base;
        if (e & 1) acc = acc * base;
    }
    return acc;
}

// ------------------------------------------------------------------ fast_math::exp2_raw (fast-math 0.1.1)
// Third-party crate, source not under /root/reference.  Published algorithm: split x*2^23
// into floor (added to the exponent field) and fraction (quadratic fit of 2^f).  Pinned by
// the golden vector decoder/spectral_noise_shaping.rs::spectral_noise_shaping_decode.
float fastmath_exp2_raw(float x) {
    const float C2 = 1.00172476f;
    const float C1 = 0.657636276f * (1.0f / 8388608.0f);
    const float C0 = 0.3371894346f * (1.0f / 8388608.0f) * (1.0f / 8388608.0f);
    int32_t mul = rust_f32_to_i32(8388608.0f * x);
    int32_t floor_ = (int32_t)((uint32_t)mul & 0xFF800000u);
    float frac = (float)(mul - floor_);
    float approx = (C0 * frac + C1) * frac + C2;
    return bitsf(fbits(approx) + (uint32_t)floor_);
}

}  // namespace lc3o
