// lc3b engine: f32 arithmetic with the reference's rounding behaviour, host + device.
//
// The reference computes in f32 without FMA contraction and takes its transcendentals from the
// `libm` crate (FreeBSD msun ports; num-traits features=["libm"], reference Cargo.toml:17).  The
// encoder's bitstream depends on those roundings (byte-identical output is the parity bar), and the
// decoder's global gain uses powf (src/decoder/global_gain.rs:20), so the engine carries its own
// statements of the msun algorithms.  Every multiply/add goes through xm/xa/xs/xd so that nvcc can
// never contract them into FMAs (__f*_rn intrinsics are contraction-proof); host code in this
// translation unit is built with -ffp-contract=off.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LC3B_HD __host__ __device__ __forceinline__
#else
#define LC3B_HD inline
#endif

namespace lc3b {

#if defined(__CUDA_ARCH__)
LC3B_HD float xm(float a, float b) { return __fmul_rn(a, b); }
LC3B_HD float xa(float a, float b) { return __fadd_rn(a, b); }
LC3B_HD float xs(float a, float b) { return __fsub_rn(a, b); }
LC3B_HD float xd(float a, float b) { return __fdiv_rn(a, b); }
LC3B_HD double dm(double a, double b) { return __dmul_rn(a, b); }
LC3B_HD double da(double a, double b) { return __dadd_rn(a, b); }
LC3B_HD double ds(double a, double b) { return __dsub_rn(a, b); }
LC3B_HD uint32_t f2u(float f) { return __float_as_uint(f); }
LC3B_HD float u2f(uint32_t u) { return __uint_as_float(u); }
LC3B_HD double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
LC3B_HD float d2f(double d) { return __double2float_rn(d); }
#else
LC3B_HD float xm(float a, float b) { return a * b; }
LC3B_HD float xa(float a, float b) { return a + b; }
LC3B_HD float xs(float a, float b) { return a - b; }
LC3B_HD float xd(float a, float b) { return a / b; }
LC3B_HD double dm(double a, double b) { return a * b; }
LC3B_HD double da(double a, double b) { return a + b; }
LC3B_HD double ds(double a, double b) { return a - b; }
LC3B_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
LC3B_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
LC3B_HD double u2d(uint64_t u) { double f; memcpy(&f, &u, 8); return f; }
LC3B_HD float d2f(double d) { return (float)d; }
#endif

LC3B_HD float trunc12(float v) { return u2f(f2u(v) & 0xfffff000u); }

// Rust `as` casts: saturating, NaN -> 0
// On the device these are single conversions: cvt.rzi.s32.f32 clamps to the integer range and turns NaN into 0 by
// definition (PTX ISA, "cvt": float-to-integer conversions saturate; NaN -> 0), cvt.rzi.s16.f32 likewise.
LC3B_HD int32_t cast_i32(float x) {
#ifdef __CUDA_ARCH__
    return __float2int_rz(x);
#else
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int32_t)x;
#endif
}
LC3B_HD int16_t cast_i16(float x) {
#ifdef __CUDA_ARCH__
    short r;
    asm("cvt.rzi.s16.f32 %0, %1;" : "=h"(r) : "f"(x));
    return (int16_t)r;
#else
    if (x != x) return 0;
    if (x >= 32767.0f) return INT16_MAX;
    if (x <= -32768.0f) return INT16_MIN;
    return (int16_t)x;
#endif
}
LC3B_HD float maxf_rs(float a, float b) { return (a != a) ? b : (b != b) ? a : (a > b ? a : b); }
LC3B_HD float minf_rs(float a, float b) { return (a != a) ? b : (b != b) ? a : (a < b ? a : b); }

// ---------------------------------------------------------------- powf, positive finite base, finite |y| < 2^27
// msun e_powf.c main path (the codec only ever raises 10.0 to moderate exponents); the special-case
// prologue is reduced to what those call sites can reach.
LC3B_HD float powf_msun(float x, float y) {
    const float L1 = u2f(0x3f19999a), L2 = u2f(0x3edb6db7), L3 = u2f(0x3eaaaaab), L4 = u2f(0x3e8ba305),
                L5 = u2f(0x3e6c3255), L6 = u2f(0x3e53f142);
    const float P1 = u2f(0x3e2aaaab), P2 = u2f(0xbb360b61), P3 = u2f(0x388ab355), P4 = u2f(0xb5ddea0e),
                P5 = u2f(0x3331bb4c);
    const float lg2 = u2f(0x3f317218), lg2_h = u2f(0x3f317200), lg2_l = u2f(0x35bfbe8c);
    const float cp = u2f(0x3f76384f), cp_h = u2f(0x3f764000), cp_l = u2f(0xb8f623c6);
    const float ovt = 4.2995665694e-08f;
    int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (iy == 0) return 1.0f;
    if (hx == 0x3f800000) return 1.0f;
    if (iy == 0x3f800000) return hy >= 0 ? x : xd(1.0f, x);
    if (hy == 0x40000000) return xm(x, x);
    if (hy == 0x3f000000 && hx >= 0) return sqrtf(x);
    float ax = fabsf(x);
    int32_t n = 0, j, k;
    if (ix < 0x00800000) { ax = xm(ax, 16777216.0f); n -= 24; ix = (int32_t)f2u(ax); }
    n += (ix >> 23) - 0x7f;
    j = ix & 0x007fffff;
    ix = j | 0x3f800000;
    if (j <= 0x1cc471) k = 0;
    else if (j < 0x5db3d7) k = 1;
    else { k = 0; n += 1; ix -= 0x00800000; }
    ax = u2f((uint32_t)ix);
    const float bpk = k ? 1.5f : 1.0f;
    const float dp_hk = k ? u2f(0x3f15c000) : 0.0f, dp_lk = k ? u2f(0x35d1cfdc) : 0.0f;
    float u = xs(ax, bpk);
    float v = xd(1.0f, xa(ax, bpk));
    float s = xm(u, v);
    float s_h = trunc12(s);
    uint32_t is = (((uint32_t)ix >> 1) & 0xfffff000u) | 0x20000000u;
    float t_h = u2f(is + 0x00400000u + ((uint32_t)k << 21));
    float t_l = xs(ax, xs(t_h, bpk));
    float s_l = xm(v, xs(xs(u, xm(s_h, t_h)), xm(s_h, t_l)));
    float s2 = xm(s, s);
    float r = xm(xm(s2, s2),
                 xa(L1, xm(s2, xa(L2, xm(s2, xa(L3, xm(s2, xa(L4, xm(s2, xa(L5, xm(s2, L6)))))))))));
    r = xa(r, xm(s_l, xa(s_h, s)));
    s2 = xm(s_h, s_h);
    t_h = trunc12(xa(xa(3.0f, s2), r));
    t_l = xs(r, xs(xs(t_h, 3.0f), s2));
    u = xm(s_h, t_h);
    v = xa(xm(s_l, t_h), xm(t_l, s));
    float p_h = trunc12(xa(u, v));
    float p_l = xs(v, xs(p_h, u));
    float z_h = xm(cp_h, p_h);
    float z_l = xa(xa(xm(cp_l, p_h), xm(p_l, cp)), dp_lk);
    float t = (float)n;
    float t1 = trunc12(xa(xa(xa(z_h, z_l), dp_hk), t));
    float t2 = xs(z_l, xs(xs(xs(t1, t), dp_hk), z_h));
    float y1 = trunc12(y);
    p_l = xa(xm(xs(y, y1), t1), xm(y, t2));
    p_h = xm(y1, t1);
    float z = xa(p_l, p_h);
    j = (int32_t)f2u(z);
    const float huge = 1.0e30f, tiny = 1.0e-30f;
    if (j > 0x43000000) return xm(huge, huge);
    else if (j == 0x43000000) { if (xa(p_l, ovt) > xs(z, p_h)) return xm(huge, huge); }
    else if ((j & 0x7fffffff) > 0x43160000) return xm(tiny, tiny);
    else if ((uint32_t)j == 0xc3160000u) { if (p_l <= xs(z, p_h)) return xm(tiny, tiny); }
    int32_t i = j & 0x7fffffff;
    k = (i >> 23) - 0x7f;
    n = 0;
    if (i > 0x3f000000) {
        n = j + (0x00800000 >> (k + 1));
        k = ((n & 0x7fffffff) >> 23) - 0x7f;
        t = u2f((uint32_t)n & ~(0x007fffffu >> k));
        n = ((n & 0x007fffff) | 0x00800000) >> (23 - k);
        if (j < 0) n = -n;
        p_h = xs(p_h, t);
    }
    t = u2f(f2u(xa(p_l, p_h)) & 0xffff8000u);
    u = xm(t, lg2_h);
    v = xa(xm(xs(p_l, xs(t, p_h)), lg2), xm(t, lg2_l));
    z = xa(u, v);
    float w = xs(v, xs(z, u));
    t = xm(z, z);
    t1 = xs(z, xm(t, xa(P1, xm(t, xa(P2, xm(t, xa(P3, xm(t, xa(P4, xm(t, P5))))))))));
    r = xs(xd(xm(z, t1), xs(t1, 2.0f)), xa(w, xm(z, w)));
    z = xs(1.0f, xs(r, z));
    j = (int32_t)f2u(z);
    j += n << 23;
    if ((j >> 23) <= 0) z = ldexpf(z, n);
    else z = u2f((uint32_t)j);
    return z;
}

// ---------------------------------------------------------------- log2f / log10f for finite x > 0 (e_log2f.c, e_log10f.c)
struct LogParts { float hi, lo; int32_t k; };
LC3B_HD LogParts log_parts(float x) {
    const float Lg1 = 0xaaaaaa.0p-24f, Lg2 = 0xccce13.0p-25f, Lg3 = 0x91e9ee.0p-25f, Lg4 = 0xf89e26.0p-26f;
    uint32_t ix = f2u(x);
    int32_t k = 0;
    if (ix < 0x00800000u) { k -= 25; x = xm(x, u2f(0x4c000000)); ix = f2u(x); }
    ix += 0x3f800000u - 0x3f3504f3u;
    k += (int32_t)(ix >> 23) - 0x7f;
    ix = (ix & 0x007fffffu) + 0x3f3504f3u;
    x = u2f(ix);
    float f = xs(x, 1.0f);
    float s = xd(f, xa(2.0f, f));
    float z = xm(s, s);
    float w = xm(z, z);
    float t1 = xm(w, xa(Lg2, xm(w, Lg4)));
    float t2 = xm(z, xa(Lg1, xm(w, Lg3)));
    float R = xa(t2, t1);
    float hfsq = xm(xm(0.5f, f), f);
    float hi = trunc12(xs(f, hfsq));
    float lo = xa(xs(xs(f, hi), hfsq), xm(s, xa(hfsq, R)));
    return {hi, lo, k};
}
LC3B_HD float log2f_msun(float x) {
    if (f2u(x) == 0x3f800000u) return 0.0f;
    const float ivln2hi = u2f(0x3fb8b000), ivln2lo = u2f(0xb9389ad4);
    LogParts p = log_parts(x);
    return xa(xa(xa(xm(xa(p.lo, p.hi), ivln2lo), xm(p.lo, ivln2hi)), xm(p.hi, ivln2hi)), (float)p.k);
}
LC3B_HD float log10f_msun(float x) {
    if (f2u(x) == 0x3f800000u) return 0.0f;
    const float ivln10hi = u2f(0x3ede6000), ivln10lo = u2f(0xb804ead9), log10_2hi = u2f(0x3e9a2080),
                log10_2lo = u2f(0x355427db);
    LogParts p = log_parts(x);
    float dk = (float)p.k;
    return xa(xa(xa(xa(xm(dk, log10_2lo), xm(xa(p.lo, p.hi), ivln10lo)), xm(p.lo, ivln10hi)), xm(p.hi, ivln10hi)),
              xm(dk, log10_2hi));
}

// ---------------------------------------------------------------- exp2f for |x| < 126 (s_exp2f.c)
#define LC3B_EXP2FT_INIT {                                                                          \
        0x3fe6a09e667f3bcdull, 0x3fe7a11473eb0187ull, 0x3fe8ace5422aa0dbull, 0x3fe9c49182a3f090ull,     \
        0x3feae89f995ad3adull, 0x3fec199bdd85529cull, 0x3fed5818dcfba487ull, 0x3feea4afa2a490daull,     \
        0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,     \
        0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull}
static const uint64_t EXP2FT_HOST[16] = LC3B_EXP2FT_INIT;
#if defined(__CUDACC__)
static __device__ const uint64_t EXP2FT_DEV[16] = LC3B_EXP2FT_INIT;
#endif
LC3B_HD float exp2f_msun(float x) {
#if defined(__CUDA_ARCH__)
    const uint64_t* T = EXP2FT_DEV;
#else
    const uint64_t* T = EXP2FT_HOST;
#endif
    const float redux = 786432.0f;
    const double P1 = (double)0x1.62e430p-1f, P2 = (double)0x1.ebfbe0p-3f, P3 = (double)0x1.c6b348p-5f,
                 P4 = (double)0x1.3b2c9cp-7f;
    uint32_t ix = f2u(x) & 0x7fffffffu;
    if (ix <= 0x33000000u) return xa(1.0f, x);
    float uf = xa(x, redux);
    uint32_t i0 = f2u(uf) + 8u;
    uint32_t k = i0 / 16u;
    double uk = u2d((uint64_t)(0x3ffu + k) << 52);
    i0 &= 15u;
    uf = xs(uf, redux);
    double z = (double)xs(x, uf);
    double r = u2d(T[i0]);
    double t = dm(r, z);
    r = da(da(r, dm(t, da(P1, dm(z, P2)))), dm(dm(t, dm(z, z)), da(P3, dm(z, P4))));
    return d2f(dm(r, uk));
}

// ---------------------------------------------------------------- asinf for |x| <= 1 (e_asinf.c)
LC3B_HD float asin_R(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f, qS1 = -7.0662963390e-01f;
    float p = xm(z, xa(pS0, xm(z, xa(pS1, xm(z, pS2)))));
    float q = xa(1.0f, xm(z, qS1));
    return xd(p, q);
}
LC3B_HD float asinf_msun(float x) {
    const double pio2 = 1.570796326794896558e+00;
    uint32_t hx = f2u(x), ix = hx & 0x7fffffffu;
    if (ix >= 0x3f800000u) {
        if (ix == 0x3f800000u) return (float)da(dm((double)x, pio2), 7.5231638452626401e-37);
        return xd(0.0f, xs(x, x));
    }
    if (ix < 0x3f000000u) {
        if (ix < 0x39800000u && ix >= 0x00800000u) return x;
        return xa(x, xm(x, asin_R(xm(x, x))));
    }
    float z = xm(xs(1.0f, fabsf(x)), 0.5f);
    double s = sqrt((double)z);
    x = (float)ds(pio2, dm(2.0, da(s, dm(s, (double)asin_R(z)))));
    return (hx >> 31) ? -x : x;
}

// ---------------------------------------------------------------- num-traits powi (no_std): recip, square-and-multiply
LC3B_HD float powi_nt(float base, int32_t exp) {
    if (exp < 0) { exp = -exp; base = xd(1.0f, base); }
    uint32_t e = (uint32_t)exp;
    if (e == 0) return 1.0f;
    while ((e & 1) == 0) { base = xm(base, base); e >>= 1; }
    if (e == 1) return base;
    float acc = base;
    while (e > 1) {
        e >>= 1;
        base = xm(base, base);
        if (e & 1) acc = xm(acc, base);
    }
    return acc;
}

// ---------------------------------------------------------------- fast_math::exp2_raw (fast-math 0.1.1)
// decoder/spectral_noise_shaping.rs:122 - the approximation is part of the parity contract.
LC3B_HD float exp2_raw_fm(float x) {
    const float C2 = 1.00172476f;
    const float C1 = 0.657636276f * (1.0f / 8388608.0f);
    const float C0 = 0.3371894346f * (1.0f / 8388608.0f) * (1.0f / 8388608.0f);
    int32_t mul = cast_i32(xm(8388608.0f, x));
    int32_t fl = (int32_t)((uint32_t)mul & 0xFF800000u);
    float frac = (float)(mul - fl);
    float approx = xa(xm(xa(xm(C0, frac), C1), frac), C2);
    return u2f(f2u(approx) + (uint32_t)fl);
}

}  // namespace lc3b
