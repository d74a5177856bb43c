// lc3b engine: shared-memory Stockham FFT pieces used by the synthesis kernels (decoder).
// Radix-2/3/4/5 butterflies and one autosort stage spread over a warp; N = nf/2 in {30, ..., 240}.
#pragma once
#include <cuda_runtime.h>

namespace lc3b {

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

template <int R>
__device__ __forceinline__ void dft(float2 (&v)[R]);
template <>
__device__ __forceinline__ void dft<2>(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = caddf(a, b);
    v[1] = csubf(a, b);
}
template <>
__device__ __forceinline__ void dft<3>(float2 (&v)[3]) {
    const float S = -0.86602540378443864676f;          // -sin(2 pi / 3)
    float2 t = caddf(v[1], v[2]);
    float2 d = csubf(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * t.x, v[0].y - 0.5f * t.y);
    float2 r = make_float2(-S * d.y, S * d.x);         // -i * sin * d  (forward transform)
    v[0] = caddf(v[0], t);
    v[1] = caddf(m, r);
    v[2] = csubf(m, r);
}
template <>
__device__ __forceinline__ void dft<4>(float2 (&v)[4]) {
    float2 a = caddf(v[0], v[2]), b = csubf(v[0], v[2]);
    float2 c = caddf(v[1], v[3]), d = csubf(v[1], v[3]);
    float2 dj = make_float2(d.y, -d.x);                // -i * d
    v[0] = caddf(a, c);
    v[1] = caddf(b, dj);
    v[2] = csubf(a, c);
    v[3] = csubf(b, dj);
}
template <>
__device__ __forceinline__ void dft<5>(float2 (&v)[5]) {
    const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;   // cos(2pi/5), cos(4pi/5)
    const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;    // sin(2pi/5), sin(4pi/5)
    float2 a1 = caddf(v[1], v[4]), b1 = csubf(v[1], v[4]);
    float2 a2 = caddf(v[2], v[3]), b2 = csubf(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x, v[0].y + C1 * a1.y + C2 * a2.y);
    float2 m2 = make_float2(v[0].x + C2 * a1.x + C1 * a2.x, v[0].y + C2 * a1.y + C1 * a2.y);
    // forward transform: X[k] = m - i*(S..)*b ; -i*(x+iy) = (y, -x)
    float2 n1 = make_float2(S1 * b1.y + S2 * b2.y, -(S1 * b1.x + S2 * b2.x));
    float2 n2 = make_float2(S2 * b1.y - S1 * b2.y, -(S2 * b1.x - S1 * b2.x));
    v[0] = make_float2(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
    v[1] = caddf(m1, n1);
    v[4] = csubf(m1, n1);
    v[2] = caddf(m2, n2);
    v[3] = csubf(m2, n2);
}

// one Stockham stage: N/R butterflies spread over the warp
template <int R>
__device__ __forceinline__ void fft_stage(const float2* __restrict__ x, float2* __restrict__ y, const float2* __restrict__ ftw,
                                          int N, int Ns, int lane) {
    const int nbf = N / R;
    const int tw_step = N / (Ns * R);
    for (int j = lane; j < nbf; j += 32) {
        const int k = j % Ns;
        float2 v[R];
#pragma unroll
        for (int t = 0; t < R; t++) {
            v[t] = x[j + t * nbf];
            if (t > 0) v[t] = cmulf(v[t], ftw[k * t * tw_step]);
        }
        dft<R>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int t = 0; t < R; t++) y[base + t * Ns] = v[t];
    }
}

// DCT-IV of X[0..nf) in place (dct_iv.rs:49-67): pre-twiddle, N-point FFT over the ping-pong buffers, post-twiddle.
__device__ __forceinline__ void dct_iv_warp(float* X, float2* bufA, float2* bufB, const float2* __restrict__ dtw,
                                            const float2* __restrict__ ftw, const int32_t* fft_radix, int nf, int N, int lane) {
    for (int n = lane; n < N; n += 32) bufA[n] = cmulf(dtw[n], make_float2(X[2 * n], X[nf - 1 - 2 * n]));
    __syncwarp();
    float2 *src = bufA, *dst = bufB;
    int Ns = 1;
    for (int sidx = 0; sidx < 8 && fft_radix[sidx] != 0; sidx++) {
        const int R = fft_radix[sidx];
        switch (R) {
            case 2: fft_stage<2>(src, dst, ftw, N, Ns, lane); break;
            case 3: fft_stage<3>(src, dst, ftw, N, Ns, lane); break;
            case 4: fft_stage<4>(src, dst, ftw, N, Ns, lane); break;
            default: fft_stage<5>(src, dst, ftw, N, Ns, lane); break;
        }
        Ns *= R;
        float2* t = src; src = dst; dst = t;
        __syncwarp();
    }
    for (int n = lane; n < N; n += 32) {
        const float2 v = cmulf(dtw[n], src[n]);
        X[2 * n] = v.x * 2.0f;
        X[nf - 1 - 2 * n] = -v.y * 2.0f;
    }
    __syncwarp();
}

}  // namespace lc3b
