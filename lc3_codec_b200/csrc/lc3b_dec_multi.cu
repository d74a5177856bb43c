// lc3b engine: time-parallel decode - many consecutive frames of every stream in ONE call.
//
// The reference decodes a file frame by frame (examples/decode.rs:85-123); with few channels that leaves a GPU idle.
// Of the decoder's state only three things really run along time:
//   * overlap-add needs the previous frame's windowed tail            (modified_dct.rs:138-151)  -> one add
//   * concealment needs the last good spectrum, a fading factor and an LCG seed (packet_loss_concealment.rs:63-85)
//                                                                     -> a scan over the frames' ok flags
//   * the LTPF is an IIR across frame boundaries, but only while it is active (long_term_post_filter.rs:252)
// so a call with F frames of S streams runs as S*F independent "units" through the expensive stages and touches the
// time axis only where it must:
//   1. entropy_kernel + dequant_kernel over the S*F units (the unchanged kernels; spectrum to a per-unit slot)
//   2. plc_scan_kernel      thread per stream: walks the ok flags, assigns each concealed unit its source spectrum,
//                           alpha and seed (the LCG is jumped ne steps at a time), updates the handle's PLC scalars
//   3. imdct_multi_kernel   warp per unit: concealment scramble, DCT-IV (same FFT as kernel 2), window -> head / tail
//   4. ola_multi_kernel     warp per unit: x_hat = head + previous unit's tail, PCM of every frame
//   5. ltpf_multi_kernel    warp per stream: replays the post filter over the spans where it is (or just was) active,
//                           overwriting those frames' PCM; then leaves the handle's per-stream state (overlap memory,
//                           LTPF ring, last good spectrum) exactly as F frame-by-frame calls would have
// Arithmetic is the frame-by-frame path's, operation for operation, so both paths produce identical PCM.
#include "lc3b_common.cuh"
#include "lc3b_imdct.cuh"
#include "lc3b_plan.cuh"
#include "lc3b_math.cuh"

namespace lc3b {

struct MultiLayout { size_t spec, xq, handoff, gband, tns_list, side, head, tail, xhat, last_good, total; };

static MultiLayout multi_layout(const lc3b_config& c, int S, int F) {
    MultiLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = (off + bytes + 255) & ~(size_t)255; return o; };
    const size_t V = (size_t)S * F;
    const size_t nblk = ((V + 127) / 128) * 4;                    // 32-stream blocks, whole entropy CTAs
    L.spec = take(sizeof(float) * V * c.ne);
    L.xq = take(sizeof(int32_t) * nblk * c.ne * 32);
    L.handoff = take(sizeof(int32_t) * nblk * 32 * HO_WORDS);
    L.gband = take(sizeof(float) * nblk * 32 * 64);             // small unit counts take the warp-per-frame dequantisation
    L.tns_list = take(sizeof(int32_t) * (2 + nblk * 32));
    L.side = take(sizeof(int32_t) * V * SIDE_WORDS);
    L.head = take(sizeof(float) * V * c.nf);
    L.tail = take(sizeof(float) * V * (c.nf - c.z));
    L.xhat = take(sizeof(float) * V * c.nf);
    L.last_good = take(sizeof(int32_t) * S);
    L.total = off;
    return L;
}

size_t multi_scratch_bytes(const DecoderState& st, int n_frames) { return multi_layout(st.cfg, st.n_streams, n_frames).total; }

struct MultiParams {
    const DevConfig* cfg;
    const float* win;
    const float2* dtw;
    const float2* ftw;
    int S, F;
    // per-unit scratch
    const float* spec_v;
    int32_t* side_v;
    float* head;
    float* tail;
    float* xhat;
    int32_t* last_good;
    // the handle's per-stream state
    float* spec;
    float* ola;
    float* ltpf_y;
    float* ltpf_xtail;
    int32_t* sstate;
    int16_t* pcm_out;
    int hist_len;
};

constexpr int MW = 4;     // warps per CTA in the warp-per-unit kernels

// ---------------------------------------------------------------- 2. concealment bookkeeping along time
// One warp per stream; the ok flags are fetched 32 frames at a time so that the walk along time is not one dependent
// global load per frame.  Every lane tracks the same running state; lane j writes frame j's record.
__global__ void __launch_bounds__(MW * 32) plc_scan_kernel(MultiParams p) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = blockIdx.x * MW + wid;
    if (s >= p.S) return;
    const int ne = p.cfg->ne;
    int32_t* ss = p.sstate + (size_t)s * SS_WORDS;
    int lost = ss[SS_PLC_LOST];
    float alpha = u2f((uint32_t)ss[SS_PLC_ALPHA]);
    uint32_t seed = (uint32_t)ss[SS_PLC_SEED];
    __syncwarp();
    uint32_t A = 1, C = 0;                                        // ne steps of seed' = (16831 + seed * 12821) & 0xFFFF
    for (int i = 0; i < ne; i++) { C = (16831u + C * 12821u) & 0xFFFFu; A = (A * 12821u) & 0xFFFFu; }
    int last = -1;
    for (int f0 = 0; f0 < p.F; f0 += 32) {
        const int f = f0 + lane;
        const int v = s * p.F + f;
        int32_t* sd = p.side_v + (size_t)v * SIDE_WORDS;
        const bool in = f < p.F;
        const bool ok = in && sd[SD_OK] != 0;
        const uint32_t okm = __ballot_sync(0xffffffffu, ok), inm = __ballot_sync(0xffffffffu, in);
        if (okm == inm) {                                          // the usual chunk: every frame decoded
            if (in) sd[SD_SRC] = -2;
            lost = 0;
            alpha = 1.0f;
            last = s * p.F + f0 + (31 - __clz(inm));
            continue;
        }
        const int cnt = __popc(inm);
        for (int j = 0; j < cnt; j++) {
            if ((okm >> j) & 1u) {
                lost = 0;
                alpha = 1.0f;
                last = s * p.F + f0 + j;
                if (lane == j) sd[SD_SRC] = -2;
            } else {
                if (lost >= 4) alpha = xm(alpha, lost < 8 ? 0.9f : 0.85f);
                if (lane == j) {
                    sd[SD_SRC] = last;
                    sd[SD_PLC_ALPHA] = (int32_t)f2u(alpha);
                    sd[SD_PLC_SEED] = (int32_t)seed;
                }
                seed = (A * seed + C) & 0xFFFFu;
                lost++;
            }
        }
    }
    if (lane == 0) {
        ss[SS_PLC_LOST] = lost;
        ss[SS_PLC_ALPHA] = (int32_t)f2u(alpha);
        ss[SS_PLC_SEED] = (int32_t)seed;
        p.last_good[s] = last;
    }
}

// ---------------------------------------------------------------- 3. spectrum -> windowed time signal, per unit
template <int NF, bool MS10>
__global__ void __launch_bounds__(MW * 32) imdct_multi_kernel(MultiParams p) {
    using G = FrameGeo<NF, MS10>;
    constexpr int NE = G::NE, Z = G::Z, NP = G::NP, NPT = G::NPT;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long v = (long long)blockIdx.x * MW + wid;
    if (v >= (long long)p.S * p.F) return;
    const int s = (int)(v / p.F);
    float* P = (float*)smem + (size_t)wid * 2 * NF;
    float* Q = P + NF;

    const int32_t* sd = p.side_v + (size_t)v * SIDE_WORDS;
    const int src = sd[SD_SRC];
    const float* sp;
    if (src == -2) sp = p.spec_v + (size_t)v * NE;
    else if (src >= 0) sp = p.spec_v + (size_t)src * NE;
    else sp = p.spec + ((size_t)p.sstate[(size_t)s * SS_WORDS + SS_SLOT] * p.S + s) * NE;   // last good of earlier calls
    if (src == -2) {
#pragma unroll
        for (int k0 = 0; k0 < NF / 4; k0 += 32) {
            const int k4 = k0 + lane;
            if (k0 + 32 <= NF / 4 || k4 < NF / 4)
                ((float4*)P)[k4] = 4 * k4 < NE ? ((const float4*)sp)[k4] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    } else {                                                      // packet_loss_concealment.rs:63-85
        const float alpha = u2f((uint32_t)sd[SD_PLC_ALPHA]);
        uint32_t sgen = (uint32_t)sd[SD_PLC_SEED];
        for (int i = 0; i <= lane; i++) sgen = (16831u + sgen * 12821u) & 0xFFFFu;
        uint32_t a32 = 1, c32 = 0;
        for (int i = 0; i < 32; i++) { c32 = (16831u + c32 * 12821u) & 0xFFFFu; a32 = (a32 * 12821u) & 0xFFFFu; }
        for (int k = lane; k < NF; k += 32) {
            float val = 0.0f;
            if (k < NE) {
                const float lg = sp[k];
                val = sgen < 0x8000u ? xm(lg, alpha) : xm(lg, -alpha);
                sgen = (a32 * sgen + c32) & 0xFFFFu;
            }
            P[k] = val;
        }
    }
    __syncwarp();
    const float* D = dct_iv_warp<NF>(P, Q, p.dtw, p.ftw, lane);
    float hd[2 * NP], tl[2 * NPT];
    imdct_unfold<NF, MS10>(D, p.win, lane, hd, tl);
    float2* head = (float2*)(p.head + (size_t)v * NF);
    float2* tail = (float2*)(p.tail + (size_t)v * (NF - Z));
#pragma unroll
    for (int j = 0; j < NP; j++)
        if (64 * j + 64 <= NF || 64 * j + 2 * lane < NF) head[32 * j + lane] = make_float2(hd[2 * j], hd[2 * j + 1]);
#pragma unroll
    for (int j = 0; j < NPT; j++)
        if (64 * j + 64 <= NF - Z || 64 * j + 2 * lane < NF - Z) tail[32 * j + lane] = make_float2(tl[2 * j], tl[2 * j + 1]);
}

__device__ __forceinline__ int16_t round_pcm(float v) {            // output_scaling.rs:13-26
    int32_t q = v > 0.0f ? cast_i32(xa(v, 0.5f)) : cast_i32(xs(v, 0.5f));
    q = q > 32767 ? 32767 : q < -32768 ? -32768 : q;
    return (int16_t)q;
}

// ---------------------------------------------------------------- 4. overlap-add across units, PCM
__global__ void __launch_bounds__(MW * 32) ola_multi_kernel(MultiParams p) {
    const DevConfig& c = *p.cfg;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long v = (long long)blockIdx.x * MW + wid;
    if (v >= (long long)p.S * p.F) return;
    const int nf = c.nf, z = c.z;
    const int s = (int)(v / p.F), f = (int)(v - (long long)s * p.F);
    const float* head = p.head + (size_t)v * nf;
    const float* prev = f > 0 ? p.tail + (size_t)(v - 1) * (nf - z) : p.ola + (size_t)s * (nf - z);
    float* xhat = p.xhat + (size_t)v * nf;
    int16_t* out = p.pcm_out + (size_t)v * nf;                     // [S][F * nf]
    for (int n = lane; n < nf; n += 32) {
        float o = head[n];
        if (n < nf - z) o = xa(prev[n], o);
        xhat[n] = o;
        out[n] = round_pcm(o);
    }
}

// ---------------------------------------------------------------- 5. LTPF over its active spans + state hand-back
struct LtpfPar { int active, p_int, p_fr, code; };
__device__ __forceinline__ LtpfPar ltpf_params(const DevConfig& c, const int32_t* sd) {   // :142-190
    LtpfPar r{sd[SD_LTPF_ACTIVE], 0, 0, 4};
    if (r.active) {
        const int pi = sd[SD_PITCH_INDEX], nbits = sd[SD_NBITS];
        int pit;
        double pfr;
        if (pi >= 440) { pit = pi - 283; pfr = 0.0; }
        else if (pi >= 380) { pit = pi / 2 - 63; pfr = (double)(2 * pi - 4 * pit - 252); }
        else { pit = pi / 4 + 32; pfr = (double)(pi + 128 - 4 * pit); }
        const double pitch = (double)pit + pfr / 4.0;
        const double pitch_fs = pitch * (8000.0 * ceil((double)c.fs / 8000.0) / 12800.0);
        const int p_up = (int)((pitch_fs * 4.0) + 0.5);
        r.p_int = p_up / 4;
        r.p_fr = p_up - 4 * r.p_int;
        const int t_nbits = c.n_ms == LC3B_7P5MS ? (int)round((double)nbits * 10.0 / 7.5) : nbits;
        const int sf = c.fs_ind * 80;
        r.code = t_nbits < 320 + sf ? 0 : t_nbits < 400 + sf ? 1 : t_nbits < 480 + sf ? 2 : t_nbits < 560 + sf ? 3 : 4;
    }
    return r;
}

__global__ void __launch_bounds__(MW * 32) ltpf_multi_kernel(MultiParams p, int smem_per_warp) {
    extern __shared__ __align__(16) uint8_t smem[];
    const DevConfig& c = *p.cfg;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int s = blockIdx.x * MW + wid;
    if (s >= p.S) return;
    const int nf = c.nf, ne = c.ne, z = c.z, F = p.F;
    const int blocks = c.ltpf_blocks, l_num = c.ltpf_l_num, l_den = c.ltpf_l_den, norm = c.ltpf_norm, s2p5 = c.ltpf_s2p5;
    float* Y = (float*)(smem + (size_t)wid * smem_per_warp);       // hist_len
    float* X = Y + p.hist_len;                                     // nf
    float* xt = X + nf;                                            // 16: x_hat tail of the previous frame
    float* scratch = xt + 16;                                      // l_num + norm
    int32_t* ss = p.sstate + (size_t)s * SS_WORDS;
    float* yhist = p.ltpf_y + (size_t)s * p.hist_len;
    float* xtail = p.ltpf_xtail + (size_t)s * XTAIL_FLOATS;      // [blocks][16], slot = ring block of the frame it closes
    const int blk0 = ss[SS_LTPF_BLK];
    const int prevw = ss[SS_LTPF_PREV];
    int prev_active = prevw & 1, prev_code = prevw >> 8, p_int_mem = ss[SS_LTPF_PINT], p_fr_mem = ss[SS_LTPF_PFR];
    const int slot = ss[SS_SLOT];
    __syncwarp();
    for (int n = lane; n < p.hist_len; n += 32) Y[n] = yhist[n];
    __syncwarp();
    bool span1 = false, span2 = false, span3 = false;              // were frames f-1 / f-2 / f-3 filtered by this kernel
    const size_t v0 = (size_t)s * F;
    uint32_t actm = 0;                                             // LTPF-active flags of the current 32-frame chunk
    for (int f = 0; f < F; f++) {
        if ((f & 31) == 0) {                                       // fetch 32 frames' flags at once; skip idle chunks
            const int fl = f + lane;
            actm = __ballot_sync(0xffffffffu, fl < F && p.side_v[(v0 + fl) * SIDE_WORDS + SD_LTPF_ACTIVE] != 0);
            if (actm == 0 && !prev_active) {
                span1 = span2 = span3 = false;
                prev_code = 4; p_int_mem = 0; p_fr_mem = 0;
                f += 31;
                continue;
            }
        }
        const size_t v = v0 + f;
        const LtpfPar cur = ltpf_params(c, p.side_v + v * SIDE_WORDS);
        const int active = cur.active, p_int = cur.p_int, p_fr = cur.p_fr, code = cur.code;
        const bool in_span = active || prev_active;
        if (in_span) {
            const int blk = ((blk0 + f) % blocks) * nf;
            // history this frame may read: frames f-1 .. f-blocks (the oldest one sits in the block this frame is about
            // to overwrite: the ring is circular and long pitch lags reach into it); refill the ones this kernel skipped
            for (int j = 1; j <= blocks; j++) {
                const int g = f - j;
                const bool have = j == 1 ? span1 : j == 2 ? span2 : span3;
                if (g >= 0 && !have) {
                    const float* src = p.xhat + (v0 + g) * nf;
                    float* dst = Y + ((blk0 + g) % blocks) * nf;
                    for (int n = lane; n < nf; n += 32) dst[n] = src[n];
                }
            }
            const float* xh = p.xhat + v * nf;
            for (int n = lane; n < nf; n += 32) X[n] = xh[n];
            if (lane < 16) xt[lane] = f > 0 ? p.xhat[(v - 1) * nf + nf - 16 + lane] : xtail[((blk0 + blocks - 1) % blocks) * 16 + lane];
            __syncwarp();
            // ---- long_term_post_filter.rs:252-343, as in synth_kernel
            auto cnum = [&](int cd, int k) -> float { return cd < 4 ? c.ltpf_num[cd][k] : 0.0f; };
            auto cden = [&](int cd, int fr, int k) -> float { return cd < 4 ? c.ltpf_den[cd][fr][k] : 0.0f; };
            const int cur_code = active ? code : 4;
            const int mem_code = prev_active ? prev_code : 4;
            auto wrapi = [&](int idx) -> int { return idx < 0 ? idx + p.hist_len : idx; };
            auto x_at = [&](int pos) -> float {
                const int rel = pos - blk;
                return rel >= 0 ? X[rel] : xt[16 + rel];
            };
            auto filt = [&](int start, int pint, int cd, int fr) -> float {
                float out = 0.0f;
                for (int k = 0; k <= l_num; k++) out = xa(out, xm(cnum(cd, k), x_at(start - k)));
                const int sden = start - pint + l_den / 2;
                for (int k = 0; k <= l_den; k++) out = xs(out, xm(cden(cd, fr, k), Y[wrapi(sden - k)]));
                return out;
            };
            const float fnorm = (float)norm;
            auto chunked = [&](int n0, int n1, int reach, auto body) {
                const int C = reach < 32 ? reach : 32;
                for (int st = n0; st < n1; st += C) {
                    const int n = st + lane;
                    if (lane < C && n < n1) body(n);
                    __syncwarp();
                }
            };
            const int reach_cur = active ? p_int - l_den / 2 : 32;
            const int reach_mem = prev_active ? p_int_mem - l_den / 2 : 32;
            auto deactivate_first = [&]() {
                chunked(0, s2p5, reach_mem, [&](int n) {
                    float fo = filt(blk + n, p_int_mem, mem_code, p_fr_mem);
                    fo = xm(fo, xs(1.0f, xd((float)n, fnorm)));
                    Y[blk + n] = xs(X[n], fo);
                });
            };
            auto plain_from = [&](int n0) {
                chunked(n0, nf, reach_cur, [&](int n) { Y[blk + n] = xs(X[n], filt(blk + n, p_int, cur_code, p_fr)); });
            };
            if (active && !prev_active) {                // case 2
                chunked(0, s2p5, reach_cur, [&](int n) {
                    float fo = filt(blk + n, p_int, cur_code, p_fr);
                    fo = xm(fo, xd((float)n, fnorm));
                    Y[blk + n] = xs(X[n], fo);
                });
                plain_from(s2p5);
            } else if (!active && prev_active) {         // case 3
                deactivate_first();
                for (int n = s2p5 + lane; n < nf; n += 32) Y[blk + n] = X[n];
            } else if (p_int == p_int_mem && p_fr == p_fr_mem) {   // case 4
                plain_from(0);
            } else {                                     // case 5
                deactivate_first();
                for (int i = lane; i < l_num + norm; i += 32) {
                    int src;
                    if (blk < l_num) src = i < l_num ? blocks * nf - l_num + i : i - l_num;
                    else src = blk - l_num + i;
                    scratch[i] = Y[src];
                }
                __syncwarp();
                chunked(0, s2p5, reach_cur, [&](int n) {
                    float fo = 0.0f;
                    for (int k = 0; k <= l_num; k++) fo = xa(fo, xm(cnum(cur_code, k), scratch[l_num + n - k]));
                    const int sden = blk + n - p_int + l_den / 2;
                    for (int k = 0; k <= l_den; k++) fo = xs(fo, xm(cden(cur_code, p_fr, k), Y[wrapi(sden - k)]));
                    fo = xm(fo, xd((float)n, fnorm));
                    Y[blk + n] = xs(scratch[n + l_num], fo);
                });
                plain_from(s2p5);
            }
            __syncwarp();
            int16_t* out = p.pcm_out + v * nf;
            for (int n = lane; n < nf; n += 32) out[n] = round_pcm(Y[blk + n]);
            __syncwarp();
        }
        span3 = span2;
        span2 = span1;
        span1 = in_span;
        prev_active = active;
        prev_code = active ? code : 4;
        p_int_mem = p_int;
        p_fr_mem = p_fr;
    }
    if (F == 0) return;
    // ---- hand the per-stream state back as F frame-by-frame calls would have left it
    for (int j = 0; j < blocks; j++) {                             // LTPF ring: the last `blocks` frames' output
        const int g = F - 1 - j;
        if (g < 0) break;
        const int bo = ((blk0 + g) % blocks) * nf;
        const bool have = j == 0 ? span1 : j == 1 ? span2 : span3;
        const float* src = have ? Y + bo : p.xhat + (v0 + g) * nf;
        for (int n = lane; n < nf; n += 32) yhist[bo + n] = src[n];
    }
    if (lane < 16) xtail[((blk0 + F - 1) % blocks) * 16 + lane] = p.xhat[(v0 + F - 1) * nf + nf - 16 + lane];
    {
        const float* t = p.tail + (v0 + F - 1) * (nf - z);
        float* ola = p.ola + (size_t)s * (nf - z);
        for (int n = lane; n < nf - z; n += 32) ola[n] = t[n];
    }
    const int lg = p.last_good[s];
    if (lg >= 0) {                                                 // PacketLossConcealment::save of the last good frame
        const float* src = p.spec_v + (size_t)lg * ne;
        float* dst = p.spec + ((size_t)(slot ^ 1) * p.S + s) * ne;
        for (int k = lane; k < ne; k += 32) dst[k] = src[k];
    }
    if (lane == 0) {
        ss[SS_LTPF_PREV] = (prev_active ? 1 : 0) | (prev_code << 8);
        ss[SS_LTPF_PINT] = p_int_mem;
        ss[SS_LTPF_PFR] = p_fr_mem;
        ss[SS_LTPF_BLK] = (blk0 + F) % blocks;
        if (lg >= 0) ss[SS_SLOT] = slot ^ 1;
    }
}

static size_t ltpf_multi_warp_bytes(const lc3b_config& c) {
    const int blocks = c.n_ms == LC3B_10MS ? 2 : 3;
    size_t b = (size_t)(blocks * c.nf + c.nf + 16 + 16 + c.nf / 3 + 16) * 4;
    return (b + 15) & ~(size_t)15;
}

namespace {
struct PrepareImdct {
    cudaError_t e = cudaSuccess;
    template <int NF, bool MS10> void operator()() {
        e = cudaFuncSetAttribute(imdct_multi_kernel<NF, MS10>, cudaFuncAttributeMaxDynamicSharedMemorySize, MW * 2 * NF * 4);
    }
};
struct LaunchImdct {
    const MultiParams& p;
    unsigned grid;
    cudaStream_t stream;
    template <int NF, bool MS10> void operator()() {
        imdct_multi_kernel<NF, MS10><<<grid, MW * 32, MW * 2 * NF * 4, stream>>>(p);
    }
};
}  // namespace

cudaError_t prepare_multi(const DecoderState& st) {
    cudaError_t e = cudaFuncSetAttribute(ltpf_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(ltpf_multi_warp_bytes(st.cfg) * MW));
    if (e == cudaSuccess) {
        PrepareImdct pi;
        if (!dispatch_frame_geo(st.cfg.nf, st.cfg.n_ms == LC3B_10MS, pi)) return cudaErrorInvalidValue;
        e = pi.e;
    }
    return e;
}

cudaError_t launch_decode_multi(const DecoderState& st, const uint8_t* frames, const int32_t* frame_nbytes, int nbytes,
                                size_t frame_stride, int n_frames, int16_t* pcm_out, int32_t* status_out, void* scratch,
                                cudaStream_t stream, PlanLanes* lanes) {
    const int S = st.n_streams, F = n_frames;
    if (F <= 0) return cudaSuccess;
    const MultiLayout L = multi_layout(st.cfg, S, F);
    uint8_t* base = (uint8_t*)scratch;
    const long long V = (long long)S * F;
    // 1. entropy decode of all units with the unchanged kernel: a view of the handle with V virtual streams
    DecoderState vs = st;
    vs.n_streams = (int)V;
    vs.spec = (float*)(base + L.spec);
    vs.xq = (int32_t*)(base + L.xq);
    vs.handoff = (int32_t*)(base + L.handoff);
    vs.gband = (float*)(base + L.gband);
    vs.tns_list = (int32_t*)(base + L.tns_list);
    vs.nsym_prev = nullptr;                                       // units of one call are not consecutive frames of a stream slot
    vs.side = (int32_t*)(base + L.side);
    vs.fixed_slot = 0;
    vs.trace = nullptr;
    vs.trace_x = nullptr;
    cudaError_t e = cudaMemsetAsync(vs.tns_list, 0, 2 * sizeof(int32_t), stream);   // caller-owned scratch: the list starts empty
    if (e == cudaSuccess) {
        // large calls: the units as four independent sub-batches whose kernels overlap (run_decode in lc3b_api.cu)
        LaunchPlan plan;
        const int n = vs.n_streams;
        const int k_sub = (lanes && n >= decode_split_min_streams() && !use_dequant_warp(n, vs.dequant_mode)) ? PLAN_MAX_LANES : 1;
        int part = ((n + k_sub - 1) / k_sub + 127) & ~127;
        if (k_sub == 1) part = n;
        for (int k = 0, b = 0; k < k_sub && b < n; k++, b += part) {
            plan.lane = k;
            plan_entropy(plan, vs, frames, frame_nbytes, nbytes, frame_stride, status_out, 3, b, k_sub == 1 ? -1 : (n - b < part ? n - b : part), -1);
        }
        e = plan_launch_direct(plan, stream, lanes);
    }
    if (e != cudaSuccess) return e;
    MultiParams p;
    p.cfg = st.dcfg; p.win = st.win; p.dtw = st.dtw; p.ftw = st.ftw;
    p.S = S; p.F = F;
    p.spec_v = vs.spec; p.side_v = vs.side;
    p.head = (float*)(base + L.head); p.tail = (float*)(base + L.tail); p.xhat = (float*)(base + L.xhat);
    p.last_good = (int32_t*)(base + L.last_good);
    p.spec = st.spec; p.ola = st.ola; p.ltpf_y = st.ltpf_y; p.ltpf_xtail = st.ltpf_xtail; p.sstate = st.sstate;
    p.pcm_out = pcm_out;
    p.hist_len = (st.cfg.n_ms == LC3B_10MS ? 2 : 3) * st.cfg.nf;
    plc_scan_kernel<<<(S + MW - 1) / MW, MW * 32, 0, stream>>>(p);
    const unsigned grid_v = (unsigned)((V + MW - 1) / MW);
    LaunchImdct li{p, grid_v, stream};
    if (!dispatch_frame_geo(st.cfg.nf, st.cfg.n_ms == LC3B_10MS, li)) return cudaErrorInvalidValue;
    ola_multi_kernel<<<grid_v, MW * 32, 0, stream>>>(p);
    const size_t lw = ltpf_multi_warp_bytes(st.cfg);
    ltpf_multi_kernel<<<(S + MW - 1) / MW, MW * 32, lw * MW, stream>>>(p, (int)lw);
    return cudaGetLastError();
}

}  // namespace lc3b
