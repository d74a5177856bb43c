// lc3b engine, decoder kernels 2 and 3: shaped spectrum -> PCM, one WARP per frame.
//
// Replaces, per stream, the second half of DecoderChannel::decode (src/decoder/lc3_decoder.rs:134-153):
//   PacketLossConcealment::save / load_into   src/decoder/packet_loss_concealment.rs:50,63
//   ModDiscreteCosTrans::run                  src/decoder/modified_dct.rs:76 (DCT-IV via N = nf/2 complex FFT,
//                                             src/common/dct_iv.rs:49, src/common/kissfft.rs:78; window; overlap-add)
//   LongTermPostFilter::run                   src/decoder/long_term_post_filter.rs:252 (five transition cases)
//   output_scaling::scale_and_round           src/decoder/output_scaling.rs:13
//
// synth_kernel<NF, MS10> (every stream): concealment, inverse MDCT (lc3b_imdct.cuh, frame length a template
// constant), overlap-add, and the post filter's PASS-THROUGH case: x_hat goes to the LTPF history ring and, rounded,
// to the PCM row.  Two consecutive samples per lane, so PCM leaves as 32-bit and f32 state as 64-bit accesses; all
// global traffic is stream-major and coalesced (spectrum 4*ne B in, overlap memory 4*(nf-z) B in/out, LTPF history
// 4*nf B out, PCM 2*nf B out).
// ltpf_kernel (streams whose post filter is, or was in the previous frame, active): the IIR itself; for those streams
// synth_kernel leaves x_hat in a side buffer (the ring block keeps the samples long pitch lags still reach), and this
// kernel writes the filtered signal to the ring block and the PCM row.  The recursion only reaches back
// p_int - l_den/2 >= 18 samples, so up to 32 consecutive outputs are computed in parallel per step.  Streams outside
// such a span cost this kernel one flag load.
#include <stdlib.h>
#include <string.h>

#include "lc3b_common.cuh"
#include "lc3b_imdct.cuh"
#include "lc3b_math.cuh"
#include "lc3b_plan.cuh"

namespace lc3b {

struct SynthParams {
    const DevConfig* cfg;
    const float* win;
    const float2* dtw;
    const float2* ftw;
    const float* spec;
    float* ola;
    float* ltpf_y;
    float* ltpf_xtail;
    float* ltpf_x;       // [n_streams][nf] x_hat of streams inside an active span (synth_kernel -> ltpf_kernel)
    int32_t* side;
    int32_t* sstate;
    int16_t* pcm_out;
    size_t pcm_stride;
    int n_streams;       // streams of this launch (a sub-batch when a call is split)
    int slot_streams;    // streams per spectrum slot = the handle's stream count
    int hist_len;        // ltpf_blocks * nf
    int pcm_pairs;       // PCM rows are 4-byte aligned: samples leave two at a time
    int group;           // ltpf_kernel: streams looked after by one warp
    int smem_per_warp;   // ltpf_kernel, bytes
    int no_ltpf;         // the handle's minimum frame length rules the post filter out: no history is kept (lc3b_decoder_set_min_nbytes)
};

constexpr int SYN_WARPS = 4;
constexpr int PIPE_WARPS = 8;    // warps per CTA of the persistent synthesis kernel (they share the CTA's copy of the tables)
// per warp: two stage buffers + one pong buffer of nf floats and two mbarriers; per CTA: DCT-IV and FFT twiddles (nf/2 float2
// each) and the window (2 nf floats reserved)
static inline size_t synth_smem_bytes(int nf) { return (size_t)PIPE_WARPS * 3 * nf * 4 + (size_t)4 * nf * 4 + (size_t)PIPE_WARPS * 2 * 8; }

__device__ __forceinline__ int32_t round_pcm_i32(float v) {          // output_scaling.rs:13-26
    int32_t q = v > 0.0f ? cast_i32(xa(v, 0.5f)) : cast_i32(xs(v, 0.5f));
    return q > 32767 ? 32767 : q < -32768 ? -32768 : q;
}

// synth_classic_kernel: one warp = one frame, spectrum through ordinary loads.  The default: at 262 144 streams it runs in
// 0.50 ms against 0.525 ms for the persistent TMA-pipelined kernel below (profiles/r2_synth_ab.json).  What bounds it is
// the SM's L1 / shared-memory data path, not instruction issue, DRAM or latency: dropping a quarter of its instructions
// (the float -> int casts) changed nothing, the time per frame is the same at 16 384 streams (everything in L2) as at
// 262 144, and MORE resident warps make it slower (8 / 10 / 12 CTAs per SM: 0.502 / 0.516 / 0.529 ms - the twiddle and
// window tables share L1 with the shared-memory carve-out).  The per-frame streams (spectrum, overlap memory, PCM)
// therefore use evict-first loads and stores, which leaves L1 to the tables (0.517 -> 0.500 ms).
template <int NF, bool MS10>
__global__ void __launch_bounds__(SYN_WARPS * 32) synth_classic_kernel(const __grid_constant__ SynthParams p) {
    using G = FrameGeo<NF, MS10>;
    constexpr int NE = G::NE, Z = G::Z, NP = G::NP, NPT = G::NPT;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int stream = blockIdx.x * SYN_WARPS + wid;
    if (stream >= p.n_streams) return;                            // warps are independent: no CTA-wide barrier below

    float* P = (float*)smem + (size_t)wid * 2 * NF;               // spectrum in; FFT ping
    float* Q = P + NF;                                            // FFT pong

    int32_t* sd = p.side + (size_t)stream * SIDE_WORDS;
    int32_t* ss = p.sstate + (size_t)stream * SS_WORDS;
    // the two records as 16-byte loads (every lane reads the same words; each load is one slot of the memory pipe)
    static_assert(SD_OK == 0 && SD_LTPF_ACTIVE == 1 && SD_SLOT == 6 && SS_PLC_LOST == 1 && SS_PLC_ALPHA == 2 && SS_LTPF_PREV == 4 && SS_LTPF_BLK == 7,
                  "record layout");
    const int4 sd_a = ((const int4*)sd)[0], sd_b = ((const int4*)sd)[1];
    const int4 ss_a = ((const int4*)ss)[0], ss_b = ((const int4*)ss)[1];
    const int ok = sd_a.x;
    const int slot = sd_b.z;
    const int active = sd_a.y;
    const int prev_active = ss_b.x & 1;
    const int blk_idx = ss_b.w;
    const float* sp = p.spec + ((size_t)slot * p.slot_streams + stream) * NE;

    // overlap memory: requested now, consumed after the FFT
    float* ola = p.ola + (size_t)stream * (NF - Z);
    float2 ola_r[NPT];
#pragma unroll
    for (int j = 0; j < NPT; j++)
        ola_r[j] = (64 * j + 64 <= NF - Z || 64 * j + 2 * lane < NF - Z) ? __ldcs((const float2*)ola + 32 * j + lane) : make_float2(0.0f, 0.0f);

    // ---- spectrum load, with concealment (packet_loss_concealment.rs:63-85) when the frame was bad
    if (ok) {
#pragma unroll
        for (int k0 = 0; k0 < NF / 4; k0 += 32) {
            const int k4 = k0 + lane;
            if (k0 + 32 <= NF / 4 || k4 < NF / 4)
                ((float4*)P)[k4] = 4 * k4 < NE ? __ldcs((const float4*)sp + k4) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
        if (lane == 0 && (ss_a.y != 0 || ss_a.z != (int32_t)f2u(1.0f))) { ss[SS_PLC_LOST] = 0; ss[SS_PLC_ALPHA] = (int32_t)f2u(1.0f); }
    } else {
        int lost = ss_a.y;
        float alpha = u2f((uint32_t)ss_a.z);
        uint32_t seed = (uint32_t)ss_a.w;
        __syncwarp();                                            // every lane has read the state before any lane updates it
        if (lost >= 4) alpha = xm(alpha, lost < 8 ? 0.9f : 0.85f);
        // lane's first line is k = lane: LCG advanced lane+1 times; then jump 32 steps at a time
        uint32_t s = seed;
        for (int i = 0; i <= lane; i++) s = (16831u + s * 12821u) & 0xFFFFu;
        uint32_t a32 = 1, c32 = 0;                               // composition of 32 LCG steps
        for (int i = 0; i < 32; i++) { c32 = (16831u + c32 * 12821u) & 0xFFFFu; a32 = (a32 * 12821u) & 0xFFFFu; }
        for (int k = lane; k < NF; k += 32) {
            float v = 0.0f;
            if (k < NE) {
                const float lg = sp[k];
                v = s < 0x8000u ? xm(lg, alpha) : xm(lg, -alpha);
                if (k == NE - 1) ss[SS_PLC_SEED] = (int32_t)s;
                s = (a32 * s + c32) & 0xFFFFu;
            }
            P[k] = v;
        }
        if (lane == 0) { ss[SS_PLC_LOST] = lost + 1; ss[SS_PLC_ALPHA] = (int32_t)f2u(alpha); }
    }
    __syncwarp();

    // ---- DCT-IV, unfold, window (dct_iv.rs:49-67, modified_dct.rs:97-136)
    const float* D = dct_iv_warp<NF>(P, Q, p.dtw, p.ftw, lane);
    float head[2 * NP], tail[2 * NPT];
    imdct_unfold<NF, MS10>(D, p.win, lane, head, tail);

    // ---- overlap-add (modified_dct.rs:138-151): two roundings, like the reference (and the time-parallel path)
#pragma unroll
    for (int j = 0; j < NPT; j++) {
        if (64 * j + 64 <= NF - Z || 64 * j + 2 * lane < NF - Z) {
            head[2 * j] = xa(ola_r[j].x, head[2 * j]);
            head[2 * j + 1] = xa(ola_r[j].y, head[2 * j + 1]);
            __stcs((float2*)ola + 32 * j + lane, make_float2(tail[2 * j], tail[2 * j + 1]));
        }
    }

    // ---- post filter, pass-through case (long_term_post_filter.rs:252-343 with zero coefficients): history + PCM
    // Inside an active span the ring block must keep its old content until the filter has run (long pitch lags reach
    // around the ring into it), so x_hat waits in a side buffer for ltpf_kernel instead.
    const bool in_span = active || prev_active;
    float* yblk = in_span ? p.ltpf_x + (size_t)stream * NF : p.ltpf_y + (size_t)stream * p.hist_len + blk_idx * NF;
    float* xtail = p.ltpf_xtail + (size_t)stream * XTAIL_FLOATS + blk_idx * 16;
    int16_t* out = p.pcm_out + (size_t)stream * p.pcm_stride;
#pragma unroll
    for (int j = 0; j < NP; j++) {
        const int n = 64 * j + 2 * lane;
        if (64 * j + 64 <= NF || n < NF) {
            const float a = head[2 * j], b = head[2 * j + 1];
            if (!p.no_ltpf) {
                ((float2*)yblk)[32 * j + lane] = make_float2(a, b);
                if (n >= NF - 16) ((float2*)xtail)[(n - (NF - 16)) >> 1] = make_float2(a, b);   // x_hat tail for the next frame's filter
            }
            const int32_t qa = round_pcm_i32(a), qb = round_pcm_i32(b);
            if (p.pcm_pairs) __stcs((uint32_t*)out + 32 * j + lane, ((uint32_t)qa & 0xffffu) | ((uint32_t)qb << 16));
            else { out[n] = (int16_t)qa; out[n + 1] = (int16_t)qb; }
        }
    }
    if (lane == 0 && !p.no_ltpf) {
        sd[SD_BLK] = blk_idx;
        int nb = blk_idx + 1;                                    // :334-337
        if (nb > (MS10 ? 1 : 2)) nb = 0;
        ss[SS_LTPF_BLK] = nb;
        if (!in_span) {                                          // otherwise ltpf_kernel carries the filter state on
            ss[SS_LTPF_PREV] = 4 << 8;
            ss[SS_LTPF_PINT] = 0;
            ss[SS_LTPF_PFR] = 0;
        }
    }
}


// ---- 1-D bulk copies global -> shared through the TMA unit, completion on an mbarrier (PTX cp.async.bulk, sm_90+)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LC3B_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LC3B_DONE;\n"
        "bra LC3B_WAIT;\n"
        "LC3B_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// generic-proxy accesses of a buffer are ordered before the async-proxy (TMA) writes issued afterwards
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// What the kernel needs to know about a frame before it can fetch its spectrum: a handful of words per stream.
struct FrameHead {
    int ok, slot, active, prev_active, blk_idx;
};
__device__ __forceinline__ FrameHead load_head(const SynthParams& p, int stream) {
    const int32_t* sd = p.side + (size_t)stream * SIDE_WORDS;
    const int32_t* ss = p.sstate + (size_t)stream * SS_WORDS;
    FrameHead h;
    h.ok = sd[SD_OK];
    h.slot = sd[SD_SLOT];
    h.active = sd[SD_LTPF_ACTIVE];
    h.prev_active = ss[SS_LTPF_PREV] & 1;
    h.blk_idx = ss[SS_LTPF_BLK];
    return h;
}

// synth_kernel: PERSISTENT, one warp per frame at a time.  A warp walks the frames w, w + W, w + 2W, ... (W = warps in
// the grid); while it transforms frame i the TMA unit already fetches the spectrum of frame i + 1 into the warp's other
// stage buffer (cp.async.bulk + mbarrier: no registers, no issue slots), and the side words of frame i + 2 are on
// their way, so a warp never waits for HBM latency once it is running.  Per warp: two stage buffers of nf floats (the
// spectrum arrives in the first ne, the rest stays zero; the stage buffer then serves as the FFT's ping buffer) and
// one pong buffer.
template <int NF, bool MS10>
__global__ void __launch_bounds__(PIPE_WARPS * 32, 4) synth_kernel(const __grid_constant__ SynthParams p) {
    using G = FrameGeo<NF, MS10>;
    constexpr int NE = G::NE, Z = G::Z, NP = G::NP, NPT = G::NPT, N = NF / 2;
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n_warps = gridDim.x * PIPE_WARPS;
    int stream = blockIdx.x * PIPE_WARPS + wid;

    // the CTA's copy of the transform tables: a persistent CTA pays for it once and every later access is a shared-memory load
    float2* s_dtw = (float2*)((float*)smem + (size_t)PIPE_WARPS * 3 * NF);
    float2* s_ftw = s_dtw + N;
    float* s_win = (float*)(s_ftw + N);                           // win[Z .. 2 NF): the part the unfold reads, indexed from 0
    for (int i = threadIdx.x; i < N; i += PIPE_WARPS * 32) { s_dtw[i] = p.dtw[i]; s_ftw[i] = p.ftw[i]; }
    for (int i = threadIdx.x; i < 2 * NF - Z; i += PIPE_WARPS * 32) s_win[i] = p.win[Z + i];
    __syncthreads();
    if (stream >= p.n_streams) return;                            // warps are independent from here on: no CTA-wide barrier below

    float* S0 = (float*)smem + (size_t)wid * 3 * NF;              // stage buffers (spectrum in; FFT ping)
    float* Q = S0 + 2 * NF;                                       // FFT pong
    uint64_t* bars = (uint64_t*)(s_win + 2 * NF) + 2 * wid;
    if (lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); }
    for (int k = lane; k < 2 * NF; k += 32) S0[k] = 0.0f;         // the zero padding beyond ne (both stages)
    __syncwarp();
    fence_proxy_async();

    auto fetch = [&](int st, int s, const FrameHead& h) {         // spectrum of stream s -> stage st
        if (lane == 0) {
            const float* sp = p.spec + ((size_t)h.slot * p.slot_streams + s) * NE;
            mbar_expect_tx(&bars[st], NE * 4);
            bulk_g2s(S0 + st * NF, sp, NE * 4, &bars[st]);
        }
    };
    FrameHead head = load_head(p, stream);
    fetch(0, stream, head);
    int nstream = stream + n_warps;
    FrameHead nhead = head;
    if (nstream < p.n_streams) nhead = load_head(p, nstream);
    uint32_t phase0 = 0, phase1 = 0;

    for (int it = 0; stream < p.n_streams; it++) {
        const int st = it & 1;
        float* P = S0 + st * NF;
        // ---- next frame: its spectrum starts moving now, the side words of the one after are requested
        const int n2stream = nstream + n_warps;
        FrameHead n2head = nhead;
        if (nstream < p.n_streams) {
            // the other stage buffer was this warp's FFT ping buffer one frame ago: restore its zero tail, then hand it to the TMA
            float* O = S0 + (st ^ 1) * NF;
            if (it > 0) {
                for (int k = NE + lane; k < NF; k += 32) O[k] = 0.0f;
                __syncwarp();
                fence_proxy_async();
            }
            fetch(st ^ 1, nstream, nhead);
            if (n2stream < p.n_streams) n2head = load_head(p, n2stream);
        }

        int32_t* sd = p.side + (size_t)stream * SIDE_WORDS;
        int32_t* ss = p.sstate + (size_t)stream * SS_WORDS;
        const int ok = head.ok, active = head.active, prev_active = head.prev_active, blk_idx = head.blk_idx;

        // overlap memory: requested now, consumed after the FFT
        float* ola = p.ola + (size_t)stream * (NF - Z);
        float2 ola_r[NPT];
#pragma unroll
        for (int j = 0; j < NPT; j++)
            ola_r[j] = (64 * j + 64 <= NF - Z || 64 * j + 2 * lane < NF - Z) ? ((const float2*)ola)[32 * j + lane] : make_float2(0.0f, 0.0f);

        // ---- the spectrum has landed in P; concealment (packet_loss_concealment.rs:63-85) when the frame was bad
        mbar_wait(&bars[st], st ? phase1 : phase0);
        if (st) phase1 ^= 1; else phase0 ^= 1;
        if (ok) {
            if (lane == 0) { ss[SS_PLC_LOST] = 0; ss[SS_PLC_ALPHA] = (int32_t)f2u(1.0f); }
        } else {
            int lost = ss[SS_PLC_LOST];
            float alpha = u2f((uint32_t)ss[SS_PLC_ALPHA]);
            uint32_t seed = (uint32_t)ss[SS_PLC_SEED];
            __syncwarp();                                            // every lane has read the state before any lane updates it
            if (lost >= 4) alpha = xm(alpha, lost < 8 ? 0.9f : 0.85f);
            // lane's first line is k = lane: LCG advanced lane+1 times; then jump 32 steps at a time
            uint32_t s = seed;
            for (int i = 0; i <= lane; i++) s = (16831u + s * 12821u) & 0xFFFFu;
            uint32_t a32 = 1, c32 = 0;                               // composition of 32 LCG steps
            for (int i = 0; i < 32; i++) { c32 = (16831u + c32 * 12821u) & 0xFFFFu; a32 = (a32 * 12821u) & 0xFFFFu; }
            for (int k = lane; k < NE; k += 32) {
                const float lg = P[k];
                P[k] = s < 0x8000u ? xm(lg, alpha) : xm(lg, -alpha);
                if (k == NE - 1) ss[SS_PLC_SEED] = (int32_t)s;
                s = (a32 * s + c32) & 0xFFFFu;
            }
            if (lane == 0) { ss[SS_PLC_LOST] = lost + 1; ss[SS_PLC_ALPHA] = (int32_t)f2u(alpha); }
            __syncwarp();
        }

        // ---- DCT-IV, unfold, window (dct_iv.rs:49-67, modified_dct.rs:97-136)
        const float* D = dct_iv_warp<NF>(P, Q, s_dtw, s_ftw, lane);
        float head_s[2 * NP], tail[2 * NPT];
        imdct_unfold<NF, MS10>(D, s_win - Z, lane, head_s, tail);
        __syncwarp();                                                // every lane is done reading P / Q before they are reused

        // ---- overlap-add (modified_dct.rs:138-151): two roundings, like the reference (and the time-parallel path)
#pragma unroll
        for (int j = 0; j < NPT; j++) {
            if (64 * j + 64 <= NF - Z || 64 * j + 2 * lane < NF - Z) {
                head_s[2 * j] = xa(ola_r[j].x, head_s[2 * j]);
                head_s[2 * j + 1] = xa(ola_r[j].y, head_s[2 * j + 1]);
                ((float2*)ola)[32 * j + lane] = make_float2(tail[2 * j], tail[2 * j + 1]);
            }
        }

        // ---- post filter, pass-through case (long_term_post_filter.rs:252-343 with zero coefficients): history + PCM
        // Inside an active span the ring block must keep its old content until the filter has run (long pitch lags reach
        // around the ring into it), so x_hat waits in a side buffer for ltpf_kernel instead.  When the handle's promised
        // minimum frame length rules the filter out for good (p.no_ltpf), the history is never read and is not kept.
        const bool in_span = active || prev_active;
        float* yblk = in_span ? p.ltpf_x + (size_t)stream * NF : p.ltpf_y + (size_t)stream * p.hist_len + blk_idx * NF;
        float* xtail = p.ltpf_xtail + (size_t)stream * XTAIL_FLOATS + blk_idx * 16;
        int16_t* out = p.pcm_out + (size_t)stream * p.pcm_stride;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const int n = 64 * j + 2 * lane;
            if (64 * j + 64 <= NF || n < NF) {
                const float a = head_s[2 * j], b = head_s[2 * j + 1];
                if (!p.no_ltpf) {
                    ((float2*)yblk)[32 * j + lane] = make_float2(a, b);
                    if (n >= NF - 16) ((float2*)xtail)[(n - (NF - 16)) >> 1] = make_float2(a, b);   // x_hat tail for the next frame's filter
                }
                const int32_t qa = round_pcm_i32(a), qb = round_pcm_i32(b);
                if (p.pcm_pairs) ((uint32_t*)out)[32 * j + lane] = ((uint32_t)qa & 0xffffu) | ((uint32_t)qb << 16);
                else { out[n] = (int16_t)qa; out[n + 1] = (int16_t)qb; }
            }
        }
        if (lane == 0 && !p.no_ltpf) {
            sd[SD_BLK] = blk_idx;
            int nb = blk_idx + 1;                                    // :334-337
            if (nb > (MS10 ? 1 : 2)) nb = 0;
            ss[SS_LTPF_BLK] = nb;
            if (!in_span) {                                          // otherwise ltpf_kernel carries the filter state on
                ss[SS_LTPF_PREV] = 4 << 8;
                ss[SS_LTPF_PINT] = 0;
                ss[SS_LTPF_PFR] = 0;
            }
        }
        stream = nstream;
        head = nhead;
        nstream = n2stream;
        nhead = n2head;
    }
}

// ---------------------------------------------------------------- kernel 3: the post filter where it is not idle
__global__ void __launch_bounds__(SYN_WARPS * 32) ltpf_kernel(SynthParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const DevConfig& c = *p.cfg;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int first = (blockIdx.x * SYN_WARPS + wid) * p.group;
    if (first >= p.n_streams) return;
    // which streams of this warp's group are inside an active span
    uint32_t todo;
    {
        const int s = first + lane;
        bool in_span = false;
        if (lane < p.group && s < p.n_streams)
            in_span = p.side[(size_t)s * SIDE_WORDS + SD_LTPF_ACTIVE] != 0 || (p.sstate[(size_t)s * SS_WORDS + SS_LTPF_PREV] & 1) != 0;
        todo = __ballot_sync(0xffffffffu, in_span);
    }
    if (todo == 0) return;
    const int nf = c.nf;
    const int blocks = c.ltpf_blocks, l_num = c.ltpf_l_num, l_den = c.ltpf_l_den, norm = c.ltpf_norm, s2p5 = c.ltpf_s2p5;
    float* Y = (float*)(smem + (size_t)wid * p.smem_per_warp);    // hist_len floats: LTPF circular history
    float* X = Y + p.hist_len;                                    // nf floats: x_hat, the filter input
    float* xt = X + nf;                                           // 16: x_hat tail of the previous frame
    float* scratch = xt + 16;                                     // l_num + norm floats: frozen history for transition case 5

    while (todo) {
        const int stream = first + (__ffs(todo) - 1);
        todo &= todo - 1;
        const int32_t* sd = p.side + (size_t)stream * SIDE_WORDS;
        int32_t* ss = p.sstate + (size_t)stream * SS_WORDS;
        const int active = sd[SD_LTPF_ACTIVE];
        const int nbits = sd[SD_NBITS];
        const int blk_idx = sd[SD_BLK];
        const int blk = blk_idx * nf;
        const int prev = ss[SS_LTPF_PREV];
        const int prev_active = prev & 1, prev_code = prev >> 8;
        const int p_int_mem = ss[SS_LTPF_PINT], p_fr_mem = ss[SS_LTPF_PFR];
        int p_int = 0, p_fr = 0, code = 4;
        if (active) {                                  // compute_filter_parameters :164-190, compute_gains_params :142-161
            const int pi = sd[SD_PITCH_INDEX];
            int pit;
            double pfr;
            if (pi >= 440) { pit = pi - 283; pfr = 0.0; }
            else if (pi >= 380) { pit = pi / 2 - 63; pfr = (double)(2 * pi - 4 * pit - 252); }
            else { pit = pi / 4 + 32; pfr = (double)(pi + 128 - 4 * pit); }
            const double pitch = (double)pit + pfr / 4.0;
            const double pitch_fs = pitch * (8000.0 * ceil((double)c.fs / 8000.0) / 12800.0);
            const int p_up = (int)((pitch_fs * 4.0) + 0.5);
            p_int = p_up / 4;
            p_fr = p_up - 4 * p_int;
            const int t_nbits = c.n_ms == LC3B_7P5MS ? (int)round((double)nbits * 10.0 / 7.5) : nbits;
            const int sf = c.fs_ind * 80;
            code = t_nbits < 320 + sf ? 0 : t_nbits < 400 + sf ? 1 : t_nbits < 480 + sf ? 2 : t_nbits < 560 + sf ? 3 : 4;
        }
        float* yhist = p.ltpf_y + (size_t)stream * p.hist_len;
        const float* xtail_prev = p.ltpf_xtail + (size_t)stream * XTAIL_FLOATS + (blk_idx == 0 ? blocks - 1 : blk_idx - 1) * 16;
        __syncwarp();                                  // the previous stream's reads of the shared buffers are done
        for (int n = lane; n < p.hist_len; n += 32) Y[n] = yhist[n];
        if (lane < 16) xt[lane] = xtail_prev[lane];
        __syncwarp();
        for (int n = lane; n < nf; n += 32) X[n] = p.ltpf_x[(size_t)stream * nf + n];   // x_hat left by synth_kernel
        __syncwarp();

        // coefficient sets: current (c_num, c_den) and previous (c_num_mem, c_den_mem); code 4 = all zero
        auto cnum = [&](int cd, int k) -> float { return cd < 4 ? c.ltpf_num[cd][k] : 0.0f; };
        auto cden = [&](int cd, int fr, int k) -> float { return cd < 4 ? c.ltpf_den[cd][fr][k] : 0.0f; };
        const int cur_code = active ? code : 4;     // inactive: zeroed coefficients (:196-201)
        const int mem_code = prev_active ? prev_code : 4;
        auto wrapi = [&](int idx) -> int { return idx < 0 ? idx + p.hist_len : idx; };   // :244-250
        auto x_at = [&](int pos) -> float {          // x_hat_mem[wrap(pos)], pos relative to buffer start
            const int rel = pos - blk;               // >= -l_num
            return rel >= 0 ? X[rel] : xt[16 + rel];
        };
        // compute_filter / compute_filter_mem :380-415
        auto filt = [&](int start, int pint, int cd, int fr) -> float {
            float out = 0.0f;
            for (int k = 0; k <= l_num; k++) out = xa(out, xm(cnum(cd, k), x_at(start - k)));
            const int sden = start - pint + l_den / 2;
            for (int k = 0; k <= l_den; k++) out = xs(out, xm(cden(cd, fr, k), Y[wrapi(sden - k)]));
            return out;
        };
        const float fnorm = (float)norm;
        // run `body(n)` for n in [n0, n1) in dependency-safe chunks: outputs reach back at least `reach` samples
        auto chunked = [&](int n0, int n1, int reach, auto body) {
            const int C = reach < 32 ? reach : 32;
            for (int s = n0; s < n1; s += C) {
                const int n = s + lane;
                if (lane < C && n < n1) body(n);
                __syncwarp();
            }
        };
        const int reach_cur = active ? p_int - l_den / 2 : 32;
        const int reach_mem = prev_active ? p_int_mem - l_den / 2 : 32;
        auto deactivate_first = [&]() {              // :417-424
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
            // activate_first_2p5ms_from_mem :345-378: numerator taps read a frozen copy of y[blk-l_num .. blk+norm)
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
        int16_t* out = p.pcm_out + (size_t)stream * p.pcm_stride;
        for (int n = lane; n < nf; n += 32) {
            const float v = Y[blk + n];
            yhist[blk + n] = v;
            out[n] = (int16_t)round_pcm_i32(v);
        }
        if (lane == 0) {
            ss[SS_LTPF_PREV] = (active ? 1 : 0) | ((active ? code : 4) << 8);
            ss[SS_LTPF_PINT] = p_int;
            ss[SS_LTPF_PFR] = p_fr;
        }
    }
}

static size_t ltpf_warp_bytes(const lc3b_config& c) {
    const int blocks = c.n_ms == LC3B_10MS ? 2 : 3;
    size_t b = (size_t)(blocks * c.nf + c.nf + 16 + 16 + c.nf / 3 + 16) * 4;
    return (b + 15) & ~(size_t)15;
}

namespace {
struct PrepareSynth {
    cudaError_t e = cudaSuccess;
    template <int NF, bool MS10> void operator()() {
        e = cudaFuncSetAttribute(synth_kernel<NF, MS10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)synth_smem_bytes(NF));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(synth_classic_kernel<NF, MS10>, cudaFuncAttributeMaxDynamicSharedMemorySize, SYN_WARPS * 2 * NF * 4);
    }
};
struct PlanSynth {
    LaunchPlan& plan;
    const SynthParams& p;
    int dep, node, sm_count;
    bool pipelined;
    template <int NF, bool MS10> void operator()() {
        // persistent: at most as many CTAs as fit on the device at once (4 of 8 warps per SM by registers; shared memory may allow fewer)
        const int full = (p.n_streams + SYN_WARPS - 1) / SYN_WARPS;
        if (!pipelined) {
            node = plan.add(synth_classic_kernel<NF, MS10>, (unsigned)full, SYN_WARPS * 32, SYN_WARPS * 2 * NF * 4, p, dep);
            return;
        }
        const size_t smem = synth_smem_bytes(NF);
        int per_sm = (int)((227 * 1024) / (smem + 1024));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        const int resident = sm_count * per_sm;
        const int ctas = (p.n_streams + PIPE_WARPS - 1) / PIPE_WARPS;
        node = plan.add(synth_kernel<NF, MS10>, (unsigned)(ctas < resident ? ctas : resident), PIPE_WARPS * 32, smem, p, dep);
    }
};
}  // namespace

cudaError_t prepare_synth(const DecoderState& st) {
    PrepareSynth ps;
    if (!dispatch_frame_geo(st.cfg.nf, st.cfg.n_ms == LC3B_10MS, ps)) return cudaErrorInvalidValue;
    if (ps.e != cudaSuccess) return ps.e;
    return cudaFuncSetAttribute(ltpf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ltpf_warp_bytes(st.cfg) * SYN_WARPS));
}

// Adds the synthesis kernel (after node `dep`, -1 = none) and the post-filter kernel behind it; returns the last node.
int plan_synth(LaunchPlan& plan, const DecoderState& st, int16_t* pcm_out, size_t pcm_stride, int dep, int base, int count) {
    SynthParams p;
    memset(&p, 0, sizeof(p));
    p.cfg = st.dcfg;
    p.win = st.win;
    p.dtw = st.dtw;
    p.ftw = st.ftw;
    // a sub-batch: every per-stream pointer moved to its first stream (see entropy_params)
    const size_t b = (size_t)base;
    const int n = count < 0 ? st.n_streams : count;
    p.hist_len = (st.cfg.n_ms == LC3B_10MS ? 2 : 3) * st.cfg.nf;
    p.spec = st.spec + b * st.cfg.ne;
    p.ola = st.ola + b * (size_t)(st.cfg.nf - st.cfg.z);
    p.ltpf_y = st.ltpf_y + b * (size_t)p.hist_len;
    p.ltpf_xtail = st.ltpf_xtail + b * XTAIL_FLOATS;
    p.ltpf_x = st.ltpf_x + b * (size_t)st.cfg.nf;
    p.side = st.side + b * SIDE_WORDS;
    p.sstate = st.sstate + b * SS_WORDS;
    p.pcm_out = pcm_out + b * pcm_stride;
    p.pcm_stride = pcm_stride;
    p.n_streams = n;
    p.slot_streams = st.n_streams;
    p.pcm_pairs = (((uintptr_t)p.pcm_out & 3) == 0 && (pcm_stride & 1) == 0) ? 1 : 0;
    // ltpf_kernel: one warp per stream while that keeps the grid small, else a warp walks a group of streams
    int group = 1;
    while (group < 32 && n / group > 16384) group *= 2;
    p.group = group;
    const size_t lw = ltpf_warp_bytes(st.cfg);
    p.smem_per_warp = (int)lw;
    p.no_ltpf = st.no_ltpf;
    // which synthesis kernel: 0 = one warp per frame, ordinary loads (default: measured faster, the kernel is issue bound);
    // 1 = persistent warps with the next frame's spectrum prefetched by the TMA unit (lc3b_decoder_set_synth_mode / LC3B_SYNTH=pipe)
    static const int env_mode = [] { const char* e = getenv("LC3B_SYNTH"); return !e ? -1 : (e[0] == 'p' ? 1 : 0); }();
    const int mode = env_mode >= 0 ? env_mode : st.synth_mode;
    PlanSynth ps{plan, p, dep, -1, st.sm_count > 0 ? st.sm_count : 148, mode == 1};
    if (!dispatch_frame_geo(st.cfg.nf, st.cfg.n_ms == LC3B_10MS, ps)) return -1;
    if (st.no_ltpf) return ps.node;                           // the post filter can never be active: nothing to launch
    const int n_warps = (n + group - 1) / group;
    return plan.add(ltpf_kernel, (unsigned)((n_warps + SYN_WARPS - 1) / SYN_WARPS), SYN_WARPS * 32, lw * SYN_WARPS, p, ps.node);
}

cudaError_t launch_synth(const DecoderState& st, int16_t* pcm_out, size_t pcm_stride, cudaStream_t stream) {
    LaunchPlan plan;
    if (plan_synth(plan, st, pcm_out, pcm_stride, -1) < 0) return cudaErrorInvalidValue;
    return plan_launch_direct(plan, stream);
}

}  // namespace lc3b
