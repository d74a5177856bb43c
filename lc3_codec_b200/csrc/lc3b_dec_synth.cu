// lc3b engine, decoder kernel 2 of 2: shaped spectrum -> PCM, one WARP per frame.
//
// Replaces, per stream, the second half of DecoderChannel::decode (src/decoder/lc3_decoder.rs:134-153):
//   PacketLossConcealment::save / load_into   src/decoder/packet_loss_concealment.rs:50,63
//   ModDiscreteCosTrans::run                  src/decoder/modified_dct.rs:76 (DCT-IV via N = nf/2 complex FFT,
//                                             src/common/dct_iv.rs:49, src/common/kissfft.rs:78; window; overlap-add)
//   LongTermPostFilter::run                   src/decoder/long_term_post_filter.rs:252 (five transition cases)
//   output_scaling::scale_and_round           src/decoder/output_scaling.rs:13
// Data-parallel work: 32 lanes walk the frame's lines/samples, all global traffic is stream-major and
// coalesced (spectrum 4*ne B in, overlap memory 4*(nf-z) B in/out, LTPF history 4*nf B out, PCM 2*nf B out).
// The FFT is a shared-memory Stockham autosort with radix-2/3/4/5 butterflies (a different factor order than
// kissfft: the PCM contract is +-1 LSB, not bit equality).  The LTPF recursion only reaches back
// p_int - l_den/2 >= 18 samples, so up to 32 consecutive outputs are computed in parallel per step.
#include "lc3b_common.cuh"
#include "lc3b_fft.cuh"
#include "lc3b_math.cuh"

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
    const int32_t* side;
    int32_t* sstate;
    int16_t* pcm_out;
    size_t pcm_stride;
    int n_streams;
    int hist_len;        // ltpf_blocks * nf
    int y_floats;        // max(hist_len, 2 * nf): the FFT ping-pong buffers and the LTPF history share this space
    int smem_per_warp;   // bytes
};

constexpr int SYN_WARPS = 4;

__global__ void __launch_bounds__(SYN_WARPS * 32) synth_kernel(SynthParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const DevConfig& c = *p.cfg;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int stream = blockIdx.x * SYN_WARPS + wid;
    if (stream >= p.n_streams) return;
    const int nf = c.nf, ne = c.ne, z = c.z, N = c.n_fft, h = nf / 2;

    uint8_t* base = smem + (size_t)wid * p.smem_per_warp;
    float2* bufA = (float2*)base;                 // N complex
    float2* bufB = bufA + N;                      // N complex
    float* Y = (float*)base;                      // hist_len floats: LTPF circular history, over the (then dead) FFT buffers
    float* X = (float*)base + p.y_floats;         // nf floats: spectrum in, then DCT-IV output, then time samples
    float* scratch = X + nf;                      // l_num + norm floats: frozen history for transition case 5

    const int32_t* sd = p.side + (size_t)stream * SIDE_WORDS;
    int32_t* ss = p.sstate + (size_t)stream * SS_WORDS;
    const int ok = sd[SD_OK];
    const int slot = sd[SD_SLOT];
    const int nbits = sd[SD_NBITS];
    const float* sp = p.spec + ((size_t)slot * p.n_streams + stream) * ne;

    // ---- spectrum load, with concealment (packet_loss_concealment.rs:63-85) when the frame was bad
    // overlap memory: requested now, consumed after the FFT (ceil(300 / 32) = 10 values per lane at most)
    float* ola = p.ola + (size_t)stream * (nf - z);
    float ola_r[10];
#pragma unroll
    for (int j = 0; j < 10; j++) { const int n = lane + 32 * j; ola_r[j] = n < nf - z ? ola[n] : 0.0f; }
    if (ok) {
        for (int k4 = lane; k4 < nf / 4; k4 += 32)
            ((float4*)X)[k4] = 4 * k4 < ne ? ((const float4*)sp)[k4] : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
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
        for (int k = lane; k < nf; k += 32) {
            float v = 0.0f;
            if (k < ne) {
                const float lg = sp[k];
                v = s < 0x8000u ? xm(lg, alpha) : xm(lg, -alpha);
                if (k == ne - 1) ss[SS_PLC_SEED] = (int32_t)s;
                s = (a32 * s + c32) & 0xFFFFu;
            }
            X[k] = v;
        }
        if (lane == 0) { ss[SS_PLC_LOST] = lost + 1; ss[SS_PLC_ALPHA] = (int32_t)f2u(alpha); }
    }
    __syncwarp();

    // ---- DCT-IV: pre-twiddle, N-point FFT, post-twiddle (dct_iv.rs:49-67)
    dct_iv_warp(X, bufA, bufB, p.dtw, p.ftw, c.fft_radix, nf, N, lane);

    // ---- unfold + window + overlap-add (modified_dct.rs:97-151); t[m] below is the reference's t_hat_mdct[m] / gain
    auto t_at = [&](int m) -> float {
        if (m < h) return X[h + m];
        if (m < nf) return -X[nf - 1 - (m - h)];
        if (m < nf + h) return -X[h - 1 - (m - nf)];
        return -X[m - 3 * h];
    };
    float* T = (float*)bufA;                       // reuse: nf output samples fit in N complex
#pragma unroll
    for (int j = 0; j < 15; j++) {
        const int n = lane + 32 * j;
        if (n < nf) {
            float o;
            if (n < nf - z) o = xa(j < 10 ? ola_r[j < 10 ? j : 0] : 0.0f, xm(t_at(z + n), p.win[z + n]));   // two roundings, like the reference (and the time-parallel path)
            else o = t_at(nf + (n - (nf - z))) * p.win[nf + (n - (nf - z))];
            T[n] = o;
        }
    }
    __syncwarp();                                  // all reads of the old X-derived values for the first half are done
    for (int n = lane; n < nf - z; n += 32) ola[n] = t_at(nf + z + n) * p.win[nf + z + n];
    __syncwarp();
    for (int n = lane; n < nf; n += 32) X[n] = T[n];   // X = x_hat (input of the post filter)
    const float xtail_next = T[nf - 16 + (lane & 15)]; // last 16 samples of the filter INPUT, kept before Y reuses the buffer
    __syncwarp();

    // ---- long term post filter (long_term_post_filter.rs:252-343)
    const int active = sd[SD_LTPF_ACTIVE];
    const int prev = ss[SS_LTPF_PREV];
    const int prev_active = prev & 1, prev_code = prev >> 8;
    const int blocks = c.ltpf_blocks, l_num = c.ltpf_l_num, l_den = c.ltpf_l_den, norm = c.ltpf_norm, s2p5 = c.ltpf_s2p5;
    const int blk = ss[SS_LTPF_BLK] * nf;
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
    const int p_int_mem = ss[SS_LTPF_PINT], p_fr_mem = ss[SS_LTPF_PFR];
    float* yhist = p.ltpf_y + (size_t)stream * p.hist_len;
    float* xtail = p.ltpf_xtail + (size_t)stream * 16;

    if (!active && !prev_active) {                 // case 1: pass-through, history only
        for (int n = lane; n < nf; n += 32) yhist[blk + n] = X[n];
    } else {
        // coefficient sets: current (c_num, c_den) and previous (c_num_mem, c_den_mem); code 4 = all zero
        auto cnum = [&](int cd, int k) -> float { return cd < 4 ? c.ltpf_num[cd][k] : 0.0f; };
        auto cden = [&](int cd, int fr, int k) -> float { return cd < 4 ? c.ltpf_den[cd][fr][k] : 0.0f; };
        const int cur_code = active ? code : 4;     // inactive: zeroed coefficients (:196-201)
        const int mem_code = prev_active ? prev_code : 4;
        for (int n = lane; n < p.hist_len; n += 32) Y[n] = yhist[n];
        __syncwarp();
        auto wrapi = [&](int idx) -> int { return idx < 0 ? idx + p.hist_len : idx; };   // :244-250
        auto x_at = [&](int pos) -> float {          // x_hat_mem[wrap(pos)], pos relative to buffer start
            const int rel = pos - blk;               // >= -l_num
            return rel >= 0 ? X[rel] : xtail[16 + rel];
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
        for (int n = lane; n < nf; n += 32) { const float v = Y[blk + n]; yhist[blk + n] = v; X[n] = v; }
    }
    __syncwarp();
    // x tail for the next frame: last 16 samples of this frame's x_hat (the filter INPUT; X may hold the output by now)
    if (lane < 16) xtail[lane] = xtail_next;
    if (lane == 0) {
        ss[SS_LTPF_PREV] = (active ? 1 : 0) | ((active ? code : 4) << 8);
        ss[SS_LTPF_PINT] = p_int;
        ss[SS_LTPF_PFR] = p_fr;
        int nb = ss[SS_LTPF_BLK] + 1;                // :334-337
        if (nb * nf > (blocks - 1) * nf) nb = 0;
        ss[SS_LTPF_BLK] = nb;
    }

    // ---- output_scaling.rs:13-26: round half away from zero, saturate
    int16_t* out = p.pcm_out + (size_t)stream * p.pcm_stride;
    for (int n = lane; n < nf; n += 32) {
        const float v = X[n];
        int32_t q = v > 0.0f ? cast_i32(xa(v, 0.5f)) : cast_i32(xs(v, 0.5f));
        q = q > 32767 ? 32767 : q < -32768 ? -32768 : q;
        out[n] = (int16_t)q;
    }
}

static size_t synth_warp_bytes(const lc3b_config& c, int* hist_len, int* y_floats) {
    const int nf = c.nf, N = nf / 2;
    const int blocks = c.n_ms == LC3B_10MS ? 2 : 3;
    const int hl = blocks * nf;
    // [FFT ping-pong buffers | LTPF history] + spectrum/time samples + case-5 scratch
    const int yf = hl > 4 * N ? hl : 4 * N;
    size_t per_warp = (size_t)yf * 4 + (size_t)nf * 4 + (size_t)(16 + nf / 3 + 16) * 4;
    per_warp = (per_warp + 15) & ~(size_t)15;
    if (hist_len) *hist_len = hl;
    if (y_floats) *y_floats = yf;
    return per_warp;
}

cudaError_t prepare_synth(const DecoderState& st) {
    return cudaFuncSetAttribute(synth_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(synth_warp_bytes(st.cfg, nullptr, nullptr) * SYN_WARPS));
}

cudaError_t launch_synth(const DecoderState& st, int16_t* pcm_out, size_t pcm_stride, cudaStream_t stream) {
    SynthParams p;
    p.cfg = st.dcfg;
    p.win = st.win;
    p.dtw = st.dtw;
    p.ftw = st.ftw;
    p.spec = st.spec;
    p.ola = st.ola;
    p.ltpf_y = st.ltpf_y;
    p.ltpf_xtail = st.ltpf_xtail;
    p.side = st.side;
    p.sstate = st.sstate;
    p.pcm_out = pcm_out;
    p.pcm_stride = pcm_stride;
    p.n_streams = st.n_streams;
    const size_t per_warp = synth_warp_bytes(st.cfg, &p.hist_len, &p.y_floats);
    p.smem_per_warp = (int)per_warp;
    const int grid = (st.n_streams + SYN_WARPS - 1) / SYN_WARPS;
    synth_kernel<<<grid, SYN_WARPS * 32, per_warp * SYN_WARPS, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace lc3b
