// lc3b engine: mixed-rate batch decoder behind the C ABI (include/lc3b.h, lc3b_mixed_*; BASELINE config 4).
//
// The reference builds one Lc3Decoder per (frame duration, sampling frequency) - Lc3Decoder::new takes ONE of each for
// all its channels (src/decoder/lc3_decoder.rs:181) - so a mixed population is a set of decoders.  This handle is that
// set for up to twelve configurations, driven as ONE batch: streams are bucketed by configuration (stable order), the
// caller lays its rows out in bucket order, and a call is
//     1 entropy launch over every bucket  ->  1 dequantisation launch per frame duration present
//     ->  per bucket: synthesis kernel -> post-filter kernel   (frame length is a template constant there)
// issued as one cached CUDA graph, so the per-bucket kernels run side by side without fork/join streams.
// Each bucket owns an ordinary single-configuration decoder state (same kernels, same bits as lc3b_decode_frames).
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "lc3b_common.cuh"
#include "lc3b_handles.cuh"

using namespace lc3b;

namespace {

struct Bucket {
    int sf = 0, fd = 0, first_row = 0, count = 0;
    lc3b_config cfg{};
    size_t ws_off = 0, ws_bytes = 0;
    size_t host_pcm_off = 0;          // int16 elements from the start of a dense per-bucket PCM buffer
    lc3b_decoder* dec = nullptr;
};

struct Layout {
    std::vector<int32_t> order;       // row -> original stream id
    std::vector<Bucket> buckets;
    size_t host_pcm_elems = 0;
    // device workspace carving
    size_t tables_off = 0, stage_in_off = 0, stage_len_off = 0, stage_status_off = 0, stage_out_off = 0, total = 0;
};

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// bucket order: configurations sorted by (sampling frequency, frame duration), streams of a bucket in their original order
bool make_layout(int n_streams, const int32_t* sf, const int32_t* fd, int max_nbytes, Layout& L) {
    if (n_streams <= 0 || !sf || !fd) return false;
    std::vector<int> count(12, 0);
    for (int s = 0; s < n_streams; s++) {
        if (sf[s] < 0 || sf[s] > 5 || fd[s] < 0 || fd[s] > 1) return false;
        count[(size_t)(sf[s] * 2 + fd[s])]++;
    }
    L.order.assign((size_t)n_streams, 0);
    std::vector<int> next(12, 0);
    int row = 0;
    size_t off = 0, pcm = 0;
    for (int k = 0; k < 12; k++) {
        if (!count[(size_t)k]) continue;
        Bucket b;
        b.sf = k / 2;
        b.fd = k % 2;
        b.first_row = row;
        b.count = count[(size_t)k];
        if (!config_new(b.sf, b.fd, &b.cfg)) return false;
        if (max_nbytes > 0) {
            if (decoder_workspace_bytes(b.count, b.fd, b.sf, max_nbytes, false, &b.ws_bytes) != LC3B_OK) return false;
            b.ws_off = off;
            off += align256(b.ws_bytes);
        }
        b.host_pcm_off = pcm;
        pcm += (size_t)b.count * (size_t)b.cfg.nf;
        next[(size_t)k] = row;
        row += b.count;
        L.buckets.push_back(b);
    }
    for (int s = 0; s < n_streams; s++) L.order[(size_t)next[(size_t)(sf[s] * 2 + fd[s])]++] = s;
    L.host_pcm_elems = pcm;
    const size_t ns = (size_t)n_streams;
    L.tables_off = off;           off += align256(sizeof(MixedBucket) * 3 * 12);
    L.stage_in_off = off;         off += align256(ns * (size_t)(max_nbytes > 0 ? max_nbytes : 0));
    L.stage_len_off = off;        off += align256(sizeof(int32_t) * ns);
    L.stage_status_off = off;     off += align256(sizeof(int32_t) * ns);
    L.stage_out_off = off;        off += align256(sizeof(int16_t) * pcm * 2);    // double-buffered (host pipelining)
    L.total = off;
    return true;
}

}  // namespace

struct lc3b_mixed_decoder {
    int n_streams = 0, max_nbytes = 0, device = 0, max_nf = 0;
    Layout L;
    uint8_t* base = nullptr;
    MixedTables tables{};
    int graph_mode = 1;
    GraphCache graphs;
    int pipelined = 0, buf = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t compute_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    bool d2h_pending[2] = {false, false};
};

// the call's kernels: rows of `frames` / `pcm_out` in bucket order; pcm_out either one common pitch (device entry) or dense
// per bucket at the layout's host_pcm_off (host entry: dense rows keep the read-back ONE linear copy)
static cudaError_t run_mixed(lc3b_mixed_decoder* h, const uint8_t* frames, const int32_t* frame_nbytes, int nbytes,
                             size_t frame_stride, int16_t* pcm_out, size_t pcm_stride, bool dense_pcm, int32_t* status_out,
                             cudaStream_t stream) {
    LaunchPlan plan;
    int d10 = -1, d75 = -1;
    plan_entropy_mixed(plan, h->tables, frames, frame_nbytes, nbytes, frame_stride, status_out, &d10, &d75);
    for (const Bucket& b : h->L.buckets) {
        int16_t* out = dense_pcm ? pcm_out + b.host_pcm_off : pcm_out + (size_t)b.first_row * pcm_stride;
        const size_t pitch = dense_pcm ? (size_t)b.cfg.nf : pcm_stride;
        if (plan_synth(plan, b.dec->st, out, pitch, b.fd == LC3B_10MS ? d10 : d75) < 0) return cudaErrorInvalidValue;
    }
    return h->graph_mode ? plan_launch_graph(h->graphs, plan, stream) : plan_launch_direct(plan, stream);
}

extern "C" {

int lc3b_mixed_decoder_layout(int n_streams, const int32_t* sampling_frequency, const int32_t* frame_duration, int32_t* order,
                              lc3b_mixed_bucket* buckets, int32_t* n_buckets, uint64_t* host_pcm_elems) {
    Layout L;
    if (!make_layout(n_streams, sampling_frequency, frame_duration, 0, L)) return LC3B_ERR_INVALID_ARG;
    if (order) memcpy(order, L.order.data(), sizeof(int32_t) * (size_t)n_streams);
    if (buckets) {
        for (size_t i = 0; i < L.buckets.size(); i++) {
            const Bucket& b = L.buckets[i];
            buckets[i].sampling_frequency = b.sf;
            buckets[i].frame_duration = b.fd;
            buckets[i].first_row = b.first_row;
            buckets[i].n_rows = b.count;
            buckets[i].nf = b.cfg.nf;
            buckets[i].reserved = 0;
            buckets[i].host_pcm_offset = (uint64_t)b.host_pcm_off;
        }
    }
    if (n_buckets) *n_buckets = (int32_t)L.buckets.size();
    if (host_pcm_elems) *host_pcm_elems = (uint64_t)L.host_pcm_elems;
    return LC3B_OK;
}

int lc3b_mixed_decoder_workspace_bytes(int n_streams, const int32_t* sampling_frequency, const int32_t* frame_duration,
                                       int max_nbytes, size_t* device_bytes) {
    Layout L;
    if (!device_bytes || max_nbytes <= 0 || max_nbytes > MAX_NBYTES ||
        !make_layout(n_streams, sampling_frequency, frame_duration, max_nbytes, L))
        return LC3B_ERR_INVALID_ARG;
    *device_bytes = L.total;
    return LC3B_OK;
}

void lc3b_mixed_decoder_destroy(lc3b_mixed_decoder* h) {
    if (!h) return;
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        for (int i = 0; i < 2; i++) { cudaEventDestroy(h->compute_done[i]); cudaEventDestroy(h->d2h_done[i]); }
        cudaStreamDestroy(h->copy_stream);
    }
    for (Bucket& b : h->L.buckets) lc3b_decoder_destroy(b.dec);
    delete h;
}

int lc3b_mixed_decoder_init(lc3b_mixed_decoder** out, int n_streams, const int32_t* sampling_frequency,
                            const int32_t* frame_duration, int max_nbytes, int device, void* dev_workspace,
                            size_t workspace_bytes, void* cuda_stream) {
    if (!out || !dev_workspace || max_nbytes <= 0 || max_nbytes > MAX_NBYTES) return LC3B_ERR_INVALID_ARG;
    lc3b_mixed_decoder* h = new (std::nothrow) lc3b_mixed_decoder();
    if (!h) return LC3B_ERR_INVALID_ARG;
    if (!make_layout(n_streams, sampling_frequency, frame_duration, max_nbytes, h->L)) { delete h; return LC3B_ERR_INVALID_ARG; }
    if (workspace_bytes < h->L.total || ((uintptr_t)dev_workspace & 255) != 0) { delete h; return LC3B_ERR_WORKSPACE; }
    DeviceGuard guard;
    h->n_streams = n_streams;
    h->max_nbytes = max_nbytes;
    h->device = device;
    h->base = (uint8_t*)dev_workspace;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    // one ordinary decoder state per bucket, carved from the caller's workspace
    MixedBucket host_tab[3][12];
    memset(host_tab, 0, sizeof(host_tab));
    int n_all = 0, n_10 = 0, n_75 = 0, cta_all = 0, cta_10 = 0, cta_75 = 0;
    for (Bucket& b : h->L.buckets) {
        const int rc = decoder_init(&b.dec, b.count, b.fd, b.sf, max_nbytes, device, h->base + b.ws_off, b.ws_bytes, cuda_stream, false);
        if (rc != LC3B_OK) { lc3b_mixed_decoder_destroy(h); return rc; }
        b.dec->graph_mode = 0;
        h->max_nf = b.cfg.nf > h->max_nf ? b.cfg.nf : h->max_nf;
        const int ctas = (b.count + 127) / 128;
        MixedBucket mb;
        memset(&mb, 0, sizeof(mb));
        mb.ep = entropy_params(b.dec->st, nullptr, nullptr, max_nbytes, (size_t)max_nbytes, nullptr);
        mb.first_row = b.first_row;
        mb.first_cta = cta_all;
        host_tab[0][n_all++] = mb;
        cta_all += ctas;
        if (b.fd == LC3B_10MS) { mb.first_cta = cta_10; host_tab[1][n_10++] = mb; cta_10 += ctas; }
        else { mb.first_cta = cta_75; host_tab[2][n_75++] = mb; cta_75 += ctas; }
    }
    MixedBucket* dev_tab = (MixedBucket*)(h->base + h->L.tables_off);
    if (cudaSetDevice(device) != cudaSuccess) { lc3b_mixed_decoder_destroy(h); return cuda_fail(cudaGetLastError()); }
    cudaError_t e = cudaMemcpyAsync(dev_tab, host_tab, sizeof(host_tab), cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);                 // host_tab is a stack object
    if (e != cudaSuccess) { lc3b_mixed_decoder_destroy(h); return cuda_fail(e); }
    h->tables.all = dev_tab;          h->tables.n_all = n_all; h->tables.cta_all = cta_all;
    h->tables.b10 = dev_tab + 12;     h->tables.n_10 = n_10;   h->tables.cta_10 = cta_10;
    h->tables.b75 = dev_tab + 24;     h->tables.n_75 = n_75;   h->tables.cta_75 = cta_75;
    h->tables.n_streams = n_streams;
    const char* env = getenv("LC3B_GRAPH");
    h->graph_mode = (env && env[0] == '0') ? 0 : 1;
    *out = h;
    return LC3B_OK;
}

int lc3b_mixed_decoder_set_dequant_mode(lc3b_mixed_decoder* h, int mode) {
    if (!h || mode < 0 || mode > 2) return LC3B_ERR_INVALID_ARG;
    h->tables.dequant_mode = mode;
    return LC3B_OK;
}

int lc3b_mixed_decoder_set_graph_mode(lc3b_mixed_decoder* h, int mode) {
    if (!h || mode < 0 || mode > 1) return LC3B_ERR_INVALID_ARG;
    h->graph_mode = mode;
    return LC3B_OK;
}

int lc3b_mixed_decode_frames(lc3b_mixed_decoder* h, int bits_per_sample, const uint8_t* frames, const int32_t* frame_nbytes,
                             int nbytes, size_t frame_stride, int16_t* pcm_out, size_t pcm_stride, int32_t* status_out,
                             void* cuda_stream) {
    if (!h || !frames || !frame_nbytes || !pcm_out) return LC3B_ERR_INVALID_ARG;
    if (bits_per_sample != 16) return LC3B_ERR_BITS_PER_SAMPLE;              // lc3_decoder.rs:80
    if (nbytes <= 0 || nbytes > h->max_nbytes || (size_t)nbytes > frame_stride || pcm_stride < (size_t)h->max_nf)
        return LC3B_ERR_INVALID_ARG;
    LC3B_CU(run_mixed(h, frames, frame_nbytes, nbytes, frame_stride, pcm_out, pcm_stride, false, status_out, (cudaStream_t)cuda_stream));
    return LC3B_OK;
}

int lc3b_mixed_decoder_set_host_pipelining(lc3b_mixed_decoder* h, int on) {
    if (!h) return LC3B_ERR_INVALID_ARG;
    if (!on && h->pipelined) {
        for (int i = 0; i < 2; i++)
            if (h->d2h_pending[i]) { LC3B_CU(cudaEventSynchronize(h->d2h_done[i])); h->d2h_pending[i] = false; }
    }
    if (on && !h->copy_stream) {
        DeviceGuard guard;
        LC3B_CU(cudaSetDevice(h->device));
        LC3B_CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            LC3B_CU(cudaEventCreateWithFlags(&h->compute_done[i], cudaEventDisableTiming));
            LC3B_CU(cudaEventCreateWithFlags(&h->d2h_done[i], cudaEventDisableTiming));
        }
    }
    h->pipelined = on ? 1 : 0;
    return LC3B_OK;
}

int lc3b_mixed_decoder_host_fence(lc3b_mixed_decoder* h, void* cuda_stream) {
    if (!h) return LC3B_ERR_INVALID_ARG;
    for (int i = 0; i < 2; i++)
        if (h->d2h_pending[i]) LC3B_CU(cudaStreamWaitEvent((cudaStream_t)cuda_stream, h->d2h_done[i], 0));
    return LC3B_OK;
}

int lc3b_mixed_decode_frames_host(lc3b_mixed_decoder* h, int bits_per_sample, const uint8_t* frames, const int32_t* frame_nbytes,
                                  int nbytes, size_t frame_stride, int16_t* pcm_out, int32_t* status_out, void* cuda_stream) {
    if (!h || !frames || !frame_nbytes || !pcm_out) return LC3B_ERR_INVALID_ARG;
    if (bits_per_sample != 16) return LC3B_ERR_BITS_PER_SAMPLE;
    if (nbytes <= 0 || nbytes > h->max_nbytes || (size_t)nbytes > frame_stride) return LC3B_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const size_t ns = (size_t)h->n_streams;
    uint8_t* stage_in = h->base + h->L.stage_in_off;
    int32_t* stage_len = (int32_t*)(h->base + h->L.stage_len_off);
    int32_t* stage_status = (int32_t*)(h->base + h->L.stage_status_off);
    int16_t* stage_out = (int16_t*)(h->base + h->L.stage_out_off) + (h->pipelined ? (size_t)h->buf * h->L.host_pcm_elems : 0);
    if (frame_stride == (size_t)nbytes) LC3B_CU(cudaMemcpyAsync(stage_in, frames, ns * (size_t)nbytes, cudaMemcpyHostToDevice, stream));
    else LC3B_CU(cudaMemcpy2DAsync(stage_in, (size_t)nbytes, frames, frame_stride, (size_t)nbytes, ns, cudaMemcpyHostToDevice, stream));
    LC3B_CU(cudaMemcpyAsync(stage_len, frame_nbytes, sizeof(int32_t) * ns, cudaMemcpyHostToDevice, stream));
    cudaStream_t out_stream = stream;
    if (h->pipelined) {
        if (h->d2h_pending[h->buf]) LC3B_CU(cudaStreamWaitEvent(stream, h->d2h_done[h->buf], 0));
        out_stream = h->copy_stream;
    }
    LC3B_CU(run_mixed(h, stage_in, stage_len, nbytes, (size_t)nbytes, stage_out, 0, true, status_out ? stage_status : nullptr, stream));
    if (h->pipelined) {
        LC3B_CU(cudaEventRecord(h->compute_done[h->buf], stream));
        LC3B_CU(cudaStreamWaitEvent(out_stream, h->compute_done[h->buf], 0));
    }
    LC3B_CU(cudaMemcpyAsync(pcm_out, stage_out, h->L.host_pcm_elems * sizeof(int16_t), cudaMemcpyDeviceToHost, out_stream));
    if (h->pipelined) {
        LC3B_CU(cudaEventRecord(h->d2h_done[h->buf], out_stream));
        h->d2h_pending[h->buf] = true;
        h->buf ^= 1;
    }
    if (status_out) LC3B_CU(cudaMemcpyAsync(status_out, stage_status, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, stream));
    return LC3B_OK;
}

}  // extern "C"
