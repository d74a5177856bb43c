// lc3b engine: one batch over the GPUs of a box (include/lc3b.h, lc3b_sharded_*; SURVEY.md 8e, BASELINE config 5).
//
// Streams never interact - DecoderChannel / EncoderChannel own all of their state (src/decoder/lc3_decoder.rs:62-69,
// src/encoder/lc3_encoder.rs:42-60) - so a batch shards by stream id with nothing to exchange: stream s belongs to
// shard floor(s * G / N), a contiguous block per GPU.  No collective, no peer access, no NCCL.
// A sharded handle owns, per GPU: an ordinary decoder / encoder handle on its own device workspace (allocated once, at
// create), a CUDA stream, and ONE HOST THREAD that issues that GPU's copies and launches, so the GPUs are fed in
// parallel and no device waits for another's launch latency.  A call hands every thread its row range of the caller's
// host buffers and returns; lc3b_sharded_*_wait joins all outstanding calls.  The host entry points of the per-GPU
// handles run with pipelining on, so the PCM read-back of call i overlaps the kernels of call i+1 on every GPU.
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "lc3b_common.cuh"
#include "lc3b_handles.cuh"

using namespace lc3b;

namespace {

struct Job {
    int kind = 0;                       // 0 decode, 1 encode, 2 flush (fence + drain), 3 quit
    const uint8_t* frames = nullptr;    // decode in
    uint8_t* frames_out = nullptr;      // encode out
    const int32_t* frame_nbytes = nullptr;
    int nbytes = 0;
    size_t frame_stride = 0;
    const int16_t* pcm_in = nullptr;    // encode in
    int16_t* pcm_out = nullptr;         // decode out
    size_t pcm_stride = 0;
    int32_t* status_out = nullptr;
};

struct Shard {
    int device = 0, first = 0, count = 0;
    void* workspace = nullptr;
    lc3b_decoder* dec = nullptr;
    lc3b_encoder* enc = nullptr;
    cudaStream_t stream = nullptr;
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> queue;
    uint64_t submitted = 0, completed = 0;
    int error = LC3B_OK, cuda_error = 0;
};

struct Pool {
    bool is_encoder = false;
    int n_streams = 0, nf = 0, max_nbytes = 0;
    std::vector<Shard*> shards;
};

void worker(Pool* pool, Shard* sh) {
    cudaSetDevice(sh->device);
    for (;;) {
        Job job;
        {
            std::unique_lock<std::mutex> lk(sh->mu);
            sh->cv_job.wait(lk, [&] { return !sh->queue.empty(); });
            job = sh->queue.front();
            sh->queue.pop_front();
        }
        int rc = LC3B_OK;
        if (job.kind == 0) {
            rc = lc3b_decode_frames_host(sh->dec, 16, job.frames + (size_t)sh->first * job.frame_stride,
                                         job.frame_nbytes ? job.frame_nbytes + sh->first : nullptr, job.nbytes, job.frame_stride,
                                         job.pcm_out + (size_t)sh->first * job.pcm_stride, job.pcm_stride,
                                         job.status_out ? job.status_out + sh->first : nullptr, sh->stream);
        } else if (job.kind == 1) {
            rc = lc3b_encode_frames_host(sh->enc, job.pcm_in + (size_t)sh->first * job.pcm_stride, job.pcm_stride,
                                         job.frames_out + (size_t)sh->first * job.frame_stride, job.nbytes, job.frame_stride, sh->stream);
        } else if (job.kind == 2) {
            if (sh->dec) rc = lc3b_decoder_host_fence(sh->dec, sh->stream);
            if (rc == LC3B_OK && cudaStreamSynchronize(sh->stream) != cudaSuccess) rc = cuda_fail(cudaGetLastError());
        }
        {
            std::lock_guard<std::mutex> lk(sh->mu);
            if (rc != LC3B_OK && sh->error == LC3B_OK) { sh->error = rc; sh->cuda_error = lc3b_last_cuda_error(); }
            sh->completed++;
        }
        sh->cv_done.notify_all();
        if (job.kind == 3) return;
    }
    (void)pool;
}

void submit(Shard* sh, const Job& job) {
    {
        std::lock_guard<std::mutex> lk(sh->mu);
        sh->queue.push_back(job);
        sh->submitted++;
    }
    sh->cv_job.notify_one();
}

int drain(Pool* pool) {
    Job f;
    f.kind = 2;
    for (Shard* sh : pool->shards) submit(sh, f);
    int rc = LC3B_OK;
    for (Shard* sh : pool->shards) {
        std::unique_lock<std::mutex> lk(sh->mu);
        sh->cv_done.wait(lk, [&] { return sh->completed == sh->submitted; });
        if (sh->error != LC3B_OK && rc == LC3B_OK) { rc = sh->error; if (rc == LC3B_ERR_CUDA) cuda_fail((cudaError_t)sh->cuda_error); }
        sh->error = LC3B_OK;
    }
    return rc;
}

void destroy(Pool* pool) {
    if (!pool) return;
    Job q;
    q.kind = 3;
    for (Shard* sh : pool->shards) {
        if (sh->thread.joinable()) {
            submit(sh, q);
            sh->thread.join();
        }
        DeviceGuard guard;
        cudaSetDevice(sh->device);
        if (sh->stream) cudaStreamSynchronize(sh->stream);
        if (sh->dec) lc3b_decoder_destroy(sh->dec);
        if (sh->enc) lc3b_encoder_destroy(sh->enc);
        if (sh->stream) cudaStreamDestroy(sh->stream);
        if (sh->workspace) cudaFree(sh->workspace);
        delete sh;
    }
    delete pool;
}

// stream s -> shard floor(s * G / N): shard g owns [ceil(g N / G), ceil((g + 1) N / G))
int shard_first(int g, int n_streams, int n_shards) { return (int)(((long long)g * n_streams + n_shards - 1) / n_shards); }

int create(Pool** out, bool encoder, int n_streams, int frame_duration, int sampling_frequency, int max_nbytes, const int* devices,
           int n_devices) {
    lc3b_config c;
    if (!out || n_streams <= 0 || n_devices <= 0 || n_devices > 64 || n_streams < n_devices || max_nbytes <= 0 ||
        max_nbytes > MAX_NBYTES || !config_new(sampling_frequency, frame_duration, &c))
        return LC3B_ERR_INVALID_ARG;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess) return cuda_fail(cudaGetLastError());
    Pool* pool = new (std::nothrow) Pool();
    if (!pool) return LC3B_ERR_INVALID_ARG;
    pool->is_encoder = encoder;
    pool->n_streams = n_streams;
    pool->nf = c.nf;
    pool->max_nbytes = max_nbytes;
    DeviceGuard guard;
    for (int g = 0; g < n_devices; g++) {
        Shard* sh = new (std::nothrow) Shard();
        if (!sh) { destroy(pool); return LC3B_ERR_INVALID_ARG; }
        pool->shards.push_back(sh);
        sh->device = devices ? devices[g] : g;
        sh->first = shard_first(g, n_streams, n_devices);
        sh->count = shard_first(g + 1, n_streams, n_devices) - sh->first;
        if (sh->device < 0 || sh->device >= visible) { destroy(pool); return LC3B_ERR_INVALID_ARG; }
        size_t bytes = 0;
        int rc = encoder ? lc3b_encoder_workspace_bytes(sh->count, frame_duration, sampling_frequency, max_nbytes, &bytes)
                         : lc3b_decoder_workspace_bytes(sh->count, frame_duration, sampling_frequency, max_nbytes, &bytes);
        if (rc != LC3B_OK) { destroy(pool); return rc; }
        cudaError_t e = cudaSetDevice(sh->device);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&sh->stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&sh->workspace, bytes);
        if (e != cudaSuccess) { destroy(pool); return cuda_fail(e); }
        rc = encoder ? lc3b_encoder_init(&sh->enc, sh->count, frame_duration, sampling_frequency, max_nbytes, sh->device, sh->workspace, bytes, sh->stream)
                     : lc3b_decoder_init(&sh->dec, sh->count, frame_duration, sampling_frequency, max_nbytes, sh->device, sh->workspace, bytes, sh->stream);
        if (rc == LC3B_OK) rc = encoder ? lc3b_encoder_set_host_pipelining(sh->enc, 1) : lc3b_decoder_set_host_pipelining(sh->dec, 1);
        if (rc != LC3B_OK) { destroy(pool); return rc; }
    }
    for (Shard* sh : pool->shards) sh->thread = std::thread(worker, pool, sh);
    *out = pool;
    return LC3B_OK;
}

int shard_info(const Pool* pool, int shard, int* device, int* first_stream, int* n_streams) {
    if (!pool || shard < 0 || shard >= (int)pool->shards.size()) return LC3B_ERR_INVALID_ARG;
    const Shard* sh = pool->shards[(size_t)shard];
    if (device) *device = sh->device;
    if (first_stream) *first_stream = sh->first;
    if (n_streams) *n_streams = sh->count;
    return LC3B_OK;
}

}  // namespace

struct lc3b_sharded_decoder { Pool* pool; };
struct lc3b_sharded_encoder { Pool* pool; };

extern "C" {

int lc3b_host_alloc(void** out, size_t bytes) {
    if (!out || bytes == 0) return LC3B_ERR_INVALID_ARG;
    LC3B_CU(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
    return LC3B_OK;
}

void lc3b_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int lc3b_sharded_decoder_create(lc3b_sharded_decoder** out, int n_streams, int frame_duration, int sampling_frequency,
                                int max_nbytes, const int* devices, int n_devices) {
    if (!out) return LC3B_ERR_INVALID_ARG;
    Pool* pool = nullptr;
    const int rc = create(&pool, false, n_streams, frame_duration, sampling_frequency, max_nbytes, devices, n_devices);
    if (rc != LC3B_OK) return rc;
    *out = new lc3b_sharded_decoder{pool};
    return LC3B_OK;
}

int lc3b_sharded_decoder_n_shards(const lc3b_sharded_decoder* h) { return h ? (int)h->pool->shards.size() : 0; }

int lc3b_sharded_decoder_shard(const lc3b_sharded_decoder* h, int shard, int* device, int* first_stream, int* n_streams) {
    return h ? shard_info(h->pool, shard, device, first_stream, n_streams) : LC3B_ERR_INVALID_ARG;
}

int lc3b_sharded_decode_frames_host(lc3b_sharded_decoder* h, int bits_per_sample, const uint8_t* frames,
                                    const int32_t* frame_nbytes, int nbytes, size_t frame_stride, int16_t* pcm_out,
                                    size_t pcm_stride, int32_t* status_out) {
    if (!h || !frames || !pcm_out) return LC3B_ERR_INVALID_ARG;
    if (bits_per_sample != 16) return LC3B_ERR_BITS_PER_SAMPLE;              // lc3_decoder.rs:80
    const Pool* pool = h->pool;
    if (nbytes < 0 || nbytes > pool->max_nbytes || (size_t)nbytes > frame_stride || pcm_stride < (size_t)pool->nf)
        return LC3B_ERR_INVALID_ARG;
    Job job;
    job.kind = 0;
    job.frames = frames;
    job.frame_nbytes = frame_nbytes;
    job.nbytes = nbytes;
    job.frame_stride = frame_stride;
    job.pcm_out = pcm_out;
    job.pcm_stride = pcm_stride;
    job.status_out = status_out;
    for (Shard* sh : h->pool->shards) submit(sh, job);
    return LC3B_OK;
}

int lc3b_sharded_decoder_wait(lc3b_sharded_decoder* h) { return h ? drain(h->pool) : LC3B_ERR_INVALID_ARG; }

int lc3b_sharded_decoder_set_min_nbytes(lc3b_sharded_decoder* h, int min_nbytes) {
    if (!h) return LC3B_ERR_INVALID_ARG;
    const int rc0 = drain(h->pool);                       // no call in flight while the shards' settings change
    if (rc0 != LC3B_OK) return rc0;
    for (Shard* sh : h->pool->shards) {
        const int rc = lc3b_decoder_set_min_nbytes(sh->dec, min_nbytes);
        if (rc != LC3B_OK) return rc;
    }
    return LC3B_OK;
}

void lc3b_sharded_decoder_destroy(lc3b_sharded_decoder* h) {
    if (!h) return;
    destroy(h->pool);
    delete h;
}

int lc3b_sharded_encoder_create(lc3b_sharded_encoder** out, int n_streams, int frame_duration, int sampling_frequency,
                                int max_nbytes, const int* devices, int n_devices) {
    if (!out) return LC3B_ERR_INVALID_ARG;
    Pool* pool = nullptr;
    const int rc = create(&pool, true, n_streams, frame_duration, sampling_frequency, max_nbytes, devices, n_devices);
    if (rc != LC3B_OK) return rc;
    *out = new lc3b_sharded_encoder{pool};
    return LC3B_OK;
}

int lc3b_sharded_encoder_n_shards(const lc3b_sharded_encoder* h) { return h ? (int)h->pool->shards.size() : 0; }

int lc3b_sharded_encoder_shard(const lc3b_sharded_encoder* h, int shard, int* device, int* first_stream, int* n_streams) {
    return h ? shard_info(h->pool, shard, device, first_stream, n_streams) : LC3B_ERR_INVALID_ARG;
}

int lc3b_sharded_encode_frames_host(lc3b_sharded_encoder* h, const int16_t* pcm_in, size_t pcm_stride, uint8_t* frames_out,
                                    int nbytes, size_t frame_stride) {
    if (!h || !pcm_in || !frames_out) return LC3B_ERR_INVALID_ARG;
    const Pool* pool = h->pool;
    if (nbytes < 20 || nbytes > pool->max_nbytes || (size_t)nbytes > frame_stride || pcm_stride < (size_t)pool->nf)
        return LC3B_ERR_INVALID_ARG;
    Job job;
    job.kind = 1;
    job.pcm_in = pcm_in;
    job.pcm_stride = pcm_stride;
    job.frames_out = frames_out;
    job.nbytes = nbytes;
    job.frame_stride = frame_stride;
    for (Shard* sh : h->pool->shards) submit(sh, job);
    return LC3B_OK;
}

int lc3b_sharded_encoder_wait(lc3b_sharded_encoder* h) { return h ? drain(h->pool) : LC3B_ERR_INVALID_ARG; }

void lc3b_sharded_encoder_destroy(lc3b_sharded_encoder* h) {
    if (!h) return;
    destroy(h->pool);
    delete h;
}

}  // extern "C"
