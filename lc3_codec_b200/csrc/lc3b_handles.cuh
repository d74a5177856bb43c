// lc3b engine: the opaque handle types of the C ABI (include/lc3b.h), shared by the translation units that implement it.
#pragma once
#include "lc3b_common.cuh"
#include "lc3b_plan.cuh"

struct lc3b_decoder {
    lc3b::DecoderState st{};
    int stage_mask = 7;
    // how a call's kernels are issued: 0 = one launch each on the caller's stream, 1 = one cached CUDA graph per call
    int graph_mode = 0;
    lc3b::GraphCache graphs;
    // a large batch is issued as up to four independent sub-batches on auxiliary streams (run_decode in lc3b_api.cu)
    int split = 0;                        // 0 = by batch size, 1 = never, 2 / 4 = that many sub-batches
    lc3b::PlanLanes lanes;
    // optional pipelining of the host entry point: PCM leaves on an internal copy stream from a double-buffered
    // staging area, so the device->host copy of call i overlaps the kernels of call i+1
    int pipelined = 0;
    int buf = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t compute_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    bool d2h_pending[2] = {false, false};
};

namespace lc3b {

// one thread-local slot for the last CUDA error of any entry point (decoder, encoder, mixed, sharded)
int cuda_fail(cudaError_t e);
#define LC3B_CU(x)                                         \
    do {                                                   \
        cudaError_t _e = (x);                              \
        if (_e != cudaSuccess) return lc3b::cuda_fail(_e); \
    } while (0)

// graph mode default: LC3B_GRAPH=0/1 in the environment overrides; otherwise graphs are used for batches small enough
// to be launch-sensitive
int default_graph_mode(int n_streams);

// internals of lc3b_api.cu used by the mixed-rate and sharded front ends
bool config_new(int sampling_frequency, int frame_duration, lc3b_config* c);
int decoder_workspace_bytes(int n_streams, int frame_duration, int sampling_frequency, int max_nbytes, bool staging,
                            size_t* device_bytes);
int decoder_init(lc3b_decoder** out, int n_streams, int frame_duration, int sampling_frequency, int max_nbytes, int device,
                 void* dev_workspace, size_t workspace_bytes, void* cuda_stream, bool staging);

// Restores the caller's current CUDA device when it goes out of scope (the init functions select the handle's device).
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace lc3b
