// lc3b engine: one call's kernel launches as data.
//
// Every decode / encode entry point first writes down WHAT it will launch (kernel, geometry, the one parameter struct
// each kernel takes, which earlier launches it depends on) and then hands that plan to one of two executors:
//   * plan_launch_direct - the launches in order on the caller's stream (what round 1 did); a plan whose nodes carry
//     lanes (independent sub-batches of one call) is issued with lane l > 0 on the l-th auxiliary stream of a PlanLanes,
//     forked from and joined back into the caller's stream with events, so that the sub-batches' kernels overlap;
//   * plan_launch_graph  - the same launches as a CUDA graph.  Independent nodes (the per-configuration synthesis kernels
//     of a mixed-rate batch) run concurrently without any fork/join streams, and a whole call costs one launch on the
//     host, which is what small batches (BASELINE configs 2-4: 8 192 .. 65 536 streams) are sensitive to.
// Graphs are kept in a small per-handle cache keyed by the plan's bytes: a caller that cycles through a few sets of
// buffers (the usual ring) replays instantiated graphs; a plan with the same topology but new pointers updates the
// least recently used executable graph in place (cudaGraphExecKernelNodeSetParams, microseconds) instead of
// re-instantiating it.  Nothing here allocates device memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace lc3b {

constexpr int PLAN_PARAM_MAX = 232;      // bytes; every kernel takes ONE parameter struct by value
constexpr int PLAN_MAX_NODES = 48;

struct PlanNode {
    const void* func;
    unsigned grid, block, smem;
    int dep[2];                          // indices of earlier nodes this one waits for, -1 = none; no deps = a root
    int lane;                            // direct launches: which stream (0 = the caller's); dependencies stay inside a lane
    int param_bytes;
    alignas(16) unsigned char param[PLAN_PARAM_MAX];
};

constexpr int PLAN_MAX_LANES = 4;

struct LaunchPlan {
    int n = 0;
    int lane = 0;                        // lane of the nodes added from now on
    PlanNode nodes[PLAN_MAX_NODES];
    LaunchPlan() { memset(nodes, 0, sizeof(nodes)); }   // padding bytes are part of the cache key

    // Adds kernel<<<grid, block, smem>>>(p); returns the node's index.  dep0 = -2 means "the node added just before".
    template <class P>
    int add(void (*kernel)(P), unsigned grid, unsigned block, size_t smem, const P& p, int dep0 = -2, int dep1 = -1) {
        static_assert(sizeof(P) <= PLAN_PARAM_MAX, "kernel parameter struct too large for a plan node");
        if (n >= PLAN_MAX_NODES || grid == 0) return n - 1;
        PlanNode& k = nodes[n];
        k.func = (const void*)kernel;
        k.grid = grid; k.block = block; k.smem = (unsigned)smem;
        k.dep[0] = dep0 == -2 ? n - 1 : dep0;
        k.dep[1] = dep1;
        k.lane = lane;
        k.param_bytes = (int)sizeof(P);
        memcpy(k.param, &p, sizeof(P));
        return n++;
    }
};

// Auxiliary streams and events for plans with lanes; owned by a handle, created on first use on the current device.
struct PlanLanes {
    cudaStream_t aux[PLAN_MAX_LANES - 1] = {nullptr, nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[PLAN_MAX_LANES - 1] = {nullptr, nullptr, nullptr};
    cudaError_t ensure();
    ~PlanLanes();
};

cudaError_t plan_launch_direct(const LaunchPlan& plan, cudaStream_t stream, PlanLanes* lanes = nullptr);

struct GraphCacheEntry {
    uint64_t shape = 0, args = 0, last_use = 0;
    cudaGraphExec_t exec = nullptr;
    cudaGraph_t graph = nullptr;
    std::vector<cudaGraphNode_t> gnodes;
    LaunchPlan plan;
};

struct GraphCache {
    static constexpr size_t MAX_ENTRIES = 24;
    std::vector<GraphCacheEntry*> entries;
    uint64_t clock = 0;
    uint64_t hits = 0, updates = 0, builds = 0;   // statistics (lc3b_*_graph_stats)
    ~GraphCache();
};

cudaError_t plan_launch_graph(GraphCache& cache, const LaunchPlan& plan, cudaStream_t stream);

}  // namespace lc3b
