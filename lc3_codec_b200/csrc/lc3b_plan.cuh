// lc3b engine: one call's kernel launches as data.
//
// Every decode / encode entry point first writes down WHAT it will launch (kernel, geometry, the one parameter struct
// each kernel takes, which earlier launches it depends on) and then hands that plan to one of two executors:
//   * plan_launch_direct - the launches in order on the caller's stream (what round 1 did);
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
    int param_bytes;
    alignas(16) unsigned char param[PLAN_PARAM_MAX];
};

struct LaunchPlan {
    int n = 0;
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
        k.param_bytes = (int)sizeof(P);
        memcpy(k.param, &p, sizeof(P));
        return n++;
    }
};

cudaError_t plan_launch_direct(const LaunchPlan& plan, cudaStream_t stream);

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
