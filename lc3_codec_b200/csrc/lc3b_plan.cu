// lc3b engine: executors for a LaunchPlan (lc3b_plan.cuh) - direct stream launches or a cached CUDA graph.
#include "lc3b_plan.cuh"

namespace lc3b {

cudaError_t PlanLanes::ensure() {
    if (fork) return cudaSuccess;
    cudaError_t e = cudaEventCreateWithFlags(&fork, cudaEventDisableTiming);
    for (int i = 0; i < PLAN_MAX_LANES - 1 && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithFlags(&aux[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&join[i], cudaEventDisableTiming);
    }
    return e;
}

PlanLanes::~PlanLanes() {
    for (int i = 0; i < PLAN_MAX_LANES - 1; i++) {
        if (aux[i]) { cudaStreamSynchronize(aux[i]); cudaStreamDestroy(aux[i]); }
        if (join[i]) cudaEventDestroy(join[i]);
    }
    if (fork) cudaEventDestroy(fork);
}

cudaError_t plan_launch_direct(const LaunchPlan& plan, cudaStream_t stream, PlanLanes* lanes) {
    int max_lane = 0;
    for (int i = 0; i < plan.n; i++) max_lane = plan.nodes[i].lane > max_lane ? plan.nodes[i].lane : max_lane;
    if (!lanes || max_lane >= PLAN_MAX_LANES) max_lane = 0;           // no auxiliary streams: everything in order on one
    cudaError_t e = cudaSuccess;
    if (max_lane > 0) {
        e = lanes->ensure();
        if (e == cudaSuccess) e = cudaEventRecord(lanes->fork, stream);
        for (int l = 1; l <= max_lane && e == cudaSuccess; l++) e = cudaStreamWaitEvent(lanes->aux[l - 1], lanes->fork, 0);
        if (e != cudaSuccess) return e;
    }
    for (int i = 0; i < plan.n; i++) {
        const PlanNode& k = plan.nodes[i];
        void* args[1] = {(void*)k.param};
        cudaStream_t s = (max_lane > 0 && k.lane > 0) ? lanes->aux[k.lane - 1] : stream;
        e = cudaLaunchKernel(k.func, dim3(k.grid), dim3(k.block), args, k.smem, s);
        if (e != cudaSuccess) return e;
    }
    for (int l = 1; l <= max_lane; l++) {
        e = cudaEventRecord(lanes->join[l - 1], lanes->aux[l - 1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, lanes->join[l - 1], 0);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h * 0xBF58476D1CE4E5B9ull;
}

// shape = what cannot be changed on an instantiated graph (kernels, dependencies, count); args = everything else
static void plan_keys(const LaunchPlan& plan, uint64_t& shape, uint64_t& args) {
    uint64_t s = 0x4C433362ull + (uint64_t)plan.n, a = 1;
    for (int i = 0; i < plan.n; i++) {
        const PlanNode& k = plan.nodes[i];
        s = mix64(s, (uint64_t)(uintptr_t)k.func);
        s = mix64(s, ((uint64_t)(uint32_t)k.dep[0] << 32) | (uint32_t)k.dep[1]);
        s = mix64(s, (uint64_t)k.param_bytes);
        a = mix64(a, ((uint64_t)k.grid << 32) | k.block);
        a = mix64(a, k.smem);
        const uint64_t* w = (const uint64_t*)k.param;
        for (int j = 0; j < (k.param_bytes + 7) / 8; j++) a = mix64(a, w[j]);
    }
    shape = s;
    args = a;
}

static bool same_args(const LaunchPlan& x, const LaunchPlan& y) {
    if (x.n != y.n) return false;
    for (int i = 0; i < x.n; i++) {
        const PlanNode &a = x.nodes[i], &b = y.nodes[i];
        if (a.func != b.func || a.grid != b.grid || a.block != b.block || a.smem != b.smem || a.param_bytes != b.param_bytes ||
            memcmp(a.param, b.param, (size_t)a.param_bytes) != 0)
            return false;
    }
    return true;
}

static void node_params(const PlanNode& k, void** slot, cudaKernelNodeParams& kp) {
    *slot = (void*)k.param;
    memset(&kp, 0, sizeof(kp));
    kp.func = (void*)k.func;
    kp.gridDim = dim3(k.grid);
    kp.blockDim = dim3(k.block);
    kp.sharedMemBytes = k.smem;
    kp.kernelParams = slot;
    kp.extra = nullptr;
}

static cudaError_t build_entry(GraphCacheEntry& en, const LaunchPlan& plan) {
    en.plan = plan;
    en.gnodes.assign((size_t)plan.n, nullptr);
    cudaError_t e = cudaGraphCreate(&en.graph, 0);
    if (e != cudaSuccess) return e;
    for (int i = 0; i < plan.n; i++) {
        const PlanNode& k = en.plan.nodes[i];
        cudaGraphNode_t deps[2];
        int nd = 0;
        for (int d = 0; d < 2; d++)
            if (k.dep[d] >= 0 && k.dep[d] < i) deps[nd++] = en.gnodes[(size_t)k.dep[d]];
        void* slot;
        cudaKernelNodeParams kp;
        node_params(k, &slot, kp);
        e = cudaGraphAddKernelNode(&en.gnodes[(size_t)i], en.graph, deps, (size_t)nd, &kp);
        if (e != cudaSuccess) return e;
    }
    return cudaGraphInstantiate(&en.exec, en.graph, 0);
}

static void destroy_entry(GraphCacheEntry* en) {
    if (en->exec) cudaGraphExecDestroy(en->exec);
    if (en->graph) cudaGraphDestroy(en->graph);
    delete en;
}

GraphCache::~GraphCache() {
    for (GraphCacheEntry* en : entries) destroy_entry(en);
}

cudaError_t plan_launch_graph(GraphCache& cache, const LaunchPlan& plan, cudaStream_t stream) {
    if (plan.n == 0) return cudaSuccess;
    uint64_t shape, args;
    plan_keys(plan, shape, args);
    cache.clock++;
    GraphCacheEntry* lru_same_shape = nullptr;
    GraphCacheEntry* lru_any = nullptr;
    for (GraphCacheEntry* en : cache.entries) {
        if (en->shape == shape && en->args == args && same_args(en->plan, plan)) {
            en->last_use = cache.clock;
            cache.hits++;
            return cudaGraphLaunch(en->exec, stream);
        }
        if (en->shape == shape && (!lru_same_shape || en->last_use < lru_same_shape->last_use)) lru_same_shape = en;
        if (!lru_any || en->last_use < lru_any->last_use) lru_any = en;
    }
    if (cache.entries.size() >= GraphCache::MAX_ENTRIES && lru_same_shape) {
        // same kernels and dependencies, other pointers / sizes: patch the executable graph in place
        GraphCacheEntry& en = *lru_same_shape;
        for (int i = 0; i < plan.n; i++) {
            const PlanNode &a = en.plan.nodes[i], &b = plan.nodes[i];
            if (a.grid == b.grid && a.block == b.block && a.smem == b.smem && memcmp(a.param, b.param, (size_t)a.param_bytes) == 0) continue;
            en.plan.nodes[i] = b;
            void* slot;
            cudaKernelNodeParams kp;
            node_params(en.plan.nodes[i], &slot, kp);
            cudaError_t e = cudaGraphExecKernelNodeSetParams(en.exec, en.gnodes[(size_t)i], &kp);
            if (e != cudaSuccess) return e;
        }
        en.args = args;
        en.last_use = cache.clock;
        cache.updates++;
        return cudaGraphLaunch(en.exec, stream);
    }
    if (cache.entries.size() >= GraphCache::MAX_ENTRIES) {
        for (size_t i = 0; i < cache.entries.size(); i++)
            if (cache.entries[i] == lru_any) { cache.entries.erase(cache.entries.begin() + (long)i); break; }
        destroy_entry(lru_any);
    }
    GraphCacheEntry* en = new GraphCacheEntry();
    en->shape = shape;
    en->args = args;
    en->last_use = cache.clock;
    cudaError_t e = build_entry(*en, plan);
    if (e != cudaSuccess) { destroy_entry(en); return e; }
    cache.entries.push_back(en);
    cache.builds++;
    return cudaGraphLaunch(en->exec, stream);
}

}  // namespace lc3b
