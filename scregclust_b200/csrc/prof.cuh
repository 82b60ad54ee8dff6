// Optional per-kernel-class timing with CUDA events on the launching stream (used by bench.py
// for the roofline figures).  Disabled by default: zero overhead beyond one branch per launch.
#pragma once
#include <vector>

#include "common.cuh"

namespace scrc {

enum ProfClass : int {
    PROF_GEMM_DENSE = 0,   // Gram / correlation / eigen-solve GEMMs (DensePolicy)
    PROF_GEMM_ADMM = 1,    // the two eigenbasis GEMMs of every ADMM sweep
    PROF_ADMM_ELEM = 2,    // rhs + relax/shrink/dual update + finalize kernels
    PROF_NNLS = 3,
    PROF_RESID_VOTE = 4,   // fused predict/residual/SSE/vote kernel
    PROF_OTHER = 5,
    PROF_NCLASS = 6
};

struct Profiler {
    bool enabled = false;
    struct Rec { int cls; cudaEvent_t e0, e1; };
    std::vector<Rec> pending;
    std::vector<cudaEvent_t> pool;
    double ms[PROF_NCLASS] = {0};
    double work[PROF_NCLASS] = {0};          // algorithmic flops (GEMM classes) or bytes
    unsigned long long launches[PROF_NCLASS] = {0};

    cudaEvent_t get_event();
    void begin(int cls, cudaStream_t s);
    void end(cudaStream_t s) noexcept;        // called from ~ProfScope: records the error instead of throwing
    cudaError_t deferred_error = cudaSuccess; // first failure seen by end(); rethrown by collect()
    void add_work(int cls, double w) { if (enabled) work[cls] += w; }
    void collect();                           // synchronises and folds pending records into ms[]
    void reset();
};
// One profiler per host thread: in thread-per-device mode every worker times its own stream.
extern thread_local Profiler g_prof;

struct ProfScope {
    cudaStream_t s;
    bool on;
    ProfScope(int cls, cudaStream_t st, double w = 0.0) : s(st), on(g_prof.enabled) {
        if (on) { g_prof.begin(cls, s); g_prof.work[cls] += w; }
    }
    ~ProfScope() { if (on) g_prof.end(s); }
    ProfScope(const ProfScope&) = delete;
    ProfScope& operator=(const ProfScope&) = delete;
};

}  // namespace scrc
