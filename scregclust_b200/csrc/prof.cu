#include "prof.cuh"

namespace scrc {

thread_local Profiler g_prof;

cudaEvent_t Profiler::get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    SCRC_CUDA(cudaEventCreate(&e));
    return e;
}
void Profiler::begin(int cls, cudaStream_t s) {
    Rec r{cls, get_event(), get_event()};
    SCRC_CUDA(cudaEventRecord(r.e0, s));
    // collect() blocks the host until the stream has drained, which stalls the launch pipeline of whatever phase it
    // lands in (100+ ms seen inside session creation): only as a safety valve, far above what a profiled pass records
    if (pending.size() >= (1u << 20)) collect();
    pending.push_back(r);
    launches[cls] += 1;
}
void Profiler::end(cudaStream_t s) noexcept {
    // scopes are never nested: the record opened by the matching begin() is the last one
    if (pending.empty()) return;
    const cudaError_t e = cudaEventRecord(pending.back().e1, s);
    if (e != cudaSuccess && deferred_error == cudaSuccess) deferred_error = e;
}
void Profiler::collect() {
    if (deferred_error != cudaSuccess) {
        const cudaError_t e = deferred_error;
        deferred_error = cudaSuccess;
        pending.clear();
        SCRC_CUDA(e);
    }
    for (Rec& r : pending) {
        SCRC_CUDA(cudaEventSynchronize(r.e1));
        float t = 0.f;
        SCRC_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
        ms[r.cls] += t;
        pool.push_back(r.e0); pool.push_back(r.e1);
    }
    pending.clear();
}
void Profiler::reset() {
    collect();
    for (int i = 0; i < PROF_NCLASS; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
}

}  // namespace scrc
