// Shared device/host helpers for libscregclust_b200 (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <stdexcept>

namespace scrc {

// ----------------------------------------------------------------------------
// Error plumbing: internal code throws, the extern "C" layer converts to codes.
// ----------------------------------------------------------------------------
struct Error : public std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define SCRC_CUDA(expr)                                                                   \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            throw ::scrc::Error(-100 - (int)_e, std::string("CUDA error ") +              \
                                cudaGetErrorName(_e) + " (" + cudaGetErrorString(_e) +    \
                                ") at " + __FILE__ + ":" + std::to_string(__LINE__));     \
        }                                                                                 \
    } while (0)

inline int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }
inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Count of kernels of this library launched by the calling thread's process
// (reported as `gpu_launches` by bench.py).
extern std::atomic<unsigned long long> g_launch_count;
#define SCRC_LAUNCHED() (::scrc::g_launch_count.fetch_add(1ull, std::memory_order_relaxed))

// ----------------------------------------------------------------------------
// Device buffer (RAII)
//
// Allocation is stream-ordered (cudaMallocAsync / cudaFreeAsync on the stream all of the calling session's work
// runs on) whenever the public entry point has opened an AllocScope: the per-cycle work buffers change size from
// cycle to cycle and from session to session, and synchronous cudaMalloc / cudaFree of multi-GB buffers cost
// ~0.5 s per penalty grid.  The device's default memory pool keeps freed blocks (release threshold = unlimited,
// Device::init), so steady-state allocation never reaches the driver.  Outside a scope (result handles, teardown)
// plain cudaMalloc / cudaFree are used; cudaFree is valid for stream-ordered allocations too.
// ----------------------------------------------------------------------------
extern thread_local cudaStream_t g_alloc_stream;
extern thread_local bool g_alloc_async;
struct AllocScope {
    cudaStream_t prev_s; bool prev_a;
    explicit AllocScope(cudaStream_t s) : prev_s(g_alloc_stream), prev_a(g_alloc_async) { g_alloc_stream = s; g_alloc_async = true; }
    ~AllocScope() { g_alloc_stream = prev_s; g_alloc_async = prev_a; }
    AllocScope(const AllocScope&) = delete;
    AllocScope& operator=(const AllocScope&) = delete;
};

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) { if (g_alloc_async) cudaFreeAsync(p, g_alloc_stream); else cudaFree(p); }
        p = nullptr; n = 0;
    }
    void alloc(size_t n_) {
        release();
        n = n_;
        if (!n) return;
        if (g_alloc_async) SCRC_CUDA(cudaMallocAsync((void**)&p, n * sizeof(T), g_alloc_stream));
        else SCRC_CUDA(cudaMalloc((void**)&p, n * sizeof(T)));
    }
    void ensure(size_t n_) { if (n_ > n) alloc(n_); }
    void zero(cudaStream_t s) { if (n) SCRC_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), s)); }
};

#ifdef __CUDACC__
// ----------------------------------------------------------------------------
// PTX wrappers: mbarrier, TMA (cp.async.bulk.tensor), FP64 mma (DMMA)
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    while (!mbar_try_wait(addr, parity)) {
    }
}
// Warp-collective wait used by the consumer warps: ONE lane polls the mbarrier, the rest of the
// warp joins at __syncwarp() (which also orders their subsequent shared-memory reads after the
// polling lane's acquire).  Do NOT let all 32 lanes spin on try_wait in front of warp-collective
// instructions: ptxas treats the try_wait predicate as warp-uniform (uniform operands) and emits
// no reconvergence, but on B200 lanes were observed to leave the spin loop on different
// iterations when the phase completes mid-poll; the early lanes then ran LDS + DMMA with a
// partial warp (rare, run-to-run varying ~1e-3 errors in one warp's tile).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
    if (lane == 0) mbar_wait(bar, parity);
    __syncwarp();
}
// Consumer release of a TMA stage.  The fragment loads are generic-proxy reads; the refill is an
// async-proxy (TMA) write.  Every lane fences its own reads against the async proxy, the warp
// joins, then one lane arrives on the `empty` barrier.  Without the cross-proxy fence the arrive
// (which ptxas is free to schedule right behind the last LDS *issue*) can be observed before
// queued LDS have read the stage, and the producer's next TMA load overwrites data still in use.
__device__ __forceinline__ void stage_release_warp(uint64_t* empty_bar, int lane) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar);
}
// 2-D tiled TMA load: coordinates are (c0 = contiguous dim, c1 = strided dim).
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, "
        "{%2, %3}], [%4];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// a / b with the bits of the IEEE division but ~5 FP64 instructions: rb = RN(1 / b) (computed once with a real
// division), then two Markstein quotient corrections q <- q + RN(a - b q) rb.  After the first one q is a
// faithful quotient, so the second returns RN(a / b).  Valid while the remainders are exact, i.e. for
// 2^-930 <= |a| < 2^930 (or a == 0) and b in the same range: callers test quot_in_range() and fall back to `/`.
__device__ __forceinline__ bool quot_in_range(double a) {
    const unsigned e = ((unsigned)__double2hiint(a) >> 20) & 0x7ffu;
    return (e - 93u) < 1860u;
}
__device__ __forceinline__ double quot_exact(double a, double b, double rb) {
    double q = a * rb;
    q = fma(fma(-q, b, a), rb, q);
    q = fma(fma(-q, b, a), rb, q);
    return q;
}

// D(8x8) += A(8x4, row) * B(4x8, col), all FP64 -> SASS DMMA.8x8x4
__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#endif  // __CUDACC__

// ----------------------------------------------------------------------------
// Host: TMA descriptor for a column-major FP64 matrix with `rows` contiguous.
// Tensor dims = (rows, cols); box = (16, box_cols); 128-byte swizzle.
// ----------------------------------------------------------------------------
// Opt-in to more than 48 KB of dynamic shared memory, remembered per kernel AND per device (the
// attribute belongs to the device's copy of the function; a process may drive several devices).
struct SmemAttr {
    int sizes[64] = {};
    template <class Kernel>
    void ensure(Kernel kernel, size_t smem) {
        int dev = 0;
        SCRC_CUDA(cudaGetDevice(&dev));
        dev &= 63;
        if ((int)smem > sizes[dev]) {
            SCRC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sizes[dev] = (int)smem;
        }
    }
};

CUtensorMap make_tmap_f64(const double* base, int64_t rows, int64_t cols, int64_t ld,
                          int box_cols);
// Same matrix view, unswizzled box of (box_rows x box_cols): shared-memory image is dense
// column-major [box_cols][box_rows] (used to stage slabs of the ADMM state).
CUtensorMap make_tmap_f64_plain(const double* base, int64_t rows, int64_t cols, int64_t ld,
                                int box_rows, int box_cols);

}  // namespace scrc
