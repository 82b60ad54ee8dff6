// Probe: do the FP64 tensor (DMMA, mma.sync.m8n8k4.f64) and the FP64 FMA pipe of sm_100a run concurrently?
// Four launches of the same kernel: DMMA warps only, DFMA warps only, both kinds side by side on every SM
// sub-partition, and all warps DMMA.  If "both" takes max(DMMA, DFMA) the pipes are independent and a hybrid
// GEMM could exceed the DMMA rate; if it takes the sum they share the datapath.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/probe/fp64_pipes tools/probe/fp64_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode bit 0: DMMA warps work, bit 1: DFMA warps work, mode 4: every warp DMMA
__global__ void __launch_bounds__(512, 1) probe(int mode, int iters, double* out, double seed) {
    const int warp = threadIdx.x >> 5;
    const bool tensor_warp = (mode == 4) || (((warp >> 2) & 1) == 0);
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seed * (i + 1);
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 - seed;
    if (tensor_warp) {
        if (!(mode & 1) && mode != 4) return;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma(acc[i], acc[i + 1], a, b);       // 8 independent DMMA per iteration
        }
    } else {
        if (!(mode & 2)) return;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);           // 64 DFMA per iteration, 16 chains
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    if (s == 123.456) out[0] = s;
}

int main() {
    double* out; cudaMalloc(&out, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, grid = 148;
    const char* names[] = {"", "dmma warps only (2 per sub-partition)", "dfma warps only (2 per sub-partition)",
                           "dmma + dfma warps together", "all 16 warps dmma"};
    for (int rep = 0; rep < 2; ++rep)
        for (int mode = 1; mode <= 4; ++mode) {
            cudaEventRecord(e0);
            probe<<<grid, 512>>>(mode, iters, out, 0.5);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            // flop: one DMMA = 8*8*4*2 = 512 flop per warp; one DFMA = 2 flop per thread
            double tw = (mode == 4) ? 16 : ((mode & 1) ? 8 : 0), fw = (mode & 2) && mode != 4 ? 8 : 0;
            double flop_t = tw * grid * (double)iters * 8 * 512, flop_f = fw * grid * (double)iters * 64 * 2 * 32;
            if (rep) printf("mode %d %-40s %8.3f ms  dmma %6.2f TF/s  dfma %6.2f TF/s  sum %6.2f\n", mode, names[mode], ms,
                            flop_t / ms * 1e-9, flop_f / ms * 1e-9, (flop_t + flop_f) / ms * 1e-9);
        }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(err));
    return err != cudaSuccess;
}
