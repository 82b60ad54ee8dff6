#include "misc.cuh"

#include <cusolverDn.h>

namespace scrc {

__global__ void transpose_kernel(const double* __restrict__ in, int64_t ldi, int64_t rows, int64_t cols,
                                 double* __restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = r0 + threadIdx.x, c = c0 + j;
        tile[j][threadIdx.x] = (r < rows && c < cols) ? in[r + c * ldi] : 0.0;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = c0 + threadIdx.x, r = r0 + j;   // out[c, r] = in[r, c]
        if (r < rows && c < cols) out[c + r * ldo] = tile[threadIdx.x][j];
    }
}

void transpose(const double* in, int64_t ldi, int64_t rows, int64_t cols, double* out, int64_t ldo,
               cudaStream_t s) {
    if (rows == 0 || cols == 0) return;
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(cols, 32));
    transpose_kernel<<<grid, dim3(32, 8), 0, s>>>(in, ldi, rows, cols, out, ldo);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// Deterministic block sum (fixed shuffle tree + fixed-order cross-warp pass).
__device__ double block_sum_256(double v, double* s_buf) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_buf[w];
    return t;
}

__global__ void __launch_bounds__(256) col_center_ss_kernel(double* x, int64_t ld, int64_t rows, double* ss) {
    __shared__ double s_buf[8];
    double* col = x + (int64_t)blockIdx.x * ld;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) a += col[i];
    const double mean = block_sum_256(a, s_buf) / (double)rows;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) {
        const double v = col[i] - mean;
        col[i] = v;
        q += v * v;
    }
    q = block_sum_256(q, s_buf);
    if (threadIdx.x == 0) ss[blockIdx.x] = q;
}

void col_center_ss(double* x, int64_t ld, int64_t rows, int64_t cols, double* ss, cudaStream_t s) {
    if (cols == 0) return;
    col_center_ss_kernel<<<(unsigned)cols, 256, 0, s>>>(x, ld, rows, ss);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

__global__ void __launch_bounds__(256) col_centered_ss_kernel(const double* x, int64_t ld, int64_t rows, double* css) {
    __shared__ double s_buf[8];
    const double* col = x + (int64_t)blockIdx.x * ld;
    double a = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) a += col[i];
    const double mean = block_sum_256(a, s_buf) / (double)rows;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) { const double v = col[i] - mean; q += v * v; }
    q = block_sum_256(q, s_buf);
    if (threadIdx.x == 0) css[blockIdx.x] = q;
}
void col_centered_ss(const double* x, int64_t ld, int64_t rows, int64_t cols, double* css, cudaStream_t s) {
    if (cols == 0) return;
    col_centered_ss_kernel<<<(unsigned)cols, 256, 0, s>>>(x, ld, rows, css);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

__global__ void __launch_bounds__(256) col_sumsq_kernel(const double* x, int64_t ld, int64_t rows, double* ss) {
    __shared__ double s_buf[8];
    const double* col = x + (int64_t)blockIdx.x * ld;
    double q = 0.0;
    for (int64_t i = threadIdx.x; i < rows; i += 256) q += col[i] * col[i];
    q = block_sum_256(q, s_buf);
    if (threadIdx.x == 0) ss[blockIdx.x] = q;
}

void col_sumsq(const double* x, int64_t ld, int64_t rows, int64_t cols, double* ss, cudaStream_t s) {
    if (cols == 0) return;
    col_sumsq_kernel<<<(unsigned)cols, 256, 0, s>>>(x, ld, rows, ss);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// One cuSOLVER handle per (host thread, device), created on first use and kept: creating one costs several
// milliseconds (it brings up a cuBLAS handle with its workspace) and a session is created per scregclust() call.
static cusolverDnHandle_t solver_handle() {
    struct Cache { cusolverDnHandle_t h[64] = {}; };   // never destroyed: library teardown order at exit is not ours
    static thread_local Cache cache;
    int dev = 0;
    SCRC_CUDA(cudaGetDevice(&dev));
    dev &= 63;
    if (!cache.h[dev] && cusolverDnCreate(&cache.h[dev]) != CUSOLVER_STATUS_SUCCESS) {
        cache.h[dev] = nullptr;
        throw Error(-80, "cusolverDnCreate failed");
    }
    return cache.h[dev];
}

void sym_eig(double* A, int64_t lda, int n, double* w, cudaStream_t s) {
    cusolverDnHandle_t h = solver_handle();
    cusolverDnSetStream(h, s);
    int lwork = 0;
    if (cusolverDnDsyevd_bufferSize(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, w,
                                    &lwork) != CUSOLVER_STATUS_SUCCESS)
        throw Error(-81, "cusolverDnDsyevd_bufferSize failed");
    DevBuf<double> work((size_t)lwork);
    DevBuf<int> info(1);
    if (cusolverDnDsyevd(h, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, (int)lda, w, work.p,
                         lwork, info.p) != CUSOLVER_STATUS_SUCCESS)
        throw Error(-82, "cusolverDnDsyevd failed");
    int hinfo = 0;
    SCRC_CUDA(cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    SCRC_CUDA(cudaStreamSynchronize(s));
    if (hinfo != 0) throw Error(-83, "symmetric eigen-decomposition did not converge (info=" + std::to_string(hinfo) + ")");
}

}  // namespace scrc
