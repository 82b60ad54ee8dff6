// Column-major FP64 device matrix with a padded (even, 64-byte multiple) leading dimension.
#pragma once
#include "common.cuh"

namespace scrc {

struct DMat {
    DevBuf<double> buf;
    int64_t rows = 0, cols = 0, ld = 0;

    DMat() = default;
    DMat(int64_t r, int64_t c) { alloc(r, c); }
    void alloc(int64_t r, int64_t c) {
        rows = r; cols = c; ld = round_up(r > 0 ? r : 1, 8);
        buf.alloc((size_t)ld * (size_t)(c > 0 ? c : 1));
    }
    // (re)shape inside the existing allocation when it is big enough
    void reshape(int64_t r, int64_t c) {
        const int64_t nld = round_up(r > 0 ? r : 1, 8);
        const size_t need = (size_t)nld * (size_t)(c > 0 ? c : 1);
        if (need > buf.n) buf.alloc(need);
        rows = r; cols = c; ld = nld;
    }
    double* data() { return buf.p; }
    const double* data() const { return buf.p; }
    double* col(int64_t c) { return buf.p + c * ld; }
    const double* col(int64_t c) const { return buf.p + c * ld; }
    size_t bytes() const { return (size_t)ld * (size_t)cols * 8; }

    // Column-major source (host OR device pointer, leading dimension `hld`) -> device; padding rows are zeroed.
    void upload(const double* host, int64_t hld, cudaStream_t s) {
        if (rows == 0 || cols == 0) return;
        if (ld != rows) SCRC_CUDA(cudaMemsetAsync(buf.p, 0, bytes(), s));
        SCRC_CUDA(cudaMemcpy2DAsync(buf.p, ld * 8, host, hld * 8, rows * 8, cols, cudaMemcpyDefault, s));
    }
    void download(double* host, int64_t hld, cudaStream_t s) const {
        if (rows == 0 || cols == 0) return;
        SCRC_CUDA(cudaMemcpy2DAsync(host, hld * 8, buf.p, ld * 8, rows * 8, cols, cudaMemcpyDeviceToHost, s));
    }
};

}  // namespace scrc
