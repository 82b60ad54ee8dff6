// Small data-movement / statistics kernels shared by the session and the drop-ins.
#pragma once
#include "dmat.cuh"

namespace scrc {

// out (cols x rows) = in' ; both column-major.
void transpose(const double* in, int64_t ldi, int64_t rows, int64_t cols, double* out, int64_t ldo,
               cudaStream_t s);

// Column-centre `x` in place (mean removed) and return the centred sums of squares
// (R/utils.R:534-537).  One CTA per column, deterministic reductions.
void col_center_ss(double* x, int64_t ld, int64_t rows, int64_t cols, double* ss, cudaStream_t s);

// css[c] = sum_r (x[r,c] - mean_c)^2 without modifying x (R: colSums(scale(x, center=TRUE, scale=FALSE)^2))
void col_centered_ss(const double* x, int64_t ld, int64_t rows, int64_t cols, double* css, cudaStream_t s);

// ss[c] = sum_r x[r,c]^2
void col_sumsq(const double* x, int64_t ld, int64_t rows, int64_t cols, double* ss, cudaStream_t s);

// Symmetric eigen-decomposition A = V diag(w) V' (A: n x n, lower triangle read;
// overwritten by V).  One-time setup per session / per drop-in call.
void sym_eig(double* A, int64_t lda, int n, double* w, cudaStream_t s);

}  // namespace scrc
