// Batched NNLS kernels.  Per-column operation order follows src/optim.cpp:403-549
// and greedy_coord_descent (optim.cpp:335-377); see SURVEY.md section 3.3.
#include "nnls.cuh"
#include "prof.cuh"

#include <algorithm>
#include <vector>

namespace scrc {

struct NnlsDev {
    const NnlsProblem* prob;
    const int* sel;
    const int* pos;
    const int* cols;
    const double* sign;
    double* Q;
    double* dvec;
    double* grad0; int64_t ldg;
    double* beta; int64_t ldb;
    int* iters;
    int m;
    int col0, col1;          // column slice [col0, col1) processed by this launch
    const int* ids;          // problems processed by this launch (nullptr: all, in order)
};

static NnlsDev make_dev(NnlsBatch& b) {
    NnlsDev d{b.prob.p, b.sel.p, b.pos.p, b.cols.p, b.sign.p, b.Q.p, b.dvec.p, b.grad0.data(), b.grad0.ld,
              b.beta.data(), b.beta.ld, b.iters.p, b.m, b.col0, b.col1 < 0 ? b.m : b.col1, nullptr};
    return d;
}

// ---------------------------------------------------------------- setup: Q = (S G S) o (d d'),  d = 1/sqrt(diag)
__global__ void __launch_bounds__(256) nnls_setup_q_kernel(NnlsDev d, const double* __restrict__ xtx, int64_t ldx) {
    const NnlsProblem pr = d.prob[d.ids ? d.ids[blockIdx.x] : blockIdx.x];
    const int s = pr.s;
    const int* sel = d.sel + pr.sel_off;
    const double* sg = d.sign + pr.sel_off;
    double* Q = d.Q + pr.q_off;
    for (int a = threadIdx.x; a < s; a += blockDim.x) {
        const int ia = sel[a];
        d.dvec[pr.row_off + a] = 1.0 / sqrt(xtx[ia + (int64_t)ia * ldx]);   // optim.cpp:416
    }
    __syncthreads();
    for (int e = threadIdx.x; e < s * s; e += blockDim.x) {
        const int a = e % s, b = e / s;
        const int ia = sel[a], ib = sel[b];
        const double g = (sg[a] * sg[b]) * xtx[ia + (int64_t)ib * ldx];
        Q[e] = g * (d.dvec[pr.row_off + a] * d.dvec[pr.row_off + b]);      // optim.cpp:418-419
    }
}

// grad0 = (-(S X'Y colscale)) * d   (optim.cpp:421)
__global__ void __launch_bounds__(256) nnls_setup_grad_kernel(NnlsDev d, const double* __restrict__ xty,
                                                              int64_t ldy, const double* __restrict__ colscale) {
    // 1-D grid (problem-major): gridDim.y is capped at 65535, the importance step can exceed that
    const int chunks = (d.col1 - d.col0 + 31) / 32;
    const int pi = blockIdx.x / chunks;
    const NnlsProblem pr = d.prob[d.ids ? d.ids[pi] : pi];
    const int s = pr.s;
    const int* sel = d.sel + pr.sel_off;
    const double* sg = d.sign + pr.sel_off;
    const int col0 = d.col0 + (blockIdx.x % chunks) * 32;
    const int ncols = min(pr.ncols < 0 ? d.m : pr.ncols, d.col1);
    for (int e = threadIdx.x; e < s * 32; e += blockDim.x) {
        const int a = e % s, col = col0 + e / s;
        if (col >= ncols) break;
        const int xc = pr.xcol_off < 0 ? col : d.cols[pr.xcol_off + col];
        const double cs = colscale ? colscale[xc] : 1.0;
        const double v = sg[a] * (xty[sel[a] + (int64_t)xc * ldy] * cs);
        d.grad0[pr.row_off + a + (int64_t)col * d.ldg] = (-v) * d.dvec[pr.row_off + a];
    }
}

void nnls_setup(NnlsBatch& b, const double* xtx, int64_t ldx, const double* xty, int64_t ldy,
                const double* colscale, cudaStream_t st) {
    if (b.nprob == 0) return;
    if ((int)b.h_prob.size() != b.nprob) throw Error(-73, "nnls_setup: host problem descriptors missing");
    NnlsDev d = make_dev(b);
    if (d.col1 <= d.col0) return;
    // problems handled by this rank / launch: b.active (ids into prob), or all of them
    int nact = b.nprob;
    if (b.use_active) {
        nact = (int)b.active.size();
        if (nact == 0) return;
        b.ids.ensure(2 * (size_t)b.nprob + 2);
        SCRC_CUDA(cudaMemcpyAsync(b.ids.p + b.nprob, b.active.data(), (size_t)nact * 4, cudaMemcpyHostToDevice, st));
        d.ids = b.ids.p + b.nprob;
    }
    nnls_setup_q_kernel<<<nact, 256, 0, st>>>(d, xtx, ldx);
    SCRC_LAUNCHED();
    const int64_t gblocks = ceil_div(d.col1 - d.col0, 32) * (int64_t)nact;
    if (gblocks > 0x7fffffffLL) throw Error(-72, "NNLS: too many (problem, column chunk) pairs in one batch");
    nnls_setup_grad_kernel<<<(unsigned)gblocks, 256, 0, st>>>(d, xty, ldy, colscale);
    SCRC_LAUNCHED();
    SCRC_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- warp-level helpers
__device__ __forceinline__ double wdot(const double* u, const double* v, int s, int lane) {
    double a = 0.0;
    for (int i = lane; i < s; i += 32) a += u[i] * v[i];
    return warp_sum(a);
}

__device__ __forceinline__ void wmatvec(const double* __restrict__ Q, const double* v, double* out, int s,
                                        int lane) {
    for (int a = lane; a < s; a += 32) {
        double acc = 0.0;
        for (int b = 0; b < s; ++b) acc += Q[a + (int64_t)b * s] * v[b];
        out[a] = acc;
    }
}

// optim.cpp:451-457 / :475-481 -- ascending i, remove negative entries one at a time
__device__ __forceinline__ void wnegcorr(const double* __restrict__ Q, double* beta, double* grad, int s,
                                         int lane) {
    for (int base = 0; base < s; base += 32) {
        const int a = base + lane;
        unsigned mask = __ballot_sync(0xffffffffu, (a < s) && (beta[a] < 0.0));
        while (mask) {
            const int i = base + __ffs(mask) - 1;
            const double bi = beta[i];
            for (int r = lane; r < s; r += 32) grad[r] -= bi * Q[r + (int64_t)i * s];
            __syncwarp();
            if (lane == 0) beta[i] = 0.0;
            __syncwarp();
            mask &= mask - 1;
        }
    }
}

// optim.cpp:335-377 restricted to one column (columns are independent)
__device__ __forceinline__ void wgcd(const double* __restrict__ Q, double* beta, double* grad, int s,
                                     int lane) {
    for (int t = 0; t < s; ++t) {
        double best = -1.0;
        int bi = 0x7fffffff;
        for (int a = lane; a < s; a += 32) {
            const double g = grad[a];
            const double sc = ((beta[a] > 0.0) || (g < 0.0)) ? fabs(g) : 0.0;
            if (sc > best) { best = sc; bi = a; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (!(best > 0.0)) break;   // max_val == 0: empty passive set (optim.cpp:363-366)
        const int p = bi;
        const double bp = beta[p];
        const double dbeta = fmax(0.0, bp - grad[p] / Q[p + (int64_t)p * s]) - bp;
        __syncwarp();
        if (lane == 0) beta[p] = bp + dbeta;
        for (int a = lane; a < s; a += 32) grad[a] += dbeta * Q[a + (int64_t)p * s];
        __syncwarp();
    }
}

__device__ __forceinline__ bool alpha_ok(double a) {
    return (a == a) && fabs(a) >= 1e-20 && fabs(a) < 1e30;   // optim.cpp:448,472
}

template <bool QSMEM>
__global__ void nnls_kernel(NnlsDev d, const int* __restrict__ prob_ids, int cols_per_cta, int chunks, double eps,
                            int max_iter, const double* __restrict__ out_scale, int n_out_scale) {
    extern __shared__ double sm[];
    // 1-D grid (problem-major) -- gridDim.y would cap the batch at 65535 problems
    const int pid = prob_ids[blockIdx.x / chunks];
    const int chunk = blockIdx.x % chunks;
    const NnlsProblem pr = d.prob[pid];
    const int s = pr.s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const double* Qg = d.Q + pr.q_off;
    if (QSMEM) {
        for (int i = threadIdx.x; i < s * s; i += blockDim.x) sm[i] = Qg[i];
        __syncthreads();
    }
    const double* Q = QSMEM ? sm : Qg;
    double* vec = sm + (QSMEM ? s * s : 0) + (size_t)warp * 6 * s;
    double *beta = vec, *grad = vec + s, *gb = vec + 2 * s, *bsave = vec + 3 * s, *db = vec + 4 * s,
           *tmp = vec + 5 * s;
    const double* sg = d.sign + pr.sel_off;
    const int ncols = min(pr.ncols < 0 ? d.m : pr.ncols, d.col1);
    const int col_end = min(d.col0 + (chunk + 1) * cols_per_cta, ncols);
    const double* oscale = out_scale ? out_scale + pr.scale_off : nullptr;
    const int oscale_len = pr.scale_len > 0 ? pr.scale_len : n_out_scale;
    const int* opos = pr.pos_off < 0 ? nullptr : d.pos + pr.pos_off;

    for (int col = d.col0 + chunk * cols_per_cta + warp; col < col_end; col += nwarps) {
        const double* g0 = d.grad0 + pr.row_off + (int64_t)col * d.ldg;
        for (int a = lane; a < s; a += 32) {
            const double g = g0[a];
            beta[a] = 0.0;
            grad[a] = g;
            gb[a] = (g > 0.0) ? 0.0 : g;   // optim.cpp:425-428 with beta == 0
        }
        __syncwarp();
        int iters = max_iter;
        bool conv = false;
        for (int l = 0; l < max_iter; ++l) {
            for (int a = lane; a < s; a += 32) bsave[a] = beta[a];
            // exact line search over the passive set (optim.cpp:442-459)
            wmatvec(Q, gb, tmp, s, lane);
            __syncwarp();
            {
                const double num = wdot(gb, gb, s, lane);
                const double den = wdot(gb, tmp, s, lane);
                const double a1 = num / den;
                if (alpha_ok(a1)) {
                    for (int a = lane; a < s; a += 32) {
                        beta[a] -= a1 * gb[a];
                        grad[a] -= a1 * tmp[a];
                    }
                    __syncwarp();
                    wnegcorr(Q, beta, grad, s, lane);
                }
            }
            __syncwarp();
            wgcd(Q, beta, grad, s, lane);
            // accelerated search (optim.cpp:465-483)
            for (int a = lane; a < s; a += 32) db[a] = bsave[a] - beta[a];
            __syncwarp();
            wmatvec(Q, db, tmp, s, lane);
            __syncwarp();
            {
                double nacc = 0.0;
                for (int a = lane; a < s; a += 32) {
                    const double gd = grad[a] * db[a];
                    nacc += gd * gd;
                }
                const double num = warp_sum(nacc);
                const double den = wdot(db, tmp, s, lane);
                const double a2 = num / den;
                if (alpha_ok(a2)) {
                    for (int a = lane; a < s; a += 32) {
                        beta[a] -= a2 * db[a];
                        grad[a] -= a2 * tmp[a];
                    }
                    __syncwarp();
                    wnegcorr(Q, beta, grad, s, lane);
                }
            }
            __syncwarp();
            wgcd(Q, beta, grad, s, lane);
            // KKT-masked gradient and stopping rule (optim.cpp:489-503)
            double nacc = 0.0;
            for (int a = lane; a < s; a += 32) {
                const double g = grad[a];
                const double v = ((beta[a] == 0.0) && (g > 0.0)) ? 0.0 : g;
                gb[a] = v;
                nacc += v * v;
            }
            __syncwarp();
            const double nrm = sqrt(warp_sum(nacc));
            if (nrm < eps) { conv = true; iters = l + 1; break; }
        }
        // unconverged columns return zeros (optim.cpp:423,502); rescale (optim.cpp:540) and
        // apply sign and R-recycled column scale (R/scregclust.R:1588)
        double* out = d.beta + pr.row_off + (int64_t)col * d.ldb;
        for (int a = lane; a < s; a += 32) {
            double v = conv ? beta[a] * d.dvec[pr.row_off + a] : 0.0;
            v = v * sg[a];
            if (oscale) v = v * oscale[((int64_t)a + (int64_t)col * s) % oscale_len];
            out[opos ? opos[a] : a] = v;
        }
        // rows up to the 16-row box of the stacked layout are exact zeros (the prediction kernel contracts over them)
        if (!opos) for (int a = s + lane; a < ((s + 15) & ~15); a += 32) out[a] = 0.0;
        if (lane == 0) d.iters[(int64_t)pid * d.m + col] = iters;
        __syncwarp();
    }
}

void nnls_solve(NnlsBatch& b, double eps, int max_iter, const double* out_scale, int n_out_scale,
                int num_sms, cudaStream_t st) {
    (void)num_sms;
    if (b.nprob == 0 || b.m == 0) return;
    if ((int)b.h_prob.size() != b.nprob) throw Error(-73, "nnls_solve: host problem descriptors missing");
    NnlsDev d = make_dev(b);
    if (d.col1 <= d.col0) return;
    // classify the problems by size from the host copy of the descriptors (no device round trip)
    std::vector<int> small_ids, large_ids;
    int smax_small = 0, smax_large = 0;
    auto classify = [&](int i) {
        const int s = b.h_prob[i].s;
        if (s <= 0) return;
        if (s <= 112) { small_ids.push_back(i); smax_small = std::max(smax_small, s); }
        else { large_ids.push_back(i); smax_large = std::max(smax_large, s); }
    };
    if (b.use_active) for (int i : b.active) classify(i);
    else for (int i = 0; i < b.nprob; ++i) classify(i);
    if (small_ids.empty() && large_ids.empty()) return;
    b.ids.ensure(2 * (size_t)b.nprob + 2);
    b.h_ids.clear();
    b.h_ids.insert(b.h_ids.end(), small_ids.begin(), small_ids.end());
    b.h_ids.insert(b.h_ids.end(), large_ids.begin(), large_ids.end());
    SCRC_CUDA(cudaMemcpyAsync(b.ids.p, b.h_ids.data(), b.h_ids.size() * 4, cudaMemcpyHostToDevice, st));
    const int cols_per_cta = 64;
    const unsigned chunks = (unsigned)ceil_div(d.col1 - d.col0, cols_per_cta);
    auto grid1d = [&](size_t nids) {
        const uint64_t g = (uint64_t)chunks * nids;
        if (g > 0x7fffffffULL) throw Error(-72, "NNLS: too many (problem, column chunk) pairs in one batch");
        return (unsigned)g;
    };
    ProfScope pscope(PROF_NNLS, st);
    if (!small_ids.empty()) {
        const int warps = 8;
        const size_t smem = ((size_t)smax_small * smax_small + (size_t)warps * 6 * smax_small) * 8;
        static SmemAttr attr;
        attr.ensure(nnls_kernel<true>, smem);
        nnls_kernel<true><<<grid1d(small_ids.size()), warps * 32, smem, st>>>(
            d, b.ids.p, cols_per_cta, (int)chunks, eps, max_iter, out_scale, n_out_scale);
        SCRC_LAUNCHED();
    }
    if (!large_ids.empty()) {
        int warps = (int)std::min<size_t>(8, (200 * 1024) / ((size_t)48 * smax_large));
        if (warps < 1) throw Error(-70, "NNLS: too many selected regulators for the shared-memory work vectors");
        const size_t smem = (size_t)warps * 6 * smax_large * 8;
        static SmemAttr attr;
        attr.ensure(nnls_kernel<false>, smem);
        nnls_kernel<false><<<grid1d(large_ids.size()), warps * 32, smem, st>>>(
            d, b.ids.p + small_ids.size(), cols_per_cta, (int)chunks, eps, max_iter, out_scale, n_out_scale);
        SCRC_LAUNCHED();
    }
    SCRC_CUDA(cudaGetLastError());
}

}  // namespace scrc
