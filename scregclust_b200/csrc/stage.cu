// Input staging on the device: sample split, per-stratum centring, constant-gene filter and
// scaling of a raw p x n expression matrix (replaces split_sample, R/scregclust.R:2280-2326, the
// constant-gene removal :899-964 and the scaling :996-1009).  One upload of the raw matrix; every
// later pass streams it from HBM.
#include "session.cuh"
#include "prof.cuh"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace scrc {

constexpr int ST_GENES = 128;   // genes per CTA (one thread each: warps read 32 consecutive genes of one cell)

// partial[(seg * CH + ch) * p + g] = sum over the cells of chunk ch of segment seg of E[g, cell]
__global__ void __launch_bounds__(ST_GENES) stage_stratum_sum_kernel(const double* __restrict__ E, int64_t p,
                                                                      const int* __restrict__ cells,
                                                                      const int* __restrict__ seg_off, int CH,
                                                                      double* partial) {
    const int64_t g = (int64_t)blockIdx.x * ST_GENES + threadIdx.x;
    const int seg = blockIdx.y, ch = blockIdx.z;
    const int c0 = seg_off[seg], c1 = seg_off[seg + 1];
    const int per = (c1 - c0 + CH - 1) / CH;
    const int a = c0 + ch * per, b = min(a + per, c1);
    if (g >= p) return;
    double s = 0.0;
    for (int c = a; c < b; ++c) s += E[g + (int64_t)cells[c] * p];
    partial[((int64_t)seg * CH + ch) * p + g] = s;
}
// smean[seg * p + g] = (sum of the chunk partials in chunk order) / count      (rowMeans, scregclust.R:2314)
__global__ void stage_stratum_mean_kernel(const double* __restrict__ partial, int64_t p, const int* __restrict__ seg_off,
                                          int CH, double* smean) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int seg = blockIdx.y;
    if (g >= p) return;
    double s = 0.0;
    for (int ch = 0; ch < CH; ++ch) s += partial[((int64_t)seg * CH + ch) * p + g];
    smean[(int64_t)seg * p + g] = s / (double)(seg_off[seg + 1] - seg_off[seg]);
}

__device__ __forceinline__ double staged_value(const double* __restrict__ E, int64_t p, int64_t g, int cell,
                                               const double* __restrict__ smean, const int* __restrict__ strat, int c) {
    double v = E[g + (int64_t)cell * p];
    if (smean) v -= smean[(int64_t)strat[c] * p + g];      // z_[, s] - rowMeans(z_[, idx])
    return v;
}

// pass 1 over the cells of one split: chunk sums, minima and maxima of the (stratum-centred) values
__global__ void __launch_bounds__(ST_GENES) stage_sum_minmax_kernel(const double* __restrict__ E, int64_t p,
                                                                     const int* __restrict__ cells,
                                                                     const int* __restrict__ strat, int ncell,
                                                                     const double* __restrict__ smean, int CH,
                                                                     double* psum, double* pmin, double* pmax) {
    const int64_t g = (int64_t)blockIdx.x * ST_GENES + threadIdx.x;
    const int ch = blockIdx.y;
    const int per = (ncell + CH - 1) / CH;
    const int a = ch * per, b = min(a + per, ncell);
    if (g >= p) return;
    double s = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int c = a; c < b; ++c) {
        const double v = staged_value(E, p, g, cells[c], smean, strat, c);
        s += v; lo = fmin(lo, v); hi = fmax(hi, v);
    }
    psum[(int64_t)ch * p + g] = s; pmin[(int64_t)ch * p + g] = lo; pmax[(int64_t)ch * p + g] = hi;
}
// mean[g] (fixed chunk order) and the constant flag: a column has sd == 0 exactly when all its values coincide
__global__ void stage_mean_kernel(const double* __restrict__ psum, const double* __restrict__ pmin,
                                  const double* __restrict__ pmax, int64_t p, int CH, int ncell, double* mean,
                                  unsigned char* is_const) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p) return;
    double s = 0.0, lo = INFINITY, hi = -INFINITY;
    for (int ch = 0; ch < CH; ++ch) {
        s += psum[(int64_t)ch * p + g];
        lo = fmin(lo, pmin[(int64_t)ch * p + g]); hi = fmax(hi, pmax[(int64_t)ch * p + g]);
    }
    if (mean) mean[g] = s / (double)ncell;
    if (!(lo < hi)) is_const[g] = 1;
}
// pass 2: chunk sums of squared deviations from the mean (two-pass variance like R's sd)
__global__ void __launch_bounds__(ST_GENES) stage_ss_kernel(const double* __restrict__ E, int64_t p,
                                                             const int* __restrict__ cells,
                                                             const int* __restrict__ strat, int ncell,
                                                             const double* __restrict__ smean,
                                                             const double* __restrict__ mean, int CH, double* pss) {
    const int64_t g = (int64_t)blockIdx.x * ST_GENES + threadIdx.x;
    const int ch = blockIdx.y;
    const int per = (ncell + CH - 1) / CH;
    const int a = ch * per, b = min(a + per, ncell);
    if (g >= p) return;
    const double mu = mean[g];
    double s = 0.0;
    for (int c = a; c < b; ++c) {
        const double d = staged_value(E, p, g, cells[c], smean, strat, c) - mu;
        s += d * d;
    }
    pss[(int64_t)ch * p + g] = s;
}
__global__ void stage_sd_kernel(const double* __restrict__ pss, int64_t p, int CH, int ncell, double* sd) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p) return;
    double s = 0.0;
    for (int ch = 0; ch < CH; ++ch) s += pss[(int64_t)ch * p + g];
    sd[g] = sqrt(s / (double)(ncell - 1));
}

// out[i, a] = (E[gene[a], cell[i]] - stratum mean - mean1[gene[a]]) (/ sd1[gene[a]] for regulators):
// a 32 x 32 tile is read gene-fastest (the raw layout) and written cell-fastest (the layout of Z).
__global__ void __launch_bounds__(256) stage_write_kernel(const double* __restrict__ E, int64_t p,
                                                          const int* __restrict__ genes, int ngene,
                                                          const int* __restrict__ cells, const int* __restrict__ strat,
                                                          int ncell, const double* __restrict__ smean,
                                                          const double* __restrict__ mean, const double* __restrict__ sd,
                                                          int scale, double* out, int64_t ldo) {
    __shared__ double tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int a0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int a = a0 + tx;
    const int64_t g = a < ngene ? genes[a] : -1;
    double mu = 0.0, s = 1.0;
    if (g >= 0) { mu = mean[g]; s = sd[g]; }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = i0 + ty + r * 8;
        double v = 0.0;
        if (g >= 0 && i < ncell) {
            v = staged_value(E, p, g, cells[i], smean, strat, i) - mu;
            if (scale) v = v / s;                                   // scale(): (x - mean) / sd
        }
        tile[ty + r * 8][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int aa = a0 + ty + r * 8;
        const int i = i0 + tx;
        if (aa < ngene && i < ncell) out[i + (int64_t)aa * ldo] = tile[tx][ty + r * 8];
    }
}
__global__ void stage_gather_kernel(const double* __restrict__ src, const int* __restrict__ idx, int n, double* dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

void Ctx::create_raw(int device, const double* expr, int64_t p, int64_t n, const int* is_regulator,
                     const int* split_indices, const int* stratification, int center, int K_,
                     unsigned char* keep_gene, int64_t* dims_out) {
    if (p <= 1 || n <= 2) throw Error(SCRC_ERR_ARG, "scrc_ctx_create_raw: non-positive dimension");
    if (p > 0x7fffffffLL || n > 0x7fffffffLL) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_ctx_create_raw: more than 2^31 genes or cells");
    dev.init(device);
    cudaStream_t st = dev.stream;
    AllocScope as(st);

    // ---- host-side index logic of split_sample (scregclust.R:2303-2323), explicit split_indices only
    std::vector<int> included;                       // cells that take part (split 1 or 2), original order
    for (int64_t c = 0; c < n; ++c)
        if (split_indices[c] == 1 || split_indices[c] == 2) included.push_back((int)c);
    const int ninc = (int)included.size();
    // stratum levels in increasing order (R: split() orders by the sorted factor levels)
    std::vector<int> level_of(ninc, 0);
    int S = 1;
    if (stratification) {
        std::vector<int> lv;
        for (int q = 0; q < ninc; ++q) lv.push_back(stratification[included[q]]);
        std::sort(lv.begin(), lv.end());
        lv.erase(std::unique(lv.begin(), lv.end()), lv.end());
        S = (int)lv.size();
        for (int q = 0; q < ninc; ++q)
            level_of[q] = (int)(std::lower_bound(lv.begin(), lv.end(), stratification[included[q]]) - lv.begin());
    }
    // With centring the reference rebuilds z_ by cbind over the strata, which reorders its columns
    // stratum by stratum, and then still selects split 1 by the positions computed before (:2309-2323).
    std::vector<int> perm(ninc);
    std::iota(perm.begin(), perm.end(), 0);
    if (center && S > 1)
        std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return level_of[a] < level_of[b]; });
    std::vector<char> pos_is1(ninc, 0);
    for (int q = 0; q < ninc; ++q) pos_is1[q] = split_indices[included[q]] == 1;
    std::vector<int> cells1, cells2, strat1, strat2;
    for (int q = 0; q < ninc; ++q) {
        const int src = perm[q];                     // the cell now sitting at position q
        if (pos_is1[q]) { cells1.push_back(included[src]); strat1.push_back(level_of[src]); }
        else { cells2.push_back(included[src]); strat2.push_back(level_of[src]); }
    }
    const int64_t n1_ = (int64_t)cells1.size(), n2_ = (int64_t)cells2.size();
    if (n1_ <= 1 || n2_ <= 0) throw Error(SCRC_ERR_ARG, "scrc_ctx_create_raw: both data splits need cells");
    // cells whose mean centres stratum s: intersect(s, is_split1) on the ORIGINAL positions (:2313)
    std::vector<int> mcells, seg_off(S + 1, 0);
    if (center) {
        for (int s = 0; s < S; ++s) {
            for (int q = 0; q < ninc; ++q)
                if (level_of[q] == s && pos_is1[q]) mcells.push_back(included[q]);
            seg_off[s + 1] = (int)mcells.size();
            if (seg_off[s + 1] == seg_off[s])
                throw Error(SCRC_ERR_ARG, "scrc_ctx_create_raw: a stratum has no cells in the first data split");
        }
    }

    // ---- device passes
    DevBuf<double> E((size_t)p * (size_t)n);
    SCRC_CUDA(cudaMemcpyAsync(E.p, expr, (size_t)p * n * 8, cudaMemcpyDefault, st));
    auto up = [&](const std::vector<int>& v, DevBuf<int>& d) {
        d.alloc(std::max<size_t>(v.size(), 1));
        if (!v.empty()) SCRC_CUDA(cudaMemcpyAsync(d.p, v.data(), v.size() * 4, cudaMemcpyHostToDevice, st));
    };
    DevBuf<int> d_c1, d_c2, d_s1, d_s2, d_mc, d_seg;
    up(cells1, d_c1); up(cells2, d_c2); up(strat1, d_s1); up(strat2, d_s2); up(mcells, d_mc); up(seg_off, d_seg);
    const unsigned gx = (unsigned)ceil_div(p, ST_GENES);
    const int CH = 32;                               // fixed: the summation order depends on the data only
    DevBuf<double> smean, part((size_t)std::max(S, 3) * CH * p);
    const double* sm = nullptr;
    if (center) {
        smean.alloc((size_t)S * p);
        stage_stratum_sum_kernel<<<dim3(gx, S, CH), ST_GENES, 0, st>>>(E.p, p, d_mc.p, d_seg.p, CH, part.p);
        SCRC_LAUNCHED();
        stage_stratum_mean_kernel<<<dim3((unsigned)ceil_div(p, 256), S), 256, 0, st>>>(part.p, p, d_seg.p, CH, smean.p);
        SCRC_LAUNCHED();
        sm = smean.p;
    }
    DevBuf<double> mean1(p), sd1(p);
    DevBuf<unsigned char> d_const(p);
    SCRC_CUDA(cudaMemsetAsync(d_const.p, 0, p, st));
    double* psum = part.p; double* pmin = part.p + (size_t)CH * p; double* pmax = part.p + 2 * (size_t)CH * p;
    stage_sum_minmax_kernel<<<dim3(gx, CH), ST_GENES, 0, st>>>(E.p, p, d_c1.p, d_s1.p, (int)n1_, sm, CH, psum, pmin, pmax);
    SCRC_LAUNCHED();
    stage_mean_kernel<<<(unsigned)ceil_div(p, 256), 256, 0, st>>>(psum, pmin, pmax, p, CH, (int)n1_, mean1.p, d_const.p);
    SCRC_LAUNCHED();
    stage_sum_minmax_kernel<<<dim3(gx, CH), ST_GENES, 0, st>>>(E.p, p, d_c2.p, d_s2.p, (int)n2_, sm, CH, psum, pmin, pmax);
    SCRC_LAUNCHED();
    stage_mean_kernel<<<(unsigned)ceil_div(p, 256), 256, 0, st>>>(psum, pmin, pmax, p, CH, (int)n2_, nullptr, d_const.p);
    SCRC_LAUNCHED();
    stage_ss_kernel<<<dim3(gx, CH), ST_GENES, 0, st>>>(E.p, p, d_c1.p, d_s1.p, (int)n1_, sm, mean1.p, CH, psum);
    SCRC_LAUNCHED();
    stage_sd_kernel<<<(unsigned)ceil_div(p, 256), 256, 0, st>>>(psum, p, CH, (int)n1_, sd1.p);
    SCRC_LAUNCHED();

    // ---- constant-gene filter (:899-964): a gene constant in either split is dropped
    std::vector<unsigned char> h_const(p);
    SCRC_CUDA(cudaMemcpyAsync(h_const.data(), d_const.p, p, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    std::vector<int> regs, tgts;
    for (int64_t g = 0; g < p; ++g) {
        const bool keep = !h_const[g];
        if (keep_gene) keep_gene[g] = keep ? 1 : 0;
        if (!keep) continue;
        if (is_regulator[g] == 1) regs.push_back((int)g);          // z_[...][is_regulator == 1L, ]
        else if (is_regulator[g] == 0) tgts.push_back((int)g);     // z_[...][is_regulator == 0L, ]
        else if (keep_gene) keep_gene[g] = 0;                      // in neither matrix
    }
    if (regs.empty() || tgts.empty()) throw Error(SCRC_ERR_ARG, "scrc_ctx_create_raw: no regulators or no targets left");
    begin(n1_, n2_, (int64_t)regs.size(), (int64_t)tgts.size(), K_);
    if (dims_out) { dims_out[0] = n1; dims_out[1] = n2; dims_out[2] = P; dims_out[3] = T; }

    DevBuf<int> d_regs, d_tgts;
    up(regs, d_regs); up(tgts, d_tgts);
    auto write = [&](DMat& Z, const DevBuf<int>& genes, int ng, const DevBuf<int>& cells, const DevBuf<int>& strat,
                     int64_t nc, int scale) {
        if (Z.ld != Z.rows) SCRC_CUDA(cudaMemsetAsync(Z.data(), 0, Z.bytes(), st));
        stage_write_kernel<<<dim3((unsigned)ceil_div(ng, 32), (unsigned)ceil_div(nc, 32)), 256, 0, st>>>(
            E.p, p, genes.p, ng, cells.p, strat.p, (int)nc, sm, mean1.p, sd1.p, scale, Z.data(), Z.ld);
        SCRC_LAUNCHED();
    };
    write(Z1r, d_regs, (int)P, d_c1, d_s1, n1, 1);
    write(Z2r, d_regs, (int)P, d_c2, d_s2, n2, 1);     // second split scaled like the first (:1003-1009)
    write(Z1t, d_tgts, (int)T, d_c1, d_s1, n1, 0);
    write(Z2t, d_tgts, (int)T, d_c2, d_s2, n2, 0);
    stage_gather_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(sd1.p, d_tgts.p, (int)T, sds_t.p);
    SCRC_LAUNCHED();
    SCRC_CUDA(cudaStreamSynchronize(st));              // E and the index buffers die here
    finish();
}

}  // namespace scrc
