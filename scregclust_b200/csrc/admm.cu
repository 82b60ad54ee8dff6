// Batched cooperative-lasso ADMM kernels.  Operation order per problem follows
// src/optim.cpp:134-308 step by step; only the linear
// solve is restructured (eigenbasis GEMMs, admm.cuh).
#include "admm.cuh"
#include "prof.cuh"

#include <algorithm>
#include <cstdlib>

namespace scrc {

struct AdmmDev {
    int P, C, nprob, nslab;
    int64_t ld, ldw;
    const int* col_prob;
    const int* c0;
    const int* c1;
    const double* lambda;
    double* rho;
    double* alpha;
    int* active;
    int* iters;
    int* n_active;
    unsigned long long* col_iters;   // [0] sum over sweeps of the number of active columns, [1] the same over update sweeps only
    const double* w;
    const double* xty;
    double *R, *W, *beta, *zeta, *mult, *beta_old, *zeta_old, *mult_old, *mult_hat_old;
    double* partial;
    const int* prob_R;     // rows per update slab of each problem (32 / 16 / 8)
    int max_nslab;
    int* tiles;            // compact lists of column tiles that hold an active problem (see admm_tile_lists)
    int n64, n128;         // number of 64- / 128-column tiles of the batch
    int nsm, bm_big;       // SMs of the device, row height of the large GEMM tile (160 or 128)
    int force_small;       // batch too small for large tiles: always 64 x 64
    double small_cost;     // time of a 64 x 64 tile relative to a large one (measured, see admm_solve)
};
constexpr int TILE_HDR = 4;   // tiles[0] = live 64-col tiles, [1] = live 128-col tiles, [2] = 128-col tiles given to the large-tile kernel,
                              // [3] = position in the 64-col list where the 64 x 64 kernel starts

void AdmmBatch::resize(int P_, int C_, int nprob_) {
    P = P_; C = C_; nprob = nprob_;
    nslab = (int)ceil_div(P, 8);   // worst case: 8-row slabs
    col_prob.ensure(C); prob_c0.ensure(nprob); prob_c1.ensure(nprob);
    prob_lambda.ensure(nprob); prob_rho.ensure(nprob); prob_alpha.ensure(nprob);
    prob_active.ensure(nprob); prob_iters.ensure(nprob);
    n_active.ensure(2); col_iters.ensure(2);   // n_active[0] = active problems, [1] = their columns
    w.reshape(P, nprob);
    DMat* mats[] = {&xty, &R, &W, &beta, &zeta, &mult, &beta_old, &zeta_old, &mult_old, &mult_hat_old};
    for (DMat* m : mats) m->reshape(P, C);
    partial.ensure((size_t)nprob * nslab * ADMM_NSUM);
    prob_R.ensure(nprob); class_ids.ensure(nprob);
    tiles.ensure(TILE_HDR + (size_t)ceil_div(C, 64) + (size_t)ceil_div(C, 128));
}

static AdmmDev make_dev(AdmmBatch& b) {
    AdmmDev d;
    d.P = b.P; d.C = b.C; d.nprob = b.nprob; d.nslab = b.nslab;
    d.ld = b.xty.ld; d.ldw = b.w.ld;
    d.col_prob = b.col_prob.p; d.c0 = b.prob_c0.p; d.c1 = b.prob_c1.p;
    d.lambda = b.prob_lambda.p; d.rho = b.prob_rho.p; d.alpha = b.prob_alpha.p;
    d.active = b.prob_active.p; d.iters = b.prob_iters.p; d.n_active = b.n_active.p;
    d.col_iters = b.col_iters.p;
    d.w = b.w.data(); d.xty = b.xty.data();
    d.R = b.R.data(); d.W = b.W.data(); d.beta = b.beta.data(); d.zeta = b.zeta.data();
    d.mult = b.mult.data(); d.beta_old = b.beta_old.data(); d.zeta_old = b.zeta_old.data();
    d.mult_old = b.mult_old.data(); d.mult_hat_old = b.mult_hat_old.data();
    d.partial = b.partial.p;
    d.prob_R = b.prob_R.p; d.max_nslab = b.nslab;
    d.tiles = b.tiles.p; d.n64 = (int)ceil_div(b.C, 64); d.n128 = (int)ceil_div(b.C, 128);
    d.nsm = 148; d.bm_big = 128; d.force_small = 0; d.small_cost = 0.26;
    return d;
}

// ------------------------------------------------------------------ compact lists of live column tiles
// tiles[0] / tiles[1] = number of 64- / 128-column tiles that contain a column of an active problem, followed by
// the two lists (increasing tile index).  The GEMM CTAs walk these lists round-robin, so the tiles of converged
// problems cost nothing -- not even an imbalance between the CTAs that would have drawn them.  One CTA.
__device__ void admm_tile_lists(const AdmmDev& d) {
    __shared__ int s_wtot[32];
    __shared__ int s_run;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int pass = 0; pass < 2; ++pass) {
        const int bn = pass == 0 ? 64 : 128;
        const int nt = pass == 0 ? d.n64 : d.n128;
        int* list = d.tiles + TILE_HDR + (pass == 0 ? 0 : d.n64);
        if (threadIdx.x == 0) s_run = 0;
        __syncthreads();
        for (int t0 = 0; t0 < nt; t0 += blockDim.x) {
            const int t = t0 + threadIdx.x;
            int live = 0;
            if (t < nt) {
                const int cf = t * bn, cl = min(cf + bn, d.C) - 1;
                for (int q = d.col_prob[cf]; q <= d.col_prob[cl]; ++q) live |= d.active[q];
            }
            const unsigned bal = __ballot_sync(0xffffffffu, live);
            if (lane == 0) s_wtot[warp] = __popc(bal);
            __syncthreads();
            int off = s_run, tot = 0;
            for (int w = 0; w < nw; ++w) { if (w < warp) off += s_wtot[w]; tot += s_wtot[w]; }
            if (live) list[off + __popc(bal & ((1u << lane) - 1u))] = t;
            __syncthreads();
            if (threadIdx.x == 0) s_run += tot;
            __syncthreads();
        }
        if (threadIdx.x == 0) d.tiles[pass] = s_run;
        __syncthreads();
    }
    // Tile shapes of the next sweep, from the exact live-tile lists.  The large-tile kernel takes a PREFIX of the
    // 128-column list sized to whole waves of the device (tiles[2] entries), the 64 x 64 kernel takes the live
    // 64-column tiles to the right of that prefix (from list position tiles[3]): the last, partly filled wave of
    // large tiles is replaced by waves of tiles a fifth of the size.  Candidates = every whole number of large waves
    // near the total (and none at all); cost in units of one large tile.  Every split gives the same bits: the
    // accumulation order along k does not depend on the tile.
    if (threadIdx.x < 32) {
        const int t64 = d.tiles[0], t128 = d.tiles[1];
        const int* l64 = d.tiles + TILE_HDR;
        const int* l128 = l64 + d.n64;
        const long long rb = (d.P + d.bm_big - 1) / d.bm_big, rs = (d.P + 63) / 64;
        const long long wb = (rb * t128 + d.nsm - 1) / d.nsm;            // waves if every tile were large
        // lane 0: no large tiles; lanes 1..31: wb - 30 + lane whole waves of large tiles (clamped)
        long long w = (lane == 0) ? 0 : wb - 31 + lane;
        if (w < 0) w = 0;
        long long nb = d.force_small ? 0 : (w * d.nsm) / rb;
        if (nb > t128) nb = t128;
        int start = t64;                                                  // first 64-column tile right of the prefix
        if (nb < t128) {
            const int bound = l128[nb] * 2;
            int lo = 0, hi = t64;
            while (lo < hi) { const int mid = (lo + hi) >> 1; if (l64[mid] < bound) lo = mid + 1; else hi = mid; }
            start = lo;
        }
        const double cost = (double)((rb * nb + d.nsm - 1) / d.nsm) +
                            (double)((rs * (t64 - start) + d.nsm - 1) / d.nsm) * d.small_cost;
        // arg-min over the warp; ties go to the candidate with more large tiles
        double bc = cost; long long bn = nb; int bs = start;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double oc = __shfl_xor_sync(0xffffffffu, bc, o);
            const long long on = __shfl_xor_sync(0xffffffffu, bn, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (oc < bc || (oc == bc && on > bn)) { bc = oc; bn = on; bs = os; }
        }
        if (lane == 0) { d.tiles[2] = (int)bn; d.tiles[3] = bs; }
    }
}
__global__ void __launch_bounds__(1024) admm_tiles_kernel(AdmmDev d) { admm_tile_lists(d); }

// ------------------------------------------------------------------ init state
__global__ void admm_init_prob_kernel(AdmmDev d, double rho0, double alpha0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < d.nprob) {
        d.rho[p] = rho0; d.alpha[p] = alpha0;
        d.active[p] = (d.c1[p] > d.c0[p]) ? 1 : 0;
        d.iters[p] = 0;
    }
    if (p == 0) { d.n_active[0] = d.nprob; d.n_active[1] = d.C; d.col_iters[0] = 0ull; d.col_iters[1] = 0ull; }
}

// ------------------------------------------------------------------ rhs = xty + rho*zeta + mult   (optim.cpp:143)
__global__ void __launch_bounds__(256) admm_rhs_kernel(AdmmDev d, const double* __restrict__ zeta,
                                                       const double* __restrict__ mult) {
    const int c = blockIdx.x;
    const int p = d.col_prob[c];
    if (!d.active[p]) return;
    const double rho = d.rho[p];
    const int64_t base = (int64_t)c * d.ld;
    for (int i = threadIdx.x; i < d.P; i += blockDim.x) {
        const int64_t idx = base + i;
        d.R[idx] = (d.xty[idx] + rho * zeta[idx]) + mult[idx];
    }
}

// ------------------------------------------------------------------ GEMM policies
// STEP 1: W = diag(1/(lam + rho_col)) V' R        STEP 2: beta = V W
template <int BM_, int BN_, int STAGES_, int STEP>
struct AdmmGemmPolicy {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_;
    struct Params {
        CUtensorMap ta, tb;
        double* out;
        int64_t ld;
        int P, C;
        const int* col_prob;
        const int* active;
        const double* rho;
        const double* lam;
        const int* hdr;          // tile header written by admm_tile_lists (counts and the split between the shapes)
        const int* tile_list;    // live column tiles of this tile width
    };
    __device__ static int tiles_m(const Params& p) { return (p.P + BM - 1) / BM; }
    // both shapes are launched every sweep: the large one walks the first hdr[2] entries of the 128-column list,
    // the 64 x 64 one the 64-column list from position hdr[3]; either may find nothing to do and exits
    __device__ static int first(const Params& p) { return BN == 64 ? p.hdr[3] : 0; }
    __device__ static int num_tiles(const Params& p) { return tiles_m(p) * (BN == 64 ? p.hdr[0] - p.hdr[3] : p.hdr[2]); }
    __device__ static bool tile(const Params& p, int t, TileInfo& ti) {
        const int tm = t % tiles_m(p), tn = p.tile_list[first(p) + t / tiles_m(p)];
        ti.ta = &p.ta; ti.tb = &p.tb;
        ti.m0 = tm * BM; ti.n0 = tn * BN; ti.k0 = 0; ti.k1 = p.P;
        return true;
    }
    template <int MT, int NT>
    __device__ static void epilogue(const Params& p, const TileInfo& ti, double (&acc)[MT][NT][2],
                                    int warp_m, int warp_n, int lane) {
        // (Measured and dropped: one real division per (row, problem) plus the exact quotient correction per element
        // instead of a division per element -- the cache of the last divisor serialises the col_prob / active / rho
        // loads of the epilogue; 33.4 -> 32.3 TFLOP/s over a config-3 grid.)
        for_each_acc<MT, NT>(acc, warp_m, warp_n, lane, [&](int rr, int cc, double v0, double v1) {
            const int row = ti.m0 + rr;
            if (row >= p.P) return;
            const double lam = (STEP == 1) ? p.lam[row] : 0.0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = ti.n0 + cc + e;
                if (col >= p.C) continue;
                const int q = p.col_prob[col];
                if (!p.active[q]) continue;
                double v = e ? v1 : v0;
                if (STEP == 1) v = v / (lam + p.rho[q]);
                p.out[row + (int64_t)col * p.ld] = v;
            }
        });
    }
};

template <int BM, int BN, int STAGES, int STEP>
static void launch_admm_gemm(const AdmmDev& d, const DMat& A, const double* Bm, double* out,
                             int num_sms, const double* lam, cudaStream_t stream) {
    using Pol = AdmmGemmPolicy<BM, BN, STAGES, STEP>;
    typename Pol::Params p;
    p.ta = make_tmap_f64(A.data(), d.P, d.P, A.ld, BM);
    p.tb = make_tmap_f64(Bm, d.P, d.C, d.ld, BN);
    p.out = out; p.ld = d.ld; p.P = d.P; p.C = d.C;
    p.col_prob = d.col_prob; p.active = d.active; p.rho = d.rho; p.lam = lam;
    static_assert(BN == 64 || BN == 128, "tile lists exist for 64- and 128-column tiles");
    p.hdr = d.tiles;
    p.tile_list = d.tiles + TILE_HDR + (BN == 64 ? 0 : d.n64);
    constexpr int smem = GemmSmem<BM, BN, STAGES>::TOTAL;
    static SmemAttr attr;
    attr.ensure(gemm_tn_kernel<Pol>, smem);
    const int64_t tiles = ceil_div(d.P, BM) * ceil_div(d.C, BN);
    const int grid = (int)std::min<int64_t>(tiles, num_sms);
    {
        ProfScope ps(PROF_GEMM_ADMM, stream);
        gemm_tn_kernel<Pol><<<grid, GEMM_THREADS, smem, stream>>>(p);
    }
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// ------------------------------------------------------------------ relax / shrink / dual update / reductions
// optim.cpp:146-176 and the sums needed by :192-254, one CTA per (R-row slab, problem).
// Lane layout: row = lane % R, column sub-lane = lane / R, so a warp step covers 32/R columns and
// 8 warps cover 8*32/R columns.  R is chosen per problem so that the whole slab of the three
// input matrices (beta, zeta, mult: R x m x 24 B <= 96 KB) stays RESIDENT IN SHARED MEMORY between
// the group-norm pass and the shrink / dual-update pass (m <= 128: R = 32, m <= 256: R = 16,
// otherwise R = 8 with the columns beyond 512 re-read from L2/HBM).  Loads are issued in batches
// before any store so that every thread keeps >= 12 independent 8-byte loads in flight.
// FUSE_RHS (sweeps on which rho cannot change): also writes next sweep's rhs = xty + rho*zeta + mult.
constexpr int UPD_CACHE = 4096;   // doubles per cached array (R * cached columns)
constexpr int UPD_BOXC = 64;      // columns per TMA box

struct UpdMaps { CUtensorMap b, z, m; };   // beta, zeta, mult viewed as (P x C) with an (R x 64) box
// zeta and mult live in two buffers each (the matrices called zeta / zeta_old, mult / mult_old).  A rho-adaptation
// sweep must leave "the snapshot" equal to the values it just produced (optim.cpp:282-285), so instead of storing
// them twice it writes them ONCE, over the previous snapshot, and that buffer becomes both the current state and
// the snapshot; the first sweep after it writes the other buffer, later plain sweeps update in place.
struct UpdPtrs {
    const double *z_in, *m_in;    // current zeta / mult (the buffers the TMA maps z, m point at)
    double *z_out, *m_out;        // where this sweep's zeta / mult go
    const double *z_snap, *m_snap;   // zeta_old / mult_old as of the last rho-adaptation sweep (UPDATE_ITER only)
};

template <int R, bool UPDATE_ITER, bool FUSE_RHS>
__global__ void __launch_bounds__(ADMM_UPD_WARPS * 32) admm_update_kernel(AdmmDev d, const int* __restrict__ prob_ids,
                                                                          const __grid_constant__ UpdMaps maps,
                                                                          const UpdPtrs ptr) {
    constexpr int CPW = 32 / R;
    constexpr int CS = UPD_CACHE / R;
    constexpr int STEP = ADMM_UPD_WARPS * CPW;
    constexpr int U = UPDATE_ITER ? 2 : 4;
    extern __shared__ __align__(128) double upd_sm[];
    double* sb = upd_sm;
    double* sz = upd_sm + UPD_CACHE;
    double* smu = upd_sm + 2 * UPD_CACHE;
    __shared__ double s_norm[ADMM_UPD_WARPS][32][2];
    __shared__ double s_sum[ADMM_UPD_WARPS][ADMM_NSUM];
    __shared__ __align__(8) uint64_t s_bar[UPD_CACHE / (8 * UPD_BOXC)];   // one per staged box (R >= 8)

    const int p = prob_ids[blockIdx.y];
    if (!d.active[p]) return;
    const int slab = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row_l = lane % R, csub = lane / R;
    const int row = slab * R + row_l;
    const bool rv = row < d.P;
    const int c0 = d.c0[p];
    const int m = d.c1[p] - c0;
    const double rho = d.rho[p], alpha = d.alpha[p], lam = d.lambda[p];
    const int64_t base = row + (int64_t)c0 * d.ld;
    // mult / rho (optim.cpp:149,160) is needed for every element in both passes: one real division
    // per thread gives RN(1 / rho), the elements use the exact quotient correction (common.cuh).
    const double inv_rho = 1.0 / rho;
    const bool rho_ok = quot_in_range(rho);

    // ---- stage the slab of beta / zeta / mult with TMA: every box is in flight at once, no registers
    // are tied up, and the image in shared memory is [column][row] exactly as the passes index it.
    // (Boxes may run past the problem's last column or past row P; those entries are never used.)
    const int mc = min(m, CS);
    const int nbox = (mc + UPD_BOXC - 1) / UPD_BOXC;
    if (threadIdx.x == 0) {
        for (int bx = 0; bx < nbox; ++bx) mbar_init(&s_bar[bx], 1);
        mbar_fence_init();
        for (int bx = 0; bx < nbox; ++bx) {       // one barrier per 64-column box: pass 1 starts on the first arrival
            const int cc = c0 + bx * UPD_BOXC;
            mbar_expect_tx(&s_bar[bx], (uint32_t)(3 * UPD_BOXC * R * 8));
            tma_load_2d(sb + bx * UPD_BOXC * R, &maps.b, slab * R, cc, &s_bar[bx]);
            tma_load_2d(sz + bx * UPD_BOXC * R, &maps.z, slab * R, cc, &s_bar[bx]);
            tma_load_2d(smu + bx * UPD_BOXC * R, &maps.m, slab * R, cc, &s_bar[bx]);
        }
    }
    __syncthreads();                       // barriers initialised and armed before anyone polls them

    // ---- pass 1: group norms of the positive / negative parts (optim.cpp:149-158), box by box as the
    // loads land (every lane walks its columns in increasing order, exactly as a single loop would)
    double np2 = 0.0, nn2 = 0.0;
    auto norm_term = [&](double b, double z, double mu) {
        const double br = alpha * b + (1.0 - alpha) * z;
        const double qm = (rho_ok && (quot_in_range(mu) || mu == 0.0)) ? quot_exact(mu, rho, inv_rho) : mu / rho;
        const double zn = br - qm;
        const double pos = fmax(zn, 0.0), neg = fmin(zn, 0.0);
        np2 += pos * pos;
        nn2 += neg * neg;
    };
    int lc = warp * CPW + csub;
    for (int bx = 0; bx < nbox; ++bx) {
        const int end = min((bx + 1) * UPD_BOXC, mc);
        if (lc - csub < end) mbar_wait_warp(&s_bar[bx], 0, lane);    // warp-uniform: the warp has columns in this box
#pragma unroll 4
        for (; lc < end; lc += STEP) {
            double b = 0.0, z = 0.0, mu = 0.0;
            if (rv) { b = sb[lc * R + row_l]; z = sz[lc * R + row_l]; mu = smu[lc * R + row_l]; }
            norm_term(b, z, mu);
        }
    }
#pragma unroll 4
    for (; lc < m; lc += STEP) {               // columns beyond the cached slab
        double b = 0.0, z = 0.0, mu = 0.0;
        if (rv) {
            const int64_t idx = base + (int64_t)lc * d.ld;
            b = d.beta[idx]; z = ptr.z_in[idx]; mu = ptr.m_in[idx];
        }
        norm_term(b, z, mu);
    }
#pragma unroll
    for (int o = R; o < 32; o <<= 1) {
        np2 += __shfl_xor_sync(0xffffffffu, np2, o);
        nn2 += __shfl_xor_sync(0xffffffffu, nn2, o);
    }
    if (csub == 0) { s_norm[warp][row_l][0] = np2; s_norm[warp][row_l][1] = nn2; }
    __syncthreads();
    double tp = 0.0, tn = 0.0;
#pragma unroll
    for (int w = 0; w < ADMM_UPD_WARPS; ++w) { tp += s_norm[w][row_l][0]; tn += s_norm[w][row_l][1]; }
    const double wi = rv ? d.w[row + (int64_t)p * d.ldw] : 0.0;
    // optim.cpp:151-158 (a zero norm gives 1 - inf -> 0)
    const double sp = fmax(1.0 - lam * wi / (rho * sqrt(tp)), 0.0);
    const double sn = fmax(1.0 - lam * wi / (rho * sqrt(tn)), 0.0);

    // ---- pass 2: shrink, dual update, residual / adaptation sums, stores
    double s[ADMM_NSUM];
#pragma unroll
    for (int k = 0; k < ADMM_NSUM; ++k) s[k] = 0.0;
    for (int lc0 = warp * CPW + csub; lc0 < m; lc0 += STEP * U) {
        double b[U], z[U], mu[U], xt[U], bo[U], zo[U], mo[U], mho[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int lc = lc0 + u * STEP;
            b[u] = z[u] = mu[u] = xt[u] = bo[u] = zo[u] = mo[u] = mho[u] = 0.0;
            if (lc < m) {
                const int64_t idx = base + (int64_t)lc * d.ld;
                if (lc < CS) { if (rv) { b[u] = sb[lc * R + row_l]; z[u] = sz[lc * R + row_l]; mu[u] = smu[lc * R + row_l]; } }
                else if (rv) { b[u] = d.beta[idx]; z[u] = ptr.z_in[idx]; mu[u] = ptr.m_in[idx]; }
                if (rv) {
                    if (FUSE_RHS) xt[u] = d.xty[idx];
                    if (UPDATE_ITER) { bo[u] = d.beta_old[idx]; zo[u] = ptr.z_snap[idx]; mo[u] = ptr.m_snap[idx]; mho[u] = d.mult_hat_old[idx]; }
                }
            }
        }
        bool fastq = rho_ok;
#pragma unroll
        for (int u = 0; u < U; ++u) fastq &= quot_in_range(mu[u]) || mu[u] == 0.0;
        double qm[U];
        if (fastq) {
#pragma unroll
            for (int u = 0; u < U; ++u) qm[u] = quot_exact(mu[u], rho, inv_rho);
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) qm[u] = mu[u] / rho;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int lc = lc0 + u * STEP;
            if (lc >= m || !rv) continue;
            const int64_t idx = base + (int64_t)lc * d.ld;
            const double br = alpha * b[u] + (1.0 - alpha) * z[u];
            double zn = br - qm[u];
            zn = (zn >= 0.0) ? zn * sp : zn * sn;                       // optim.cpp:160-168
            const double mn = mu[u] + rho * (-br + zn);                 // optim.cpp:171
            const double pr = -b[u] + zn;                               // optim.cpp:175
            const double du = rho * (z[u] - zn);                        // optim.cpp:176
            s[0] += pr * pr; s[1] += du * du; s[2] += b[u] * b[u]; s[3] += zn * zn; s[4] += mn * mn;
            if (UPDATE_ITER) {
                const double mh = mn + rho * (-b[u] + z[u]);            // optim.cpp:203
                const double dmh = mh - mho[u];
                const double dh = b[u] - bo[u];
                const double dm = mn - mo[u];
                const double dg = zo[u] - z[u];
                s[5] += dmh * dmh; s[6] += dh * dh; s[7] += dh * dmh;
                s[8] += dm * dm; s[9] += dg * dg; s[10] += dg * dm;
                d.beta_old[idx] = b[u]; d.mult_hat_old[idx] = mh;       // optim.cpp:282-285 (zeta_old, mult_old: see UpdPtrs)
            }
            ptr.z_out[idx] = zn;                                        // optim.cpp:292
            ptr.m_out[idx] = mn;
            if (FUSE_RHS) d.R[idx] = (xt[u] + rho * zn) + mn;           // optim.cpp:143 of the next sweep
        }
    }
#pragma unroll
    for (int k = 0; k < ADMM_NSUM; ++k) {
        const double v = warp_sum(s[k]);
        if (lane == 0) s_sum[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < ADMM_NSUM) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < ADMM_UPD_WARPS; ++w) v += s_sum[w][threadIdx.x];
        d.partial[((int64_t)p * d.max_nslab + slab) * ADMM_NSUM + threadIdx.x] = v;
    }
}

// ------------------------------------------------------------------ stop test + adaptive rho / alpha
// optim.cpp:192-196 (stop), :199-289 (spectral step-size / relaxation update), :295-305.
// One CTA; one warp per problem (lanes stride the slab partials, fixed shuffle tree).
__global__ void __launch_bounds__(1024) admm_finalize_kernel(AdmmDev d, AdmmOpts o, int it, int is_update) {
    __shared__ int s_cnt;
    __shared__ unsigned long long s_cols;
    __shared__ int s_act_cols;
    if (threadIdx.x == 0) { s_cnt = 0; s_cols = 0ull; s_act_cols = 0; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int p = warp; p < d.nprob; p += nwarps) {
        if (!d.active[p]) continue;
        const int nsl = (d.P + d.prob_R[p] - 1) / d.prob_R[p];
        double s[ADMM_NSUM];
#pragma unroll
        for (int k = 0; k < ADMM_NSUM; ++k) s[k] = 0.0;
        for (int sl = lane; sl < nsl; sl += 32) {
            const double* q = d.partial + ((int64_t)p * d.max_nslab + sl) * ADMM_NSUM;
#pragma unroll
            for (int k = 0; k < ADMM_NSUM; ++k) s[k] += q[k];
        }
#pragma unroll
        for (int k = 0; k < ADMM_NSUM; ++k) s[k] = warp_sum(s[k]);
        if (lane != 0) continue;
        atomicAdd(&s_cols, (unsigned long long)(d.c1[p] - d.c0[p]));   // columns swept this iteration
        const double n_primal = sqrt(s[0]), n_dual = sqrt(s[1]);
        const double n_beta = sqrt(s[2]), n_zeta = sqrt(s[3]), n_mult = sqrt(s[4]);
        const bool conv = (n_primal <= fmax(o.eps_rel * fmax(n_beta, n_zeta), o.eps_abs)) &&
                          (n_dual <= fmax(o.eps_rel * n_mult, o.eps_abs));
        if (conv) {
            d.active[p] = 0;
            d.iters[p] = it;
            continue;
        }
        if (is_update) {
            double rho = d.rho[p], alpha;
            const double n_dmh = sqrt(s[5]), n_dh = sqrt(s[6]), n_dm = sqrt(s[8]), n_dg = sqrt(s[9]);
            double a = 0.0, a_corr = 0.0;
            if (n_dmh > 0.0 && n_dh > 0.0) {
                const double hm = s[7];
                const double a_sd = s[5] / hm;
                const double a_mg = hm / s[6];
                a = (2.0 * a_mg > a_sd) ? a_mg : a_sd - a_mg / 2.0;
                a_corr = hm / (n_dh * n_dmh);
            }
            double b = 0.0, b_corr = 0.0;
            if (n_dm > 0.0 && n_dg > 0.0) {
                const double gm = s[10];
                const double b_sd = s[8] / gm;
                const double b_mg = gm / s[9];
                b = (2.0 * b_mg > b_sd) ? b_mg : b_sd - b_mg / 2.0;
                b_corr = gm / (n_dg * n_dm);
            }
            const bool ga = a_corr > o.eps_corr, gb = b_corr > o.eps_corr;
            if (ga && gb) {
                rho = sqrt(a * b);
                alpha = 1.0 + 2.0 / (sqrt(a * b) * (1.0 / a + 1.0 / b));
            } else if (ga && !gb) {
                rho = a; alpha = 1.9;
            } else if (!ga && gb) {
                rho = b; alpha = 1.1;
            } else {
                alpha = 1.5;
            }
            d.rho[p] = rho;
            d.alpha[p] = alpha;
        }
        if (it + 1 > o.max_iter) {
            d.active[p] = 0;
            d.iters[p] = it + 1;
            continue;
        }
        atomicAdd(&s_cnt, 1);
        atomicAdd(&s_act_cols, d.c1[p] - d.c0[p]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        d.n_active[0] = s_cnt; d.n_active[1] = s_act_cols;
        d.col_iters[0] += s_cols;
        if (is_update) d.col_iters[1] += s_cols;
    }
    __syncthreads();            // the new activity flags are visible to the whole CTA
    admm_tile_lists(d);
}

// ------------------------------------------------------------------ driver
int admm_solve(AdmmBatch& b, const EigBasis& eig, const AdmmOpts& o, bool cold, int num_sms,
               cudaStream_t stream, int* h_scratch, const std::atomic<int>* cancel) {
    b.aborted = false;
    if (b.nprob == 0 || b.C == 0) return 0;
    AdmmDev d = make_dev(b);
    const bool big = ceil_div(b.P, 128) * ceil_div(b.C, 128) >= num_sms;
    // 160-row tiles when they waste fewer padded rows than 128-row tiles (e.g. P = 1600: 0 % vs 4 %)
    const bool tall = (getenv("SCRC_NO_TALL_TILES") == nullptr) &&
                      (ceil_div(b.P, 160) * 160 < ceil_div(b.P, 128) * 128);
    d.nsm = num_sms; d.bm_big = tall ? 160 : 128; d.force_small = big ? 0 : 1;
    // time of one 64 x 64 tile relative to one large tile: MAC ratio / relative efficiency.  0.85 from a sweep on
    // B200 (tools/tile_cost_sweep.py, profiles/r02_tile_cost_sweep.log: best for the 20-value grid and for a
    // 3-value grid alike); SCRC_SMALL_TILE_COST overrides for tuning
    d.small_cost = (64.0 * 64.0) / ((double)d.bm_big * 128.0) / 0.85;
    if (const char* e = getenv("SCRC_SMALL_TILE_COST")) d.small_cost = atof(e);
    if (cold) {
        DMat* z[] = {&b.beta, &b.zeta, &b.beta_old, &b.zeta_old};
        for (DMat* m : z) SCRC_CUDA(cudaMemsetAsync(m->data(), 0, m->bytes(), stream));
    }
    DMat* z2[] = {&b.mult, &b.mult_old, &b.mult_hat_old};
    for (DMat* m : z2) SCRC_CUDA(cudaMemsetAsync(m->data(), 0, m->bytes(), stream));
    admm_init_prob_kernel<<<(b.nprob + 127) / 128, 128, 0, stream>>>(d, o.rho0, o.alpha0);
    SCRC_LAUNCHED();
    admm_tiles_kernel<<<1, 1024, 0, stream>>>(d);
    SCRC_LAUNCHED();

    // problems grouped by slab height (see admm_update_kernel); lists live in class_ids
    if ((int)b.h_cols.size() != b.nprob) throw Error(-71, "admm_solve: host column counts missing");
    std::vector<int> hR(b.nprob), ids[3];
    for (int p = 0; p < b.nprob; ++p) {
        const int m = b.h_cols[p];
        const int cls = (m <= 128) ? 0 : (m <= 256) ? 1 : 2;
        hR[p] = 32 >> cls;
        if (m > 0) ids[cls].push_back(p);
    }
    std::vector<int> flat;
    int off[3];
    for (int c = 0; c < 3; ++c) { off[c] = (int)flat.size(); flat.insert(flat.end(), ids[c].begin(), ids[c].end()); }
    SCRC_CUDA(cudaMemcpyAsync(b.prob_R.p, hR.data(), b.nprob * 4, cudaMemcpyHostToDevice, stream));
    if (!flat.empty()) SCRC_CUDA(cudaMemcpyAsync(b.class_ids.p, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, stream));
    SCRC_CUDA(cudaStreamSynchronize(stream));   // hR / flat are stack temporaries
    constexpr int upd_smem = 3 * UPD_CACHE * 8;
    {
        static SmemAttr attr[12];
#define SCRC_SET_UPD(R_, I_) \
        attr[I_].ensure(admm_update_kernel<R_, false, false>, upd_smem); attr[I_ + 1].ensure(admm_update_kernel<R_, false, true>, upd_smem); \
        attr[I_ + 2].ensure(admm_update_kernel<R_, true, false>, upd_smem); attr[I_ + 3].ensure(admm_update_kernel<R_, true, true>, upd_smem);
        SCRC_SET_UPD(32, 0) SCRC_SET_UPD(16, 4) SCRC_SET_UPD(8, 8)
#undef SCRC_SET_UPD
    }
    // zeta / mult buffer pairs (UpdPtrs): index 0 = the matrices named zeta / mult, 1 = zeta_old / mult_old
    double* zbuf[2] = {d.zeta, d.zeta_old};
    double* mbuf[2] = {d.mult, d.mult_old};
    UpdMaps umaps[3][2];
    for (int c = 0; c < 3; ++c) {
        if (ids[c].empty()) continue;
        const int R = 32 >> c;
        for (int w = 0; w < 2; ++w) {
            umaps[c][w].b = make_tmap_f64_plain(d.beta, b.P, b.C, d.ld, R, UPD_BOXC);
            umaps[c][w].z = make_tmap_f64_plain(zbuf[w], b.P, b.C, d.ld, R, UPD_BOXC);
            umaps[c][w].m = make_tmap_f64_plain(mbuf[w], b.P, b.C, d.ld, R, UPD_BOXC);
        }
    }
    // cur: buffer holding the current zeta / mult; snap: buffer holding the snapshot of the last rho-adaptation sweep.
    // At the start both matrices of a pair hold the initial value (zeros, or the warm start), so cur == snap == 0.
    int cur = 0, snap = 0;
    auto launch_update = [&](bool upd, bool fuse) {
        // rho-adaptation sweep: write over the snapshot (every element is read by its own thread first); the first
        // sweep after one: write the other buffer; otherwise in place
        const int outb = upd ? snap : (cur == snap ? 1 - cur : cur);
        UpdPtrs ptr{zbuf[cur], mbuf[cur], zbuf[outb], mbuf[outb], zbuf[snap], mbuf[snap]};
        for (int c = 0; c < 3; ++c) {
            if (ids[c].empty()) continue;
            const int R = 32 >> c;
            const dim3 grid((unsigned)ceil_div(b.P, R), (unsigned)ids[c].size());
            const int* lst = b.class_ids.p + off[c];
            const UpdMaps& um = umaps[c][cur];
#define SCRC_UPD(R_) \
            if (upd && fuse) admm_update_kernel<R_, true, true><<<grid, ADMM_UPD_WARPS * 32, upd_smem, stream>>>(d, lst, um, ptr); \
            else if (upd) admm_update_kernel<R_, true, false><<<grid, ADMM_UPD_WARPS * 32, upd_smem, stream>>>(d, lst, um, ptr); \
            else if (fuse) admm_update_kernel<R_, false, true><<<grid, ADMM_UPD_WARPS * 32, upd_smem, stream>>>(d, lst, um, ptr); \
            else admm_update_kernel<R_, false, false><<<grid, ADMM_UPD_WARPS * 32, upd_smem, stream>>>(d, lst, um, ptr);
            if (R == 32) { SCRC_UPD(32) } else if (R == 16) { SCRC_UPD(16) } else { SCRC_UPD(8) }
#undef SCRC_UPD
            SCRC_LAUNCHED();
        }
        cur = outb;
        if (upd) snap = outb;
    };
    int it = 1;
    int sweeps = 0;
    bool need_rhs = true;   // the first sweep and every sweep after a possible rho change build the rhs explicitly
    while (true) {
        if (need_rhs) {
            ProfScope ps(PROF_ADMM_ELEM, stream);
            admm_rhs_kernel<<<b.C, 256, 0, stream>>>(d, zbuf[cur], mbuf[cur]);
            SCRC_LAUNCHED();
        }
        // Tail of the solve: once the surviving problems fill less than about half of the SMs with
        // large tiles, 64 x 64 tiles give 4-5x the CTAs (the accumulation order along k does not depend
        // on the tile shape, so the bits are the same).
        // The tile shape of each sweep is chosen on the device from the exact live-tile counts (admm_tile_lists);
        // both shapes are launched, the one not chosen exits at once.
        for (int step = 1; step <= 2; ++step) {
            const DMat& A = step == 1 ? eig.V : eig.Vt;
            const double* Bm = step == 1 ? d.R : d.W;
            double* out = step == 1 ? d.W : d.beta;
            if (step == 1) {
                if (big && tall) launch_admm_gemm<160, 128, 5, 1>(d, A, Bm, out, num_sms, eig.lam.p, stream);
                else if (big) launch_admm_gemm<128, 128, 6, 1>(d, A, Bm, out, num_sms, eig.lam.p, stream);
                launch_admm_gemm<64, 64, 8, 1>(d, A, Bm, out, num_sms, eig.lam.p, stream);
            } else {
                if (big && tall) launch_admm_gemm<160, 128, 5, 2>(d, A, Bm, out, num_sms, eig.lam.p, stream);
                else if (big) launch_admm_gemm<128, 128, 6, 2>(d, A, Bm, out, num_sms, eig.lam.p, stream);
                launch_admm_gemm<64, 64, 8, 2>(d, A, Bm, out, num_sms, eig.lam.p, stream);
            }
        }
        const int is_update = (it % o.n_update == 0) ? 1 : 0;
        {
            ProfScope ps(PROF_ADMM_ELEM, stream);
            launch_update(is_update != 0, /*fuse rhs=*/!is_update);
            admm_finalize_kernel<<<1, 1024, 0, stream>>>(d, o, it, is_update);
            SCRC_LAUNCHED();
        }
        need_rhs = (is_update != 0);
        SCRC_CUDA(cudaGetLastError());
        ++sweeps;
        if ((it >= 4 && it % 2 == 0) || it >= o.max_iter) {
            SCRC_CUDA(cudaMemcpyAsync(h_scratch, d.n_active, 2 * sizeof(int), cudaMemcpyDeviceToHost, stream));
            SCRC_CUDA(cudaStreamSynchronize(stream));
            if (h_scratch[0] == 0) break;
            // cooperative cancellation (optim.cpp:307 polls the interrupt flag every iteration)
            if (cancel && cancel->load(std::memory_order_relaxed)) { b.aborted = true; break; }
        }
        ++it;
        if (it > o.max_iter + 1) break;  // safety net; finalize retires every problem at max_iter
    }
    {   // column-iterations actually executed -> algorithmic flops of the 2 GEMMs per sweep
        unsigned long long ci[2] = {0ull, 0ull};
        SCRC_CUDA(cudaMemcpyAsync(ci, d.col_iters, sizeof(ci), cudaMemcpyDeviceToHost, stream));
        SCRC_CUDA(cudaStreamSynchronize(stream));
        b.last_col_iters = ci[0];
        g_prof.add_work(PROF_GEMM_ADMM, 4.0 * (double)b.P * (double)b.P * (double)ci[0]);
        // elementwise sweep, algorithmic bytes per active column: plain sweeps read beta, zeta,
        // mult, X'y and write zeta, mult, rhs = 7 x 8 = 56 P; rho-adaptation sweeps read beta, zeta, mult + 4 snapshots
        // and write zeta, mult (which double as their snapshots) + the beta and mult_hat snapshots (88 P), and the rhs
        // needs its own pass once rho is known (3 reads + 1 write, 32 P) = 120 P
        g_prof.add_work(PROF_ADMM_ELEM, (56.0 * (double)(ci[0] - ci[1]) + 120.0 * (double)ci[1]) * (double)b.P);
    }
    return sweeps;
}

}  // namespace scrc
