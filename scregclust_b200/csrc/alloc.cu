#include "alloc.cuh"
#include "prof.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace scrc {

struct RvParams {
    CUtensorMap ta, tb;
    const double* Z; int64_t ldz;
    int ncells, T, K, L;
    int t_begin, t_end;  // target slice of this rank (tiles are counted from t_begin)
    const int* row_off;
    const int* kpad;
    const int* nsel;     // optional: rows of each block that are not zero padding
    const double* two_var;
    const double* inv_two_var;
    const double* half_log;
    int* votes;
    int CR, tiles_n, ctiles;
    const int* row_off_b;
    const int* col_index;
    const int* col_off;
    const int* col_cnt;
    double* sq_out;      // MODE 2: K x ncells x T squared test residuals of ONE fit (module fastest)
};

constexpr int RV_A_BYTES = RV_BM * GEMM_BK * 8;
constexpr int RV_B_BYTES = RV_BN * GEMM_BK * 8;
constexpr int RV_STAGE_BYTES = RV_A_BYTES + RV_B_BYTES;

__device__ __forceinline__ void consumer_bar() {   // the 8 consumer warps only
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

// MODE 1: per-cell votes; MODE 2: squared residuals written out (only used for the sequential network-prior sweep,
// which needs every gene's K x n2 slab).  The test-split sums of squares no longer come from this pass: like the
// training ones they are evaluated from the split's Gram matrices (sse_train_gram), O(s^2) per (module, target).
template <int MODE>
__global__ void __launch_bounds__(GEMM_THREADS, 2) resid_vote_kernel(const __grid_constant__ RvParams p) {
    constexpr bool VOTE = (MODE == 1);
    constexpr int MT = RV_BM / (GEMM_WARPS_M * 8);   // 4
    constexpr int NT = RV_BN / (GEMM_WARPS_N * 8);   // 2
    extern __shared__ __align__(1024) uint8_t smem[];   // see gemm_tn.cuh: keep the shared address space
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RV_STAGES * RV_STAGE_BYTES);
    uint64_t* empty = full + RV_STAGES;
    // score constants of this CTA's 32 target columns for every module, loaded once: RN(1 / (2 var)) and
    // -0.5 log(2 pi var); the slice counts of the modules (the per-module global loads they replace sat on the
    // critical path of every (cell tile, module) step: each was an L2 round trip in front of a branch)
    double* iv_s = reinterpret_cast<double*>(smem + RV_STAGES * RV_STAGE_BYTES + 2 * RV_STAGES * 8);   // [K][32]
    double* nhl_s = iv_s + (VOTE ? p.K * RV_BN : 0);                                                   // [K][32]
    int* votes_s = reinterpret_cast<int*>(nhl_s + (VOTE ? p.K * RV_BN : 0));                           // [32][K]
    int* kp_s = votes_s + (VOTE ? RV_BN * p.K : 0);                                                    // [K] padded rows
    int* s4_s = kp_s + p.K;                                                                            // [K] live rows, in 4s

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RV_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], GEMM_CONSUMER_WARPS); }
        mbar_fence_init();
    }
    __syncthreads();

    // One work item per CTA: (penalty/fit l, cell range cr, target tile tn).
    const int ct_per = (p.ctiles + p.CR - 1) / p.CR;
    const int w = blockIdx.x;
    const int tn = w % p.tiles_n, cr = (w / p.tiles_n) % p.CR, l = w / (p.tiles_n * p.CR);
    const int ct0 = cr * ct_per, ct1 = min(ct0 + ct_per, p.ctiles);
    const int ncol = p.col_cnt ? min(p.col_cnt[l], p.t_end) : p.t_end;  // one past the last column of this fit (slice)
    if (p.t_begin + tn * RV_BN >= ncol) return;           // whole CTA: nothing to do for this tile
    const int* cidx = p.col_index ? p.col_index + p.col_off[l] : nullptr;
    PipeState ps;

    if (warp < GEMM_PRODUCER_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 0 && lane == 0) {
            {
                for (int ct = ct0; ct < ct1; ++ct) {
                    for (int j = 0; j < p.K; ++j) {
                        const int kp = p.kpad[l * p.K + j];
                        const int ro = p.row_off[l * p.K + j];
                        const int rob = p.row_off_b ? p.row_off_b[l * p.K + j] : ro;
                        for (int kk = 0; kk < kp; kk += GEMM_BK) {
                            mbar_wait(&empty[ps.stage], ps.phase ^ 1);
                            uint8_t* sa = smem + ps.stage * RV_STAGE_BYTES;
                            mbar_expect_tx(&full[ps.stage], RV_STAGE_BYTES);
                            tma_load_2d(sa, &p.ta, ro + kk, ct * RV_BM, &full[ps.stage]);
                            tma_load_2d(sa + RV_A_BYTES, &p.tb, rob + kk, p.t_begin + tn * RV_BN, &full[ps.stage]);
                            ps.advance<RV_STAGES>();
                        }
                    }
                }
            }
        }
        return;
    }

    // Register pool of the CTA = 384 threads x 80 registers (2 CTAs/SM).  The producer warpgroup gives back
    // (80 - 24) x 128 = 7168 registers, which lets the 256 consumer threads grow by 28 -> 104 (multiple of 8).
    // Asking for more than the pool holds makes setmaxnreg.inc wait forever.
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int cw = warp - GEMM_PRODUCER_WARPS;
    const int ctid = threadIdx.x - GEMM_PRODUCER_WARPS * 32;
    const int warp_m = cw / GEMM_WARPS_N, warp_n = cw % GEMM_WARPS_N;
    const int a_row0 = warp_m * (MT * 8), b_col0 = warp_n * (NT * 8);
    const int r = lane >> 2, q = lane & 3;

    {
        const int t0 = p.t_begin + tn * RV_BN;
        for (int j = ctid; j < p.K; j += 256) {
            const int kp = p.kpad[l * p.K + j];
            kp_s[j] = kp;
            s4_s[j] = p.nsel ? ((p.nsel[l * p.K + j] + 3) & ~3) : kp;
        }
        if (VOTE) {
            for (int i = ctid; i < RV_BN * p.K; i += 256) {
                votes_s[i] = 0;
                const int j = i / RV_BN, lc = i % RV_BN;
                const int col = t0 + lc;
                double a = 1.0, b = 0.0;
                if (col < ncol) {
                    const int64_t vi = ((int64_t)l * p.K + j) * p.T + col;
                    a = p.inv_two_var[vi]; b = -p.half_log[vi];
                }
                iv_s[i] = a; nhl_s[i] = b;
            }
        }
        consumer_bar();
        for (int ct = ct0; ct < ct1; ++ct) {
            const int c0 = ct * RV_BM;
            double z[MT][NT][2], best[MT][NT][2];
            int bidx[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int row = c0 + a_row0 + i * 8 + r;
                        const int col = t0 + b_col0 + jj * 8 + q * 2 + e;
                        double zv = 0.0;
                        if (row < p.ncells && col < ncol) zv = p.Z[row + (int64_t)(cidx ? cidx[col] : col) * p.ldz];
                        z[i][jj][e] = zv;
                        best[i][jj][e] = -INFINITY;
                        bidx[i][jj][e] = 0;
                    }
            // (Issuing the mainloop of module j + 1 ahead of the epilogue of module j was tried and measured: no gain
            // -- 101.5 vs 99.8 ms at config 3, 749 vs 737 ms at config 4 -- the second accumulator set costs spills.)
            auto mainloop = [&](int j, double (&acc)[MT][NT][2]) {
#pragma unroll
                for (int i = 0; i < MT; ++i)
#pragma unroll
                    for (int jj = 0; jj < NT; ++jj) acc[i][jj][0] = acc[i][jj][1] = 0.0;
                if (j < p.K) {
                    // rows beyond the selection are zero padding up to the 16-row box: their 4-deep slices
                    // add exact zeros and are skipped
                    const int kp = kp_s[j], s4 = s4_s[j];
                    for (int kk = 0; kk < kp; kk += GEMM_BK) {
                        mbar_wait_warp(&full[ps.stage], ps.phase, lane);
                        const double* As = reinterpret_cast<const double*>(smem + ps.stage * RV_STAGE_BYTES);
                        const double* Bs = reinterpret_cast<const double*>(smem + ps.stage * RV_STAGE_BYTES + RV_A_BYTES);
                        // full stages take the branch-free path, which lets the fragment loads of the next 4-deep slice
                        // be scheduled above the DMMAs of the current one
                        const int ns = (s4 - kk) >> 2;
                        if (ns >= 4) mma_kslice<MT, NT>(As, Bs, a_row0, b_col0, lane, acc);
                        else mma_kslice_n<MT, NT>(As, Bs, a_row0, b_col0, lane, ns, acc);
                        stage_release_warp(&empty[ps.stage], lane);
                        ps.advance<RV_STAGES>();
                    }
                }
            };
            // score constants of a module's columns (requested before the next mainloop hides their latency)
            auto load_consts = [&](int j, double (&iv)[NT][2], double (&nhl)[NT][2]) {
#pragma unroll
                for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        nhl[jj][e] = 0.0; iv[jj][e] = 1.0;
                        if (VOTE && j < p.K) {
                            const int lc = b_col0 + jj * 8 + q * 2 + e;
                            iv[jj][e] = iv_s[j * RV_BN + lc];
                            nhl[jj][e] = nhl_s[j * RV_BN + lc];
                        }
                    }
            };
            auto epilogue = [&](int j, double (&acc)[MT][NT][2], const double (&iv)[NT][2], const double (&nhl)[NT][2]) {
                // residual and squared residual (scregclust.R:1593-1599)
#pragma unroll
                for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const double res = z[i][jj][e] - acc[i][jj][e];
                            acc[i][jj][e] = res * res;
                        }
                if (MODE == 2 && j < p.K) {
#pragma unroll
                    for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int i = 0; i < MT; ++i) {
                                const int row = c0 + a_row0 + i * 8 + r;
                                const int col = t0 + b_col0 + jj * 8 + q * 2 + e;
                                if (row < p.ncells && col < ncol)
                                    p.sq_out[j + (int64_t)p.K * (row + (int64_t)p.ncells * col)] = acc[i][jj][e];
                            }
                }
                // per-cell score -sq / (2 var) - 0.5 log(2 pi var) and running first maximum (allocation.cpp:59-62,80-85),
                // evaluated as one fused multiply-add with the correctly rounded reciprocal RN(1 / (2 var)) and
                // -0.5 log(2 pi var) from make_resid_var.  It differs from the reference's quotient-then-subtract by
                // at most ~1.5 ulp of the score -- four orders of magnitude below the rounding noise the squared
                // residual itself carries (its bits depend on the summation order of whatever BLAS / tensor kernel
                // produced the prediction) and below the 1e-12 band in which the oracle calls a cell's vote
                // noise-decided.  Structurally identical modules still tie bit for bit (same inputs, same formula),
                // so first-index tie breaking is preserved.
                if (VOTE && j < p.K) {
#pragma unroll
                    for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
#pragma unroll
                            for (int i = 0; i < MT; ++i) {
                                const double sc = fma(-acc[i][jj][e], iv[jj][e], nhl[jj][e]);
                                if (sc > best[i][jj][e]) { best[i][jj][e] = sc; bidx[i][jj][e] = j; }
                            }
                }
            };
            {
                double accA[MT][NT][2];
                double iv[NT][2], nhl[NT][2];
                for (int j = 0; j < p.K; ++j) {
                    load_consts(j, iv, nhl);
                    mainloop(j, accA);
                    epilogue(j, accA, iv, nhl);
                }
            }
            if (VOTE) {
                // The 8 lanes that share a target column mostly vote for the same module: lanes with the
                // same (column, module) key elect a leader that adds the group's count once.
#pragma unroll
                for (int jj = 0; jj < NT; ++jj)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int lc = b_col0 + jj * 8 + q * 2 + e;
                        const bool cvalid = t0 + lc < ncol;
                        const int b0 = bidx[0][jj][e];
                        int cnt = 0;
#pragma unroll
                        for (int i = 0; i < MT; ++i) {
                            const bool valid = cvalid && (c0 + a_row0 + i * 8 + r < p.ncells);
                            if (valid && bidx[i][jj][e] == b0) ++cnt;
                            else if (valid) atomicAdd(&votes_s[lc * p.K + bidx[i][jj][e]], 1);
                        }
                        const unsigned peers = __match_any_sync(0xffffffffu, (q << 8) | b0);
                        const int total = __reduce_add_sync(peers, cnt);
                        if (lane == __ffs(peers) - 1 && total > 0) atomicAdd(&votes_s[lc * p.K + b0], total);
                    }
            }
        }
        if (VOTE) {
            consumer_bar();
            for (int i = ctid; i < RV_BN * p.K; i += 256) {
                const int lc = i / p.K, j = i % p.K;
                const int v = votes_s[i];
                if (v != 0 && t0 + lc < ncol) atomicAdd(&p.votes[((int64_t)l * p.T + t0 + lc) * p.K + j], v);
            }
        }
    }
}

int resid_vote_pick_cr(int ncells, int T, int L, int num_sms) {
    // Cell ranges per (fit, target tile); T = width of the target slice the launch covers.  The kernel only adds
    // integer votes, so the split does not affect any result.  16 cell tiles (1024 cells) per CTA amortise the
    // per-CTA staging of the score constants; launches with few fits or a narrow slice (late cycles, one rank of
    // several) split finer -- down to 4 cell tiles per CTA -- so that the grid keeps about 6 waves of 2 CTAs per SM.
    const int64_t ctiles = ceil_div(ncells, RV_BM), tiles_n = std::max<int64_t>(1, ceil_div(T, RV_BN));
    int per = 16;
    while (per > 4 && (int64_t)L * tiles_n * ceil_div(ctiles, per) < (int64_t)12 * num_sms) per >>= 1;
    return std::max(1, (int)ceil_div(ctiles, per));
}
void resid_vote(const ResidVoteArgs& a, int mode, int num_sms, cudaStream_t st) {
    if (a.K + 1 > RV_MAX_K + 1) throw Error(-60, "resid_vote: at most 64 modules are supported");
    RvParams p;
    // operands are allocated with at least one 16-row block so that the maps are never empty
    p.ta = make_tmap_f64(a.ZselT, std::max<int64_t>(a.rows_total_a >= 0 ? a.rows_total_a : a.rows_total, 16), a.ncells,
                         a.ldzs, RV_BM);
    p.tb = make_tmap_f64(a.beta, std::max<int64_t>(a.rows_total, 16), a.T, a.ldb, RV_BN);
    p.Z = a.Z; p.ldz = a.ldz; p.ncells = a.ncells; p.T = a.T; p.K = a.K; p.L = a.L;
    p.row_off = a.row_off; p.kpad = a.kpad; p.nsel = a.nsel; p.two_var = a.two_var; p.inv_two_var = a.inv_two_var; p.half_log = a.half_log;
    p.votes = a.votes; p.CR = a.CR;
    p.row_off_b = a.row_off_b; p.col_index = a.col_index; p.col_off = a.col_off; p.col_cnt = a.col_cnt;
    p.sq_out = a.sq_out;
    if (mode == 1 && !a.inv_two_var) throw Error(-63, "resid_vote: vote mode needs the reciprocal table");
    if (mode == 2 && (a.L != 1 || !a.sq_out)) throw Error(-62, "resid_vote: store mode handles one fit per launch");
    p.t_begin = a.t_begin; p.t_end = a.t_end < 0 ? a.T : a.t_end;
    if (p.t_begin < 0 || p.t_end > a.T || p.t_begin > p.t_end) throw Error(-64, "resid_vote: target slice out of range");
    if (p.t_begin == p.t_end) return;
    p.tiles_n = (int)ceil_div(p.t_end - p.t_begin, RV_BN); p.ctiles = (int)ceil_div(a.ncells, RV_BM);

    const int smem = RV_STAGES * RV_STAGE_BYTES + 2 * RV_STAGES * 8 +
                     (mode == 1 ? 2 * a.K * RV_BN * 8 + RV_BN * a.K * 4 : 0) + 2 * a.K * 4;
    if (mode != 1 && mode != 2) throw Error(-65, "resid_vote: mode must be 1 (votes) or 2 (store)");
    static SmemAttr attr[3];
    if (mode == 1) attr[1].ensure(resid_vote_kernel<1>, smem);
    else attr[2].ensure(resid_vote_kernel<2>, smem);
    const int64_t nwork = (int64_t)a.L * a.CR * p.tiles_n;
    if (nwork > 0x7fffffffLL) throw Error(-61, "resid_vote: too many work items");
    const int grid = (int)nwork;
    (void)num_sms;
    {
        // algorithmic bytes: the target tile is read once per (fit, target tile, cell tile); the
        // gathered regulators and coefficients once per module slice
        const double Ts = (double)(p.t_end - p.t_begin);
        const double bytes = 8.0 * (double)a.L * a.ncells * Ts + 8.0 * (double)a.rows_total * (a.ncells * (double)p.tiles_n + Ts * p.ctiles);
        ProfScope ps(PROF_RESID_VOTE, st, bytes);
        if (mode == 1) resid_vote_kernel<1><<<grid, GEMM_THREADS, smem, st>>>(p);
        else resid_vote_kernel<2><<<grid, GEMM_THREADS, smem, st>>>(p);
    }
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// ---------------------------------------------------------------- Gram-form sums of squares / finishing kernels
// One CTA per (fit*module, 128-target chunk); the module's Gram block lives in shared memory.
constexpr int SSG_MAX_S = 72;   // 72^2 doubles = 41 KB of static shared memory
constexpr int SSG_TILE = 64;    // larger selections: 64 x 32 tiles of the Gram block
constexpr int SSG_TILE_C = 32;
__global__ void __launch_bounds__(128) sse_train_gram_kernel(const double* __restrict__ G_rr, int64_t ldrr,
                                                             const double* __restrict__ G_rt, int64_t ldrt,
                                                             const double* __restrict__ gtt,
                                                             const double* __restrict__ beta, int64_t ldb,
                                                             const int* __restrict__ sel, const int* __restrict__ row_off,
                                                             const int* __restrict__ nsel, int T, int K, double* sse,
                                                             int t0, int t1, const int* __restrict__ col_index,
                                                             const int* __restrict__ col_off,
                                                             const int* __restrict__ col_cnt, int pseudo,
                                                             const int* __restrict__ sel_row_off) {
    __shared__ double s_g[SSG_MAX_S * SSG_MAX_S];
    __shared__ int s_sel[SSG_TILE + SSG_TILE_C > SSG_MAX_S ? SSG_TILE + SSG_TILE_C : SSG_MAX_S];
    // pseudo: one extra slot per fit (j == K) that holds y'y, the "no model" sum of squares (scregclust.R:1505-1515)
    const int KK = K + (pseudo ? 1 : 0);
    const int l = blockIdx.y / KK, j = blockIdx.y % KK;
    const int fk = l * K + j;
    const int s = j < K ? nsel[fk] : 0;
    const int ro = j < K ? row_off[fk] : 0;
    const int so = j < K ? (sel_row_off ? sel_row_off[fk] : ro) : 0;   // rows of `sel` (may differ from the rows of beta)
    const int t = t0 + blockIdx.x * 128 + threadIdx.x;
    // quick mode: column t of this fit is target col_index[col_off[l] + t] (per-fit target subsets)
    const bool valid = t < t1 && !(col_cnt && t >= col_cnt[l]);
    const int tg = valid ? (col_index ? col_index[col_off[l] + t] : t) : 0;
    const double* b = beta + ro + (int64_t)(valid ? t : t0) * ldb;
    double cross = 0.0, quad = 0.0;
    if (s <= SSG_MAX_S) {
        // the module's whole Gram block in shared memory
        for (int a = threadIdx.x; a < s; a += 128) s_sel[a] = sel[so + a];
        __syncthreads();
        for (int e = threadIdx.x; e < s * s; e += 128) s_g[e] = G_rr[s_sel[e % s] + (int64_t)s_sel[e / s] * ldrr];
        __syncthreads();
        if (valid)
            for (int a = 0; a < s; ++a) {
                const double ba = b[a];
                cross += ba * G_rt[s_sel[a] + (int64_t)tg * ldrt];
                double row = 0.0;
                for (int c = 0; c < s; ++c) row += s_g[a + c * s] * b[c];
                quad += ba * row;
            }
    } else {
        // large selections (weak penalization, or the ridge-initialised n1 < P regime: hundreds of regulators per
        // module): the Gram block goes through shared memory in 64 x 32 tiles (zero padded); a thread keeps the 32
        // coefficients of its target that the column tile needs in registers, so the inner product runs on one
        // broadcast shared-memory operand per FMA
        int* s_sa = s_sel;
        int* s_sc = s_sel + SSG_TILE;
        for (int c0 = 0; c0 < s; c0 += SSG_TILE_C) {
            const int nc = min(SSG_TILE_C, s - c0);
            double bc[SSG_TILE_C];
#pragma unroll
            for (int c = 0; c < SSG_TILE_C; ++c) bc[c] = (valid && c < nc) ? b[c0 + c] : 0.0;
            for (int a0 = 0; a0 < s; a0 += SSG_TILE) {
                const int na = min(SSG_TILE, s - a0);
                __syncthreads();                                   // the previous tile has been consumed
                if (threadIdx.x < SSG_TILE) s_sa[threadIdx.x] = threadIdx.x < na ? sel[so + a0 + threadIdx.x] : -1;
                else if (threadIdx.x - SSG_TILE < SSG_TILE_C)
                    s_sc[threadIdx.x - SSG_TILE] = (int)threadIdx.x - SSG_TILE < nc ? sel[so + c0 + threadIdx.x - SSG_TILE] : -1;
                __syncthreads();
                for (int e = threadIdx.x; e < SSG_TILE * SSG_TILE_C; e += 128) {
                    const int ia = s_sa[e % SSG_TILE], ic = s_sc[e / SSG_TILE];
                    s_g[e] = (ia >= 0 && ic >= 0) ? G_rr[ia + (int64_t)ic * ldrr] : 0.0;
                }
                __syncthreads();
                if (valid) {
                    for (int a = 0; a < na; ++a) {
                        double row = 0.0;
#pragma unroll
                        for (int c = 0; c < SSG_TILE_C; ++c) row += s_g[a + c * SSG_TILE] * bc[c];
                        const double ba = b[a0 + a];
                        quad += ba * row;
                        if (c0 == 0) cross += ba * G_rt[s_sa[a] + (int64_t)tg * ldrt];
                    }
                }
            }
        }
    }
    if (!valid) return;
    double acc = gtt[tg];
    if (s > 0) acc = (acc - 2.0 * cross) + quad;
    sse[((int64_t)l * (K + 1) + j) * T + t] = acc;
}
void sse_train_gram(const double* G_rr, int64_t ldrr, const double* G_rt, int64_t ldrt, const double* gtt,
                    const double* beta, int64_t ldb, const int* sel, const int* row_off, const int* nsel, int T, int K,
                    int L, double* sse, cudaStream_t st, int t0, int t1, const int* col_index, const int* col_off,
                    const int* col_cnt, bool pseudo, const int* sel_row_off) {
    if (t1 < 0) t1 = T;
    if (t1 == t0) return;
    const int64_t gy = (int64_t)L * (K + (pseudo ? 1 : 0));
    if (gy > 65535) {      // gridDim.y limit: the leave-one-out batch of the importance step can exceed it
        const int chunk = 65535 / (K + (pseudo ? 1 : 0));
        for (int l0 = 0; l0 < L; l0 += chunk) {
            const int nl = std::min(chunk, L - l0);
            sse_train_gram(G_rr, ldrr, G_rt, ldrt, gtt, beta, ldb, sel, row_off + (size_t)l0 * K, nsel + (size_t)l0 * K, T, K, nl,
                           sse + (size_t)l0 * (K + 1) * T, st, t0, t1, col_index, col_off ? col_off + l0 : nullptr,
                           col_cnt ? col_cnt + l0 : nullptr, pseudo, sel_row_off ? sel_row_off + (size_t)l0 * K : nullptr);
        }
        return;
    }
    dim3 grid((unsigned)ceil_div(t1 - t0, 128), (unsigned)gy);
    ProfScope ps(PROF_RESID_VOTE, st, 0.0);
    sse_train_gram_kernel<<<grid, 128, 0, st>>>(G_rr, ldrr, G_rt, ldrt, gtt, beta, ldb, sel, row_off, nsel, T, K, sse, t0, t1,
                                                col_index, col_off, col_cnt, pseudo ? 1 : 0, sel_row_off);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

__global__ void resid_var_kernel(const double* __restrict__ sse_train, const int* __restrict__ nsel, int n1, int T,
                                 int K, int L, double* var, double* two_var, double* inv_two_var, double* half_log, int t0,
                                 int Ts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over L*K*Ts
    if (i >= (int64_t)L * K * Ts) return;
    const int t = t0 + (int)(i % Ts);
    const int j = (int)((i / Ts) % K);
    const int l = (int)(i / ((int64_t)Ts * K));
    const double sse = sse_train[((int64_t)l * (K + 1) + j) * T + t];
    const double v = sse / (double)(n1 - nsel[l * K + j]);             // scregclust.R:1636-1638
    const int64_t o = ((int64_t)l * K + j) * T + t;
    var[o] = v;
    two_var[o] = 2.0 * v;                                               // allocation.cpp:60
    if (inv_two_var) inv_two_var[o] = 1.0 / (2.0 * v);                  // correctly rounded (IEEE division)
    half_log[o] = 0.5 * log((2.0 * M_PI) * v);                          // allocation.cpp:62
}
void make_resid_var(const double* sse_train, const int* nsel, int n1, int T, int K, int L, double* var,
                    double* two_var, double* inv_two_var, double* half_log, cudaStream_t st, int t0, int t1) {
    if (t1 < 0) t1 = T;
    const int64_t n = (int64_t)L * K * (t1 - t0);
    if (n == 0) return;
    resid_var_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(sse_train, nsel, n1, T, K, L, var, two_var,
                                                                 inv_two_var, half_log, t0, t1 - t0);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

__global__ void finish_alloc_kernel(const double* __restrict__ sse_test, const int* __restrict__ votes, int T, int K,
                                    int L, double* r2, double* best_r2, int* best_idx, int* k_new, int t0, int Ts) {
    const int64_t ii = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // over L*Ts
    if (ii >= (int64_t)L * Ts) return;
    const int t = t0 + (int)(ii % Ts), l = (int)(ii / Ts);
    const int64_t i = (int64_t)l * T + t;
    const double ss0 = sse_test[((int64_t)l * (K + 1) + K) * T + t];
    double br = -INFINITY; int bi = 0;
    int bv = -1, bk = 0;
    for (int j = 0; j < K; ++j) {
        const double v = 1.0 - sse_test[((int64_t)l * (K + 1) + j) * T + t] / ss0;   // scregclust.R:1629-1631
        r2[((int64_t)l * K + j) * T + t] = v;
        if (v > br) { br = v; bi = j; }
        if (votes) {
            const int vt = votes[((int64_t)l * T + t) * K + j];
            if (vt > bv) { bv = vt; bk = j; }                                         // allocation.cpp:88-89
        }
    }
    best_r2[i] = br;
    best_idx[i] = bi + 1;
    if (votes) k_new[i] = bk + 1;
}
void finish_alloc(const double* sse_test, const int* votes, int T, int K, int L, double* r2, double* best_r2,
                  int* best_idx, int* k_new, cudaStream_t st, int t0, int t1) {
    if (t1 < 0) t1 = T;
    const int64_t n = (int64_t)L * (t1 - t0);
    if (n == 0) return;
    finish_alloc_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(sse_test, votes, T, K, L, r2, best_r2, best_idx, k_new,
                                                                    t0, t1 - t0);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// ---------------------------------------------------------------- gather + transpose of selected regulators
__global__ void gather_transpose_kernel(const double* __restrict__ Zr, int64_t ldz, int ncells,
                                        const int* __restrict__ row_src, int64_t rows, double* __restrict__ out,
                                        int64_t ldo) {
    __shared__ double tile[32][33];
    const int64_t r0 = (int64_t)blockIdx.x * 32;   // stacked rows
    const int64_t c0 = (int64_t)blockIdx.y * 32;   // cells
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t rr = r0 + j, c = c0 + threadIdx.x;
        double v = 0.0;
        if (rr < rows && c < ncells) {
            const int src = row_src[rr];
            if (src >= 0) v = Zr[c + (int64_t)src * ldz];
        }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t rr = r0 + threadIdx.x, c = c0 + j;
        if (rr < rows && c < ncells) out[rr + c * ldo] = tile[threadIdx.x][j];
    }
}
void gather_transpose(const double* Zr, int64_t ldz, int ncells, const int* row_src, int64_t rows, double* out,
                      int64_t ldo, cudaStream_t st) {
    if (rows == 0 || ncells == 0) return;
    dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(ncells, 32));
    gather_transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(Zr, ldz, ncells, row_src, rows, out, ldo);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

}  // namespace scrc
