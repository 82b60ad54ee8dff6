// FP64 "TN" tensor-core GEMM core for sm_100a:  C[m,n] = sum_k A[k,m] * B[k,n]
//
// Both operands are column-major with the contraction index k contiguous
// (A is K x M with leading dimension lda, B is K x N), which is the natural
// shape of every contraction on the scregclust hot path:
//   * Gram matrices X'X, X'Y                    (src/optim.cpp:17-27,99,414,421)
//   * fast_cor cross products                   (R/utils.R:538)
//   * eigenbasis ADMM solves  V'(rhs), Vt'(W)   (replaces optim.cpp:136-143)
//   * NNLS predictions  ZselT' * beta_hat       (R/scregclust.R:1590-1595)
//
// Structure (one persistent CTA per SM):
//   warps 0..3  : producer warpgroup (registers donated via setmaxnreg). One
//                 elected lane streams 16-deep k-slices of the
//                 A and B tiles into a STAGES-deep shared-memory ring with
//                 cp.async.bulk.tensor (128-byte swizzle), signalling `full`
//                 mbarriers with complete_tx.
//   warps 4..11 : consumers, arranged 2 (M) x 4 (N). Each k-slice is consumed
//                 with mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4); one lane per warp
//                 arrives on the `empty` mbarrier when the warp is done reading.
// Accumulators live in registers (tcgen05/UMMA has no FP64 kind, so TMEM is not
// an option for this data type).  The epilogue is a policy hook.
#pragma once
#include "common.cuh"

namespace scrc {

constexpr int GEMM_BK = 16;           // doubles per k-slice = one 128-byte swizzle row
constexpr int GEMM_CONSUMER_WARPS = 8;
constexpr int GEMM_PRODUCER_WARPS = 4;  // one warpgroup; only lane 0 of its first warp issues TMA
constexpr int GEMM_THREADS = (GEMM_CONSUMER_WARPS + GEMM_PRODUCER_WARPS) * 32;
constexpr int GEMM_WARPS_M = 2;
constexpr int GEMM_WARPS_N = 4;

struct TileInfo {
    const CUtensorMap* ta;
    const CUtensorMap* tb;
    int m0, n0;      // tile origin in C
    int k0, k1;      // contraction range [k0, k1)
    int aux0, aux1;  // policy-defined
};

template <int BM, int BN, int STAGES>
struct GemmSmem {
    static constexpr int A_BYTES = BM * GEMM_BK * 8;
    static constexpr int B_BYTES = BN * GEMM_BK * 8;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 2 * STAGES * 8;
};

struct PipeState {
    uint32_t stage = 0;
    uint32_t phase = 0;
    template <int STAGES>
    __device__ __forceinline__ void advance() {
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
};

// Consume one k-slice (16 deep) of the staged tiles.
template <int MT, int NT>
__device__ __forceinline__ void mma_kslice(const double* __restrict__ As,
                                           const double* __restrict__ Bs, int a_row0, int b_col0,
                                           int lane, double (&acc)[MT][NT][2]) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        // logical k = kk*4 + q -> 16-byte chunk (k>>1), swizzled with the row index (row & 7 == r)
        const int off = ((((kk * 2) + (q >> 1)) ^ r) << 1) + (q & 1);
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = As[(a_row0 + i * 8 + r) * GEMM_BK + off];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = Bs[(b_col0 + j * 8 + r) * GEMM_BK + off];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// Same, but only the first `nslices` (1..4) 4-deep slices of the stage: the rest is known to be zero padding.
template <int MT, int NT>
__device__ __forceinline__ void mma_kslice_n(const double* __restrict__ As,
                                             const double* __restrict__ Bs, int a_row0, int b_col0,
                                             int lane, int nslices, double (&acc)[MT][NT][2]) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (kk >= nslices) break;
        const int off = ((((kk * 2) + (q >> 1)) ^ r) << 1) + (q & 1);
        double a[MT], b[NT];
#pragma unroll
        for (int i = 0; i < MT; ++i) a[i] = As[(a_row0 + i * 8 + r) * GEMM_BK + off];
#pragma unroll
        for (int j = 0; j < NT; ++j) b[j] = Bs[(b_col0 + j * 8 + r) * GEMM_BK + off];
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int j = 0; j < NT; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
}

// Policy interface:
//   struct Params;                               (trivially copyable, < 4 KB)
//   static constexpr int BM, BN, STAGES;
//   __device__ static int  num_tiles(const Params&);
//   __device__ static bool tile(const Params&, int t, TileInfo&);   // false => skip tile
//   __device__ static void epilogue(const Params&, const TileInfo&, acc, warp_m, warp_n, lane);
template <class Policy>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tn_kernel(const __grid_constant__ typename Policy::Params prm) {
    constexpr int BM = Policy::BM, BN = Policy::BN, STAGES = Policy::STAGES;
    constexpr int MT = BM / (GEMM_WARPS_M * 8), NT = BN / (GEMM_WARPS_N * 8);
    using SM = GemmSmem<BM, BN, STAGES>;

    // 1024-byte alignment is required by the 128-byte TMA swizzle.  The array is addressed with
    // plain offsets only (no integer re-alignment of the pointer) so that the compiler keeps the
    // shared address space and emits LDS for the fragment loads: generic LD.E loads are not ordered
    // with the mbarrier arrive that hands the stage back to the TMA producer.
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + SM::BAR_OFFSET);
    uint64_t* empty = full + STAGES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], GEMM_CONSUMER_WARPS);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int ntiles = Policy::num_tiles(prm);
    PipeState ps;

    if (warp < GEMM_PRODUCER_WARPS) {
        // ------------------------------ TMA producer ------------------------------
        // The producer warpgroup hands its registers to the consumers (setmaxnreg).
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == 0 && lane == 0) {
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                TileInfo ti;
                if (!Policy::tile(prm, t, ti)) continue;
                for (int k = ti.k0; k < ti.k1; k += GEMM_BK) {
                    mbar_wait(&empty[ps.stage], ps.phase ^ 1);
                    uint8_t* sa = smem + ps.stage * SM::STAGE_BYTES;
                    mbar_expect_tx(&full[ps.stage], SM::STAGE_BYTES);
                    tma_load_2d(sa, ti.ta, k, ti.m0, &full[ps.stage]);
                    tma_load_2d(sa + SM::A_BYTES, ti.tb, k, ti.n0, &full[ps.stage]);
                    ps.template advance<STAGES>();
                }
            }
        }
    } else {
        // ------------------------------ DMMA consumers -----------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int cw = warp - GEMM_PRODUCER_WARPS;
        const int warp_m = cw / GEMM_WARPS_N, warp_n = cw % GEMM_WARPS_N;
        const int a_row0 = warp_m * (MT * 8), b_col0 = warp_n * (NT * 8);
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            TileInfo ti;
            if (!Policy::tile(prm, t, ti)) continue;
            double acc[MT][NT][2];
#pragma unroll
            for (int i = 0; i < MT; ++i)
#pragma unroll
                for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
            for (int k = ti.k0; k < ti.k1; k += GEMM_BK) {
                mbar_wait_warp(&full[ps.stage], ps.phase, lane);
                const double* As = reinterpret_cast<const double*>(smem + ps.stage * SM::STAGE_BYTES);
                const double* Bs = reinterpret_cast<const double*>(smem + ps.stage * SM::STAGE_BYTES +
                                                                   SM::A_BYTES);
                mma_kslice<MT, NT>(As, Bs, a_row0, b_col0, lane, acc);
                stage_release_warp(&empty[ps.stage], lane);
                ps.template advance<STAGES>();
            }
            Policy::epilogue(prm, ti, acc, warp_m, warp_n, lane);
        }
    }
}

// Iterate the accumulator fragment of one thread: f(row_in_tile, col_in_tile, v0, v1)
// where v0 -> (row, col), v1 -> (row, col + 1).
template <int MT, int NT, class F>
__device__ __forceinline__ void for_each_acc(double (&acc)[MT][NT][2], int warp_m, int warp_n,
                                             int lane, F&& f) {
    const int r = lane >> 2, q = lane & 3;
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int i = 0; i < MT; ++i)
            f(warp_m * (MT * 8) + i * 8 + r, warp_n * (NT * 8) + j * 8 + q * 2, acc[i][j][0],
              acc[i][j][1]);
}

// ----------------------------------------------------------------------------
// Generic policy: dense M x N output with a small set of fused epilogues.
// ----------------------------------------------------------------------------
enum EpiMode : int {
    EPI_PLAIN = 0,      // C = alpha * acc
    EPI_SYM_LOWER = 1,  // A == B: only tiles with m-tile >= n-tile; mirror (exact symmetry)
    EPI_COR = 2,        // C = clamp(acc / sqrt(rvec[row] * cvec[col]), -1, 1)   (R/utils.R:538-540)
    EPI_ROWDIV = 3,     // C = acc / d, d = alpha*rvec[row] + shift;  rows with |d| <= cut give 0
    EPI_SUBFROM = 4,    // C = D - acc   (D has the layout of C; may alias C)
};

template <int BM_, int BN_, int STAGES_, int MODE>
struct DensePolicy {
    static constexpr int BM = BM_, BN = BN_, STAGES = STAGES_;
    struct Params {
        CUtensorMap ta, tb;
        double* C;
        int64_t ldc;
        int M, N, K;
        double alpha;
        const double* rvec;
        const double* cvec;
        const double* D;
        int64_t ldd;
        double shift, cut;
    };
    __device__ static int tiles_m(const Params& p) { return (p.M + BM - 1) / BM; }
    __device__ static int tiles_n(const Params& p) { return (p.N + BN - 1) / BN; }
    __device__ static int num_tiles(const Params& p) { return tiles_m(p) * tiles_n(p); }
    __device__ static bool tile(const Params& p, int t, TileInfo& ti) {
        const int tm = t % tiles_m(p), tn = t / tiles_m(p);
        if (MODE == EPI_SYM_LOWER && tm * BM + BM <= tn * BN) return false;  // strictly above diag
        ti.ta = &p.ta; ti.tb = &p.tb;
        ti.m0 = tm * BM; ti.n0 = tn * BN;
        ti.k0 = 0; ti.k1 = p.K;
        return true;
    }
    template <int MT, int NT>
    __device__ static void epilogue(const Params& p, const TileInfo& ti, double (&acc)[MT][NT][2],
                                    int warp_m, int warp_n, int lane) {
        for_each_acc<MT, NT>(acc, warp_m, warp_n, lane, [&](int rr, int cc, double v0, double v1) {
            const int row = ti.m0 + rr;
            if (row >= p.M) return;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int col = ti.n0 + cc + e;
                if (col >= p.N) continue;
                double v = e ? v1 : v0;
                if (MODE == EPI_PLAIN) {
                    p.C[row + (int64_t)col * p.ldc] = p.alpha * v;
                } else if (MODE == EPI_SYM_LOWER) {
                    if (row >= col) {
                        v *= p.alpha;
                        p.C[row + (int64_t)col * p.ldc] = v;
                        p.C[col + (int64_t)row * p.ldc] = v;
                    }
                } else if (MODE == EPI_COR) {
                    v = v / sqrt(p.rvec[row] * p.cvec[col]);
                    v = fmax(fmin(v, 1.0), -1.0);
                    p.C[row + (int64_t)col * p.ldc] = v;
                } else if (MODE == EPI_ROWDIV) {
                    const double d = p.alpha * p.rvec[row] + p.shift;
                    p.C[row + (int64_t)col * p.ldc] = (fabs(d) > p.cut) ? v / d : 0.0;
                } else if (MODE == EPI_SUBFROM) {
                    p.C[row + (int64_t)col * p.ldc] = p.D[row + (int64_t)col * p.ldd] - v;
                }
            }
        });
    }
};

// Host-side launcher for DensePolicy GEMMs.  A: K x M (lda), B: K x N (ldb), C: M x N (ldc).
struct DenseArgs {
    const double* A; int64_t lda;
    const double* B; int64_t ldb;
    double* C; int64_t ldc;
    int M, N, K;
    double alpha = 1.0;
    const double* rvec = nullptr;
    const double* cvec = nullptr;
    const double* D = nullptr; int64_t ldd = 0;
    double shift = 0.0, cut = 0.0;
};
void gemm_tn_dense(int mode, const DenseArgs& a, int num_sms, cudaStream_t stream);

}  // namespace scrc
