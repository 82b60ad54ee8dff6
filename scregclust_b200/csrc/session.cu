#include "session.cuh"
#include "shard.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <numeric>

namespace scrc {

// SCRC_DEBUG_TIMING=1: host wall time per phase of a cycle (with a stream synchronisation at every phase
// boundary, so the phases do not overlap; totals are printed when the session is destroyed).
struct PhaseTimer {
    bool on;
    cudaStream_t st;
    double t_prev;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
    PhaseTimer(cudaStream_t s) : on(getenv("SCRC_DEBUG_TIMING") != nullptr), st(s), t_prev(0.0) { if (on) t_prev = now(); }
    void mark(double* acc, int idx) {
        if (!on) return;
        cudaStreamSynchronize(st);
        const double t = now();
        acc[idx] += t - t_prev;
        t_prev = t;
    }
};
static const char* kPhaseNames[] = {"tables(host)", "admm", "merge1+extract+sync", "nnls tables", "nnls", "resid/vote",
                                    "merge2", "outputs", "importance tables", "importance nnls", "importance resid",
                                    "importance host", "coeff pack"};
void Ctx::print_phase_times() const {
    if (getenv("SCRC_DEBUG_TIMING") == nullptr) return;
    double tot = 0.0;
    for (int i = 0; i < 13; ++i) tot += phase_s[i];
    fprintf(stderr, "[scrc rank %d] host wall per phase (synchronised), total %.1f ms:", rank(), tot * 1e3);
    for (int i = 0; i < 13; ++i) if (phase_s[i] > 0.0) fprintf(stderr, "  %s %.1f", kPhaseNames[i], phase_s[i] * 1e3);
    fprintf(stderr, "\n");
}

// ============================================================================ device / helpers
void Device::init(int device) {
    id = device;
    SCRC_CUDA(cudaSetDevice(id));
    cudaDeviceProp prop;
    SCRC_CUDA(cudaGetDeviceProperties(&prop, id));
    // the library is built for sm_100a only (arch-specific code, not forward compatible to sm_103 / sm_110 / sm_12x)
    if (!(prop.major == 10 && prop.minor == 0))
        throw Error(SCRC_ERR_UNSUPPORTED, std::string("libscregclust_b200 is built for sm_100a (B200) only; device ") +
                                              std::to_string(device) + " is sm_" + std::to_string(prop.major) +
                                              std::to_string(prop.minor));
    num_sms = prop.multiProcessorCount;
    if (!stream) SCRC_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    {   // keep freed work buffers in the device's pool instead of returning them to the driver (common.cuh: DevBuf)
        cudaMemPool_t pool;
        SCRC_CUDA(cudaDeviceGetDefaultMemPool(&pool, id));
        unsigned long long keep = ~0ull;
        SCRC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    if (!h_pinned) SCRC_CUDA(cudaMallocHost(&h_pinned, 64 * sizeof(int)));
}
void Device::destroy() {
    if (stream) { cudaStreamDestroy(stream); stream = nullptr; }
    if (h_pinned) { cudaFreeHost(h_pinned); h_pinned = nullptr; }
}

__global__ void scale_copy_kernel(const double* in, int64_t ldi, double* out, int64_t ldo, int64_t rows, int64_t cols,
                                  double scale) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i % rows, c = i / rows;
    out[r + c * ldo] = in[r + c * ldi] * scale;
}

void build_eig(const DMat& G, double scale, EigBasis& eig, Device& dev) {
    const int P = (int)G.rows;
    eig.P = P;
    eig.V.alloc(P, P);
    eig.Vt.alloc(P, P);
    eig.lam.alloc(P);
    SCRC_CUDA(cudaMemsetAsync(eig.V.data(), 0, eig.V.bytes(), dev.stream));
    const int64_t n = (int64_t)P * P;
    scale_copy_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, dev.stream>>>(G.data(), G.ld, eig.V.data(), eig.V.ld, P, P, scale);
    SCRC_LAUNCHED();
    sym_eig(eig.V.data(), eig.V.ld, P, eig.lam.p, dev.stream);
    SCRC_CUDA(cudaMemsetAsync(eig.Vt.data(), 0, eig.Vt.bytes(), dev.stream));
    transpose(eig.V.data(), eig.V.ld, P, P, eig.Vt.data(), eig.Vt.ld, dev.stream);
}

void eig_solve(const EigBasis& eig, double mult, double shift, double cut, const DMat& B, DMat& out, DMat& tmp,
               Device& dev) {
    const int P = eig.P;
    tmp.reshape(P, B.cols);
    out.reshape(P, B.cols);
    if (tmp.ld != P) SCRC_CUDA(cudaMemsetAsync(tmp.data(), 0, tmp.bytes(), dev.stream));
    DenseArgs a1{eig.V.data(), eig.V.ld, B.data(), B.ld, tmp.data(), tmp.ld, P, (int)B.cols, P};
    a1.alpha = mult; a1.rvec = eig.lam.p; a1.shift = shift; a1.cut = cut;
    gemm_tn_dense(EPI_ROWDIV, a1, dev.num_sms, dev.stream);
    DenseArgs a2{eig.Vt.data(), eig.Vt.ld, tmp.data(), tmp.ld, out.data(), out.ld, P, (int)B.cols, P};
    gemm_tn_dense(EPI_PLAIN, a2, dev.num_sms, dev.stream);
}

static void gram(const DMat& X, const DMat& Y, DMat& out, bool symmetric, Device& dev) {
    out.alloc(X.cols, Y.cols);
    SCRC_CUDA(cudaMemsetAsync(out.data(), 0, out.bytes(), dev.stream));
    DenseArgs a{X.data(), X.ld, Y.data(), Y.ld, out.data(), out.ld, (int)X.cols, (int)Y.cols, (int)X.rows};
    gemm_tn_dense(symmetric ? EPI_SYM_LOWER : EPI_PLAIN, a, dev.num_sms, dev.stream);
}

__global__ void vec_recip_kernel(const double* in, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0 / in[i];
}
__global__ void res_sds_kernel(const double* ss, double denom, double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = sqrt(ss[i] / denom);                       // scregclust.R:1104-1107
}
__global__ void bas_kernel(const double* beta, int64_t ldb, const double* res_sds, double* bas, int64_t ldo, int64_t P,
                           int64_t T) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * T) return;
    const int64_t r = i % P, c = i / P;
    const double v = beta[r + c * ldb] * (1.0 / res_sds[c]);        // scregclust.R:1109-1113
    bas[r + c * ldo] = v * v;
}

// ============================================================================ session creation
void Ctx::begin(int64_t n1_, int64_t n2_, int64_t P_, int64_t T_, int K_) {
    if (n1_ <= 1 || n2_ <= 0 || P_ <= 0 || T_ <= 0 || K_ <= 0) throw Error(SCRC_ERR_ARG, "scrc_ctx_create: non-positive dimension");
    if (K_ > RV_MAX_K) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_ctx_create: at most 64 modules are supported");
    n1 = n1_; n2 = n2_; P = P_; T = T_; K = K_;
    Z1r.alloc(n1, P); Z2r.alloc(n2, P); Z1t.alloc(n1, T); Z2t.alloc(n2, T);
    sds_t.alloc(T); inv_sds_t.alloc(T);
}

void Ctx::create(int device, const double* z1r, const double* z2r, const double* z1t, const double* z2t,
                 const double* sds, int64_t n1_, int64_t n2_, int64_t P_, int64_t T_, int K_, Comm* comm_) {
    if (n1_ <= 1 || n2_ <= 0 || P_ <= 0 || T_ <= 0 || K_ <= 0) throw Error(SCRC_ERR_ARG, "scrc_ctx_create: non-positive dimension");
    comm = comm_;
    dev.init(comm ? comm->device : device);
    AllocScope as(dev.stream);
    begin(n1_, n2_, P_, T_, K_);
    cudaStream_t st = dev.stream;
    // The regulators of the first split go first, on the compute stream: X'X and its eigen-decomposition (the
    // longest, latency-bound part of the precompute) only need them and run while the other three matrices are
    // still arriving on a second stream (pinned host buffers; pageable ones simply serialise).
    cudaStream_t cp = nullptr;
    cudaEvent_t ev_alloc = nullptr, ev_copy = nullptr;
    SCRC_CUDA(cudaStreamCreateWithFlags(&cp, cudaStreamNonBlocking));
    SCRC_CUDA(cudaEventCreateWithFlags(&ev_alloc, cudaEventDisableTiming));
    SCRC_CUDA(cudaEventCreateWithFlags(&ev_copy, cudaEventDisableTiming));
    struct Guard { cudaStream_t s; cudaEvent_t a, b; ~Guard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); } } guard{cp, ev_alloc, ev_copy};
    SCRC_CUDA(cudaEventRecord(ev_alloc, st));          // the buffers were allocated in stream order on `st`
    SCRC_CUDA(cudaStreamWaitEvent(cp, ev_alloc, 0));
    Z1r.upload(z1r, n1, st);
    SCRC_CUDA(cudaMemcpyAsync(sds_t.p, sds, T * 8, cudaMemcpyDefault, st));
    Z1t.upload(z1t, n1, cp); Z2r.upload(z2r, n2, cp); Z2t.upload(z2t, n2, cp);
    SCRC_CUDA(cudaEventRecord(ev_copy, cp));
    finish(ev_copy);
    SCRC_CUDA(cudaStreamSynchronize(cp));
}

// Everything after the four scaled matrices and z1_target_sds are resident.  In a multi-GPU job rank 0
// computes the precompute once and broadcasts it (0.25 GB at config 3, < 1 ms over NVLink): the
// eigen-decomposition is latency-bound and gains nothing from being repeated, and every rank then works
// from bit-identical Gram matrices and eigenvectors by construction.
void Ctx::finish(cudaEvent_t inputs_ready) {
    cudaStream_t st = dev.stream;
    const bool dbg = getenv("SCRC_DEBUG_TIMING") != nullptr;
    const bool root = rank() == 0;
    auto tick = [&](const char* what) {
        static thread_local double t_prev = 0.0;
        if (!dbg) return;
        cudaStreamSynchronize(st);
        timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
        const double t = ts.tv_sec + 1e-9 * ts.tv_nsec;
        if (what) fprintf(stderr, "[scrc create rank %d] %-28s %8.2f ms\n", rank(), what, (t - t_prev) * 1e3);
        t_prev = t;
    };
    tick(nullptr);
    vec_recip_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(sds_t.p, inv_sds_t.p, (int)T);
    SCRC_LAUNCHED();
    tick("inputs resident");
    DevBuf<double> df_d(1);
    if (root) {
        // Gram matrices (optim.cpp:17-27,99,414,421) -- computed once, the reference redoes them per call
        gram(Z1r, Z1r, G_rr, true, dev);
        build_eig(G_rr, 1.0 / (double)n1, eig, dev);   // eigenbasis of X'X with X = Z1r / sqrt(n1)
        tick("X'X + eigen-decomposition");
        if (inputs_ready) SCRC_CUDA(cudaStreamWaitEvent(st, inputs_ready, 0));   // targets and test split have arrived
        gram(Z1r, Z1t, G_rt, false, dev);
        tick("X'Y");
        gtt.alloc(T);
        col_sumsq(Z1t.data(), Z1t.ld, n1, T, gtt.p, st);
        // the same for the test split: SSE_test = y'y - 2 b'X'y + b'X'X b needs no pass over the n2 cells either
        gram(Z2r, Z2r, G2_rr, true, dev);
        gram(Z2r, Z2t, G2_rt, false, dev);
        gtt2.alloc(T);
        col_sumsq(Z2t.data(), Z2t.ld, n2, T, gtt2.p, st);
        tick("test-split gram");

        // Initial fit (scregclust.R:1049-1101): OLS when n1 >= P, ridge with the df search otherwise.
        DMat tmp;
        std::vector<double> lam(P);
        SCRC_CUDA(cudaMemcpyAsync(lam.data(), eig.lam.p, P * 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        double ridge = 0.0;
        if (n1 >= P) {
            init_df = (double)P;
        } else {
            // es = eigenvalues of Z'Z, decreasing; lambda_minimal = es[n1 - 1] + 1e-4 (1-based)
            std::vector<double> es(P);
            for (int64_t i = 0; i < P; ++i) es[i] = lam[P - 1 - i] * (double)n1;
            ridge = es[n1 - 2] + 1e-4;
            auto df = [&](double l) { double s = 0.0; for (double e : es) s += e / (e + l); return s; };
            init_df = df(ridge);
            while (init_df >= (double)n1) { ridge *= 2.0; init_df = df(ridge); }
        }
        eig_solve(eig, (double)n1, ridge, 0.0, G_rt, beta_init, tmp, dev);
        tick("initial fit solve");

        // res_sds = sqrt(colSums((Y - X beta)^2) / (n1 - df))  -- explicit residuals (scregclust.R:1104-1107)
        {
            DMat Z1rT(P, n1), resid(n1, T);
            if (Z1rT.ld != P) SCRC_CUDA(cudaMemsetAsync(Z1rT.data(), 0, Z1rT.bytes(), st));
            transpose(Z1r.data(), Z1r.ld, n1, P, Z1rT.data(), Z1rT.ld, st);
            DenseArgs a{Z1rT.data(), Z1rT.ld, beta_init.data(), beta_init.ld, resid.data(), resid.ld, (int)n1, (int)T, (int)P};
            a.D = Z1t.data(); a.ldd = Z1t.ld;
            gemm_tn_dense(EPI_SUBFROM, a, dev.num_sms, st);
            DevBuf<double> ss(T);
            col_sumsq(resid.data(), resid.ld, n1, T, ss.p, st);
            res_sds.alloc(T);
            res_sds_kernel<<<(unsigned)ceil_div(T, 256), 256, 0, st>>>(ss.p, (double)n1 - init_df, res_sds.p, (int)T);
            SCRC_LAUNCHED();
            SCRC_CUDA(cudaStreamSynchronize(st));
        }
        bas.alloc(P, T);
        bas_kernel<<<(unsigned)ceil_div(P * T, 256), 256, 0, st>>>(beta_init.data(), beta_init.ld, res_sds.p, bas.data(), bas.ld, P, T);
        SCRC_LAUNCHED();
        SCRC_CUDA(cudaMemcpyAsync(df_d.p, &init_df, 8, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        tick("residual sds + weights");
    } else {
        G_rr.alloc(P, P); G_rt.alloc(P, T);
        eig.P = (int)P; eig.V.alloc(P, P); eig.Vt.alloc(P, P); eig.lam.alloc(P);
        gtt.alloc(T); beta_init.alloc(P, T); res_sds.alloc(T); bas.alloc(P, T);
        G2_rr.alloc(P, P); G2_rt.alloc(P, T); gtt2.alloc(T);
        if (inputs_ready) SCRC_CUDA(cudaStreamWaitEvent(st, inputs_ready, 0));
    }
    if (world() > 1) {
        comm->group_begin();
        comm->broadcast_bytes(G_rr.data(), G_rr.bytes(), 0, st);
        comm->broadcast_bytes(G_rt.data(), G_rt.bytes(), 0, st);
        comm->broadcast_bytes(eig.V.data(), eig.V.bytes(), 0, st);
        comm->broadcast_bytes(eig.Vt.data(), eig.Vt.bytes(), 0, st);
        comm->broadcast_bytes(eig.lam.p, (size_t)P * 8, 0, st);
        comm->broadcast_bytes(gtt.p, (size_t)T * 8, 0, st);
        comm->broadcast_bytes(G2_rr.data(), G2_rr.bytes(), 0, st);
        comm->broadcast_bytes(G2_rt.data(), G2_rt.bytes(), 0, st);
        comm->broadcast_bytes(gtt2.p, (size_t)T * 8, 0, st);
        comm->broadcast_bytes(beta_init.data(), beta_init.bytes(), 0, st);
        comm->broadcast_bytes(res_sds.p, (size_t)T * 8, 0, st);
        comm->broadcast_bytes(bas.data(), bas.bytes(), 0, st);
        comm->broadcast_bytes(df_d.p, 8, 0, st);
        comm->group_end();
        SCRC_CUDA(cudaMemcpyAsync(&init_df, df_d.p, 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        tick("precompute broadcast");
    }
}

// ============================================================================ Step 1 helper kernels
// xty column c = X'y for target t(c):  G_rt[:, t] / res_sds[t] / n1   (scregclust.R:1353-1364, optim.cpp:99)
__global__ void __launch_bounds__(256) admm_xty_kernel(const double* __restrict__ G_rt, int64_t ldg,
                                                       const int* __restrict__ col_target,
                                                       const double* __restrict__ res_sds, double n1, double* xty,
                                                       int64_t ld, int P) {
    const int c = blockIdx.x;
    const int t = col_target[c];
    const double sd = res_sds[t];
    for (int i = threadIdx.x; i < P; i += blockDim.x) xty[i + (int64_t)c * ld] = G_rt[i + (int64_t)t * ldg] / sd / n1;
}
// ws = sqrt(m_j) / rowSums(beta_adapt_sq[, k == j])^(1/4)   (scregclust.R:1348-1351)
__global__ void __launch_bounds__(256) admm_weights_kernel(const double* __restrict__ bas, int64_t ldb,
                                                           const int* __restrict__ col_target, const int* c0s,
                                                           const int* c1s, double* w, int64_t ldw, int P) {
    const int p = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int c0 = c0s[p], c1 = c1s[p];
    double s = 0.0;
    for (int c = c0; c < c1; ++c) s += bas[i + (int64_t)col_target[c] * ldb];
    w[i + (int64_t)p * ldw] = sqrt((double)(c1 - c0)) / sqrt(sqrt(s));
}
// beta * res_sds, threshold 1e-6, models / weights of the problems solved on THIS rank (scregclust.R:1372-1385);
// the iteration count of problem p goes to the (fit, module) table that the ranks merge afterwards.
__global__ void __launch_bounds__(256) admm_extract_kernel(const double* __restrict__ beta, int64_t ld,
                                                           const int* __restrict__ col_target,
                                                           const double* __restrict__ res_sds, const int* c0s,
                                                           const int* c1s, const int* prob_fit, const int* prob_mod,
                                                           const int* __restrict__ prob_iters, int P, int K,
                                                           unsigned char* models, double* weights, int* iters_fk) {
    __shared__ double s_sum[8][32];
    __shared__ int s_any[8][32];
    const int p = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 32 + lane;
    const int c0 = c0s[p], c1 = c1s[p];
    double sum = 0.0;
    int any = 0;
    if (row < P) {
        for (int c = c0 + warp; c < c1; c += 8) {
            double v = beta[row + (int64_t)c * ld] * res_sds[col_target[c]];
            if (fabs(v) < 1e-6) v = 0.0;
            if (fabs(v) > 0.0) any = 1;
            sum += v;
        }
    }
    s_sum[warp][lane] = sum;
    s_any[warp][lane] = any;
    __syncthreads();
    const int64_t fk = (int64_t)prob_fit[p] * K + prob_mod[p];
    if (warp == 0 && row < P) {
        double t = 0.0; int a = 0;
        for (int w = 0; w < 8; ++w) { t += s_sum[w][lane]; a |= s_any[w][lane]; }
        models[fk * P + row] = (unsigned char)a;
        weights[fk * P + row] = t / (double)(c1 - c0);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) iters_fk[fk] = prob_iters[p];
}
// signs = sign(weights) where selected, NA elsewhere (scregclust.R:1386); nsel[fk] = #selected regulators.
// Runs on every rank over the merged tables, so every rank derives identical signs.
__global__ void __launch_bounds__(256) signs_nsel_kernel(const unsigned char* __restrict__ models,
                                                         const double* __restrict__ weights, int P, double* signs,
                                                         int* nsel) {
    __shared__ int s_c[8];
    const int64_t fk = blockIdx.x;
    int cnt = 0;
    for (int i = threadIdx.x; i < P; i += 256) {
        const int64_t o = fk * P + i;
        const int a = models[o];
        const double wt = weights[o];
        signs[o] = a ? (double)((wt > 0.0) - (wt < 0.0)) : nan("");
        cnt += a;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s_c[w];
        nsel[fk] = t;
    }
}
// Stacked selection tables of Step 2a from the model indicator, regulators in increasing index per
// (fit, module) as which() returns them (scregclust.R:1541-1547): sel / sign per selected regulator and
// row_src per stacked, 16-row padded row (-1 = padding).
__global__ void __launch_bounds__(256) nnls_fill_kernel(const unsigned char* __restrict__ models,
                                                        const double* __restrict__ signs, int P,
                                                        const int* __restrict__ sel_off, const int* __restrict__ row_off,
                                                        const int* __restrict__ kpad, int* sel, double* sign,
                                                        int* row_src) {
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int64_t fk = blockIdx.x;
    const int kp = kpad[fk];
    if (kp == 0) return;
    const int so = sel_off[fk], ro = row_off[fk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < P; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const bool a = (i < P) && models[fk * P + i];
        const unsigned bal = __ballot_sync(0xffffffffu, a);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) woff += s_warp[w]; tot += s_warp[w]; }
        const int base = s_base;
        if (a) {
            const int r = base + woff + __popc(bal & ((1u << lane) - 1u));
            sel[so + r] = i;
            sign[so + r] = signs[fk * P + i];
            row_src[ro + r] = i;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = base + tot;
        __syncthreads();
    }
    for (int r = s_base + threadIdx.x; r < kp; r += 256) row_src[ro + r] = -1;
}
__global__ void add_one_kernel(int* x, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] += 1;
}

// ---------------------------------------------------------------- quick mode (scregclust.R:1435-1496)
// areg[f][r] = regulator r is selected by some module of fit f (scregclust.R:1447)
__global__ void quick_areg_kernel(const unsigned char* __restrict__ models, int K, int P, unsigned char* areg) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (i >= P) return;
    int a = 0;
    for (int j = 0; j < K; ++j) a |= models[((int64_t)f * K + j) * P + i];
    areg[(int64_t)f * P + i] = (unsigned char)a;
}
// mx[f][t] = max over the active regulators r of |cor(target t, regulator r)| for the noise targets of fit f
// (scregclust.R:1455-1460); corr is cross_corr1 (T x P).  Consecutive threads read consecutive targets.
__global__ void quick_max_kernel(const double* __restrict__ corr, int64_t ldc, const int* __restrict__ idx,
                                 const unsigned char* __restrict__ areg, int T, int P, double* mx) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (t >= T || idx[(int64_t)f * T + t] != -1) return;
    const unsigned char* a = areg + (int64_t)f * P;
    double m = -INFINITY;
    for (int r = 0; r < P; ++r)
        if (a[r]) m = fmax(m, fabs(corr[t + (int64_t)r * ldc]));
    mx[(int64_t)f * T + t] = m;
}
// Local (subset) column space -> all T targets, with the reference's values for the targets Step 2 skipped:
// r2 = 0, resid_var = 1e6 (0 in the last cycle), best_r2 = 0, which.max = 1, and -- all K scores being equal -- every
// cell's vote and hence the gene itself going to module 1 (scregclust.R:1628-1642, allocation.cpp:80-94).
__global__ void quick_scatter_kernel(int F, int K, int T, const int* __restrict__ inv, int n2, int last_cycle,
                                     const double* r2_l, const double* var_l, const double* br_l, const int* bi_l,
                                     const int* kn_l, const int* votes_l, double* r2_g, double* var_g, double* br_g,
                                     int* bi_g, int* kn_g, int* votes_g) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)F * T) return;
    const int f = (int)(i / T), t = (int)(i % T);
    const int c = inv[i];
    const int64_t lo = (int64_t)f * T + (c >= 0 ? c : 0);
    if (br_g) br_g[i] = c >= 0 ? br_l[lo] : 0.0;
    if (bi_g) bi_g[i] = c >= 0 ? bi_l[lo] : 1;
    if (kn_g) kn_g[i] = c >= 0 ? kn_l[lo] : 1;
    for (int j = 0; j < K; ++j) {
        const int64_t g = ((int64_t)f * K + j) * T + t, l = ((int64_t)f * K + j) * T + (c >= 0 ? c : 0);
        if (r2_g) r2_g[g] = c >= 0 ? r2_l[l] : 0.0;
        if (var_g) var_g[g] = c >= 0 ? var_l[l] : (last_cycle ? 0.0 : 1e6);
        if (votes_g) votes_g[i * K + j] = c >= 0 ? votes_l[lo * K + j] : (j == 0 ? n2 : 0);
    }
}
// R's default quantile (type 7) with its own arithmetic (stats:::quantile.default): index = 1 + (n-1) p,
// qs = x[lo]; where index > lo and x[hi] != qs: qs = (1-h) qs + h x[hi].
static double r_quantile7(std::vector<double> x, double prob) {
    const size_t n = x.size();
    if (n == 0) return std::nan("");
    std::sort(x.begin(), x.end());
    const double index = 1.0 + (double)(n - 1) * prob;
    const double lo = std::floor(index), hi = std::ceil(index);
    double qs = x[(size_t)lo - 1];
    if (index > lo && x[(size_t)hi - 1] != qs) {
        const double h = index - lo;
        qs = (1.0 - h) * qs + h * x[(size_t)hi - 1];
    }
    return qs;
}

// The reference documents these as assumptions (allocation.cpp:4-6); on a public C ABI they are checked,
// because the sequential sweep indexes shared and global memory with them.
void validate_prior_inputs(const char* who, const int* k, int64_t T, int K, const int* prior_off, const int* prior_idx,
                           const int* order) {
    auto bad = [&](const std::string& what) { throw Error(SCRC_ERR_ARG, std::string(who) + ": " + what); };
    for (int64_t t = 0; t < T; ++t)
        if (!(k[t] == -1 || (k[t] >= 1 && k[t] <= K))) bad("module ids must be -1 (noise) or in 1..K");
    if (order) {
        std::vector<unsigned char> seen(T, 0);
        for (int64_t t = 0; t < T; ++t) {
            if (order[t] < 0 || order[t] >= T || seen[order[t]]) bad("update_order must be a permutation of 0..T-1");
            seen[order[t]] = 1;
        }
    }
    if (prior_off) {
        if (prior_off[0] != 0) bad("prior_off[0] must be 0");
        for (int64_t t = 0; t < T; ++t)
            if (prior_off[t + 1] < prior_off[t]) bad("prior_off must be non-decreasing");
        if (prior_off[T] > 0 && !prior_idx) bad("prior_idx is missing");
        for (int e = 0; e < prior_off[T]; ++e)
            if (prior_idx[e] < 0 || prior_idx[e] >= T) bad("prior_idx entries must be in 0..T-1");
    }
}

// ============================================================================ one batched cycle
// Multi-GPU layout of a cycle (world > 1; every rank executes this function with the same arguments):
//   Step 1   the non-empty (fit, module) problems are dealt to the ranks by predicted cost (shard.hpp);
//            each rank runs the lock-step ADMM on its problems only;
//   merge 1  all-reduce(SUM) of the module summaries -- selection indicator (1 B), mean coefficient (8 B)
//            per (fit, module, regulator), iteration counts -- every entry written by its owner, 0 elsewhere;
//   Step 2   every rank rebuilds the (tiny) selection tables and runs NNLS, train / test SSE and the votes
//            for ITS slice of the target genes (all fits, all modules);
//   merge 2  all-reduce(SUM) of k_new, best_r2 (and whichever further outputs the caller asked for), slice
//            entries written by their owner, 0 elsewhere; after a skip_allocation cycle the coefficient
//            columns are exchanged as well, so that every rank can return them.
// With a network prior Step 2b is sequential over the genes (allocation.cpp:33-97): Step 1 is still
// spread over the ranks, Step 2 is then replicated on every rank.
void Ctx::fit_cycle(int F, const double* lambda, const int* k_in, const scrc_options& opt, bool skip_alloc,
                    scrc_cycle_out* out, const scrc_prior* prior, const scrc_quick* quick) {
    // aggregate: the allocate_per_obs = FALSE branch of Step 2b (R/scregclust.R:1678-1728); use_prior: the per-cell vote
    // with a network prior (sequential sweep over a materialised residual array)
    const bool aggregate = prior && prior->aggregate != 0 && !skip_alloc;
    const bool agg_prior = aggregate && prior->weight != 0.0 && prior->prior_off != nullptr;
    const bool use_prior = prior && prior->weight != 0.0 && !skip_alloc && !aggregate;
    const bool use_quick = quick && quick->idx_latest;
    if (use_quick && use_prior) throw Error(SCRC_ERR_UNSUPPORTED, "quick mode together with a network prior is not supported");
    if (use_quick && aggregate) throw Error(SCRC_ERR_ARG, "quick_mode only available when allocate_per_obs is TRUE.");
    if (agg_prior && !prior->update_order)
        throw Error(SCRC_ERR_ARG, "scrc_fit_cycle_prior: update_order is required for the aggregate branch with a prior");
    if (use_quick && !(quick->percent >= 0.0 && quick->percent < 1.0))
        throw Error(SCRC_ERR_ARG, "quick_mode_percent needs to be numeric in [0, 1).");
    if (use_prior && (!prior->update_order || !prior->prior_off))
        throw Error(SCRC_ERR_ARG, "scrc_fit_cycle_prior: prior_off and update_order are required when prior_weight != 0");
    if (F <= 0) throw Error(SCRC_ERR_ARG, "scrc_fit_cycle: F must be positive");
    if ((int64_t)F * K > 65535) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_fit_cycle: more than 65535 (fit, module) pairs in one call");
    if (use_prior || agg_prior)
        for (int f = 0; f < F; ++f)
            validate_prior_inputs("scrc_fit_cycle_prior", k_in + (int64_t)f * T, T, K, f == 0 ? prior->prior_off : nullptr,
                                  prior->prior_idx, prior->update_order + (int64_t)f * T);
    SCRC_CUDA(cudaSetDevice(dev.id));
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    const int64_t FK = (int64_t)F * K;
    const int W = world(), R = rank();
    if (W == 1) check_cancel();   // (collective runs agree on an interrupt at merge 1)
    PhaseTimer pt(st);

    // ---------------- Step 1: problem table (one problem per non-empty (fit, module)) ----------------
    std::vector<int> cnt(FK, 0);
    for (int f = 0; f < F; ++f) {
        const int* k = k_in + (int64_t)f * T;
        int* c = cnt.data() + (int64_t)f * K;
        for (int64_t t = 0; t < T; ++t) { const int j = k[t]; if (j >= 1 && j <= K) ++c[j - 1]; }
    }
    std::vector<int> g_fk;
    std::vector<double> g_cost;
    const bool have_hint = (int64_t)cost_hint.size() == FK;
    for (int64_t fk = 0; fk < FK; ++fk)
        if (cnt[fk] > 0) {
            g_fk.push_back((int)fk);
            g_cost.push_back((double)cnt[fk] * (double)((have_hint && cost_hint[fk] > 0) ? cost_hint[fk] : 25));
        }
    std::vector<int> owner(g_fk.size());
    shard_assign(g_cost.data(), (int)g_fk.size(), W, owner.data());
    std::vector<int> prob_of_fk(FK, -1), h_c0, h_c1, h_fit, h_mod;
    std::vector<double> h_lambda;
    int C = 0;
    // Local problems in order of decreasing predicted iteration count (stable; plain (fit, module) order without a
    // hint): problems that converge at about the same sweep sit next to each other, so the converged part of the
    // column range is mostly whole tiles and few live tiles carry dead columns.  The order changes no number.
    std::vector<int> mine;
    for (size_t gi = 0; gi < g_fk.size(); ++gi) if (owner[gi] == R) mine.push_back((int)gi);
    std::stable_sort(mine.begin(), mine.end(), [&](int a, int b) { return g_cost[a] / cnt[g_fk[a]] > g_cost[b] / cnt[g_fk[b]]; });
    for (int gi : mine) {
        const int fk = g_fk[gi];
        prob_of_fk[fk] = (int)h_c0.size();
        h_c0.push_back(C); C += cnt[fk]; h_c1.push_back(C);
        h_fit.push_back(fk / K); h_mod.push_back(fk % K); h_lambda.push_back(lambda[fk / K]);
    }
    cost_hint.clear();
    const int nprob = (int)h_c0.size();
    std::vector<int> h_col_target(C), h_col_prob(C), fillpos(h_c0);
    for (int f = 0; f < F; ++f) {
        const int* k = k_in + (int64_t)f * T;
        const int* pf = prob_of_fk.data() + (int64_t)f * K;
        for (int64_t t = 0; t < T; ++t) {
            const int j = k[t];
            if (j < 1 || j > K) continue;
            const int p = pf[j - 1];
            if (p < 0) continue;
            const int pos = fillpos[p]++;
            h_col_target[pos] = (int)t; h_col_prob[pos] = p;
        }
    }

    models_d.ensure(FK * P); weights_d.ensure(FK * P); signs_d.ensure(FK * P);
    iters_fk_d.ensure(FK + 1); nsel_d.ensure(FK);
    SCRC_CUDA(cudaMemsetAsync(models_d.p, 0, FK * P, st));
    SCRC_CUDA(cudaMemsetAsync(weights_d.p, 0, FK * P * 8, st));
    SCRC_CUDA(cudaMemsetAsync(iters_fk_d.p, 0, (FK + 1) * 4, st));

    long long sweeps = 0;
    bool aborted = false;
    pt.mark(phase_s, 0);
    if (nprob > 0) {
        admm.resize((int)P, C, nprob);
        admm.h_cols.resize(nprob);
        for (int p = 0; p < nprob; ++p) admm.h_cols[p] = h_c1[p] - h_c0[p];
        col_target.ensure(C); prob_fit.ensure(nprob); prob_mod.ensure(nprob);
        SCRC_CUDA(cudaMemcpyAsync(col_target.p, h_col_target.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(admm.col_prob.p, h_col_prob.data(), (size_t)C * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(admm.prob_c0.p, h_c0.data(), nprob * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(admm.prob_c1.p, h_c1.data(), nprob * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(admm.prob_lambda.p, h_lambda.data(), nprob * 8, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(prob_fit.p, h_fit.data(), nprob * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(prob_mod.p, h_mod.data(), nprob * 4, cudaMemcpyHostToDevice, st));
        admm_xty_kernel<<<C, 256, 0, st>>>(G_rt.data(), G_rt.ld, col_target.p, res_sds.p, (double)n1, admm.xty.data(), admm.xty.ld, (int)P);
        SCRC_LAUNCHED();
        admm_weights_kernel<<<dim3((unsigned)ceil_div(P, 256), nprob), 256, 0, st>>>(bas.data(), bas.ld, col_target.p, admm.prob_c0.p, admm.prob_c1.p, admm.w.data(), admm.w.ld, (int)P);
        SCRC_LAUNCHED();
        AdmmOpts ao;
        ao.rho0 = opt.rho0; ao.alpha0 = opt.alpha0; ao.eps_corr = opt.eps_corr; ao.n_update = opt.n_update;
        ao.max_iter = opt.max_optim_iter; ao.eps_rel = opt.tol_coop_rel; ao.eps_abs = opt.tol_coop_abs;
        sweeps = admm_solve(admm, eig, ao, true, dev.num_sms, st, dev.h_pinned, &cancel);
        aborted = admm.aborted;
        admm_extract_kernel<<<dim3((unsigned)ceil_div(P, 32), nprob), 256, 0, st>>>(
            admm.beta.data(), admm.beta.ld, col_target.p, res_sds.p, admm.prob_c0.p, admm.prob_c1.p, prob_fit.p,
            prob_mod.p, admm.prob_iters.p, (int)P, K, models_d.p, weights_d.p, iters_fk_d.p);
        SCRC_LAUNCHED();
    }
    if (aborted || cancel.load(std::memory_order_relaxed)) {
        const int one = 1;
        SCRC_CUDA(cudaMemcpyAsync(iters_fk_d.p + FK, &one, 4, cudaMemcpyHostToDevice, st));
    }
    pt.mark(phase_s, 1);
    // ---------------- merge 1: module summaries ----------------
    if (W > 1) {
        comm->group_begin();
        comm->allreduce_sum_u8(models_d.p, (size_t)(FK * P), st);
        comm->allreduce_sum_f64(weights_d.p, (size_t)(FK * P), st);
        comm->allreduce_sum_i32(iters_fk_d.p, (size_t)(FK + 1), st);
        comm->group_end();
    }
    signs_nsel_kernel<<<(unsigned)FK, 256, 0, st>>>(models_d.p, weights_d.p, (int)P, signs_d.p, nsel_d.p);
    SCRC_LAUNCHED();
    if (h_fk_cap < (size_t)(2 * FK + 1)) {
        if (h_fk) cudaFreeHost(h_fk);
        h_fk_cap = (size_t)(2 * FK + 1) * 2;
        SCRC_CUDA(cudaMallocHost(&h_fk, h_fk_cap * sizeof(int)));
    }
    SCRC_CUDA(cudaMemcpyAsync(h_fk, iters_fk_d.p, (FK + 1) * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaMemcpyAsync(h_fk + FK + 1, nsel_d.p, FK * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    if (h_fk[FK] > 0) throw Error(SCRC_ERR_INTERRUPTED, "interrupted by scrc_ctx_request_cancel");
    const int* h_admm_it = h_fk;
    const int* h_nsel = h_fk + FK + 1;
    pt.mark(phase_s, 2);

    // ---------------- Step 1.5 (quick mode): per-fit target subsets ----------------
    std::vector<int> h_filt, h_foff, h_fcnt, h_finv;
    if (use_quick) {
        if (abs_corr.rows == 0) {     // cross_corr1 = fast_cor(z1_target_centered, z1_reg_scaled), once per session (:1116-1122)
            DMat X(n1, T), Y(n1, P);
            SCRC_CUDA(cudaMemcpyAsync(X.data(), Z1t.data(), Z1t.bytes(), cudaMemcpyDeviceToDevice, st));
            SCRC_CUDA(cudaMemcpyAsync(Y.data(), Z1r.data(), Z1r.bytes(), cudaMemcpyDeviceToDevice, st));
            DevBuf<double> xss(T), yss(P);
            col_center_ss(X.data(), X.ld, n1, T, xss.p, st);
            col_center_ss(Y.data(), Y.ld, n1, P, yss.p, st);
            abs_corr.alloc(T, P);
            DenseArgs a{X.data(), X.ld, Y.data(), Y.ld, abs_corr.data(), abs_corr.ld, (int)T, (int)P, (int)n1};
            a.rvec = xss.p; a.cvec = yss.p;
            gemm_tn_dense(EPI_COR, a, dev.num_sms, st);
            SCRC_CUDA(cudaStreamSynchronize(st));
        }
        q_areg.ensure(FK / K * P); q_idx.ensure((size_t)F * T); q_mx.ensure((size_t)F * T);
        SCRC_CUDA(cudaMemcpyAsync(q_idx.p, quick->idx_latest, (size_t)F * T * 4, cudaMemcpyHostToDevice, st));
        quick_areg_kernel<<<dim3((unsigned)ceil_div(P, 256), F), 256, 0, st>>>(models_d.p, K, (int)P, q_areg.p);
        SCRC_LAUNCHED();
        quick_max_kernel<<<dim3((unsigned)ceil_div(T, 256), F), 256, 0, st>>>(abs_corr.data(), abs_corr.ld, q_idx.p, q_areg.p,
                                                                              (int)T, (int)P, q_mx.p);
        SCRC_LAUNCHED();
        std::vector<double> h_mx((size_t)F * T);
        SCRC_CUDA(cudaMemcpyAsync(h_mx.data(), q_mx.p, h_mx.size() * 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        h_foff.assign(F, 0); h_fcnt.assign(F, 0); h_finv.assign((size_t)F * T, -1);
        for (int f = 0; f < F; ++f) {
            const int* idx = quick->idx_latest + (int64_t)f * T;
            h_foff[f] = (int)h_filt.size();
            int nsel_fit = 0;
            for (int j = 0; j < K; ++j) nsel_fit += h_nsel[(int64_t)f * K + j];
            if (nsel_fit == 0) {                                       // sum(models) == 0: Step 2 sees every target (:1444)
                for (int64_t t = 0; t < T; ++t) h_filt.push_back((int)t);
            } else {
                std::vector<int> noise;
                for (int64_t t = 0; t < T; ++t) { if (idx[t] != -1) h_filt.push_back((int)t); else noise.push_back((int)t); }
                if (!skip_alloc && !noise.empty()) {                   // the last cycle keeps the active targets only (:1451)
                    std::vector<double> v(noise.size());
                    for (size_t q = 0; q < noise.size(); ++q) v[q] = h_mx[(size_t)f * T + noise[q]];
                    const double thr = r_quantile7(v, 1.0 - quick->percent);          // :1463
                    for (size_t q = 0; q < noise.size(); ++q) if (v[q] >= thr) h_filt.push_back(noise[q]);
                }
            }
            h_fcnt[f] = (int)h_filt.size() - h_foff[f];
            for (int c = 0; c < h_fcnt[f]; ++c) h_finv[(size_t)f * T + h_filt[h_foff[f] + c]] = c;
        }
        std::vector<double> h_fsds(h_filt.size());
        if (h_sds.empty()) {
            h_sds.resize(T);
            SCRC_CUDA(cudaMemcpyAsync(h_sds.data(), sds_t.p, T * 8, cudaMemcpyDeviceToHost, st));
            SCRC_CUDA(cudaStreamSynchronize(st));
        }
        for (size_t q = 0; q < h_filt.size(); ++q) h_fsds[q] = h_sds[h_filt[q]];
        nnls.cols.ensure(std::max<size_t>(h_filt.size(), 1)); q_off.ensure(F); q_cnt.ensure(F); q_inv.ensure((size_t)F * T);
        q_sds.ensure(std::max<size_t>(h_filt.size(), 1));
        if (!h_filt.empty()) {
            SCRC_CUDA(cudaMemcpyAsync(nnls.cols.p, h_filt.data(), h_filt.size() * 4, cudaMemcpyHostToDevice, st));
            SCRC_CUDA(cudaMemcpyAsync(q_sds.p, h_fsds.data(), h_fsds.size() * 8, cudaMemcpyHostToDevice, st));
        }
        SCRC_CUDA(cudaMemcpyAsync(q_off.p, h_foff.data(), F * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(q_cnt.p, h_fcnt.data(), F * 4, cudaMemcpyHostToDevice, st));
        SCRC_CUDA(cudaMemcpyAsync(q_inv.p, h_finv.data(), (size_t)F * T * 4, cudaMemcpyHostToDevice, st));
    }
    last_quick = use_quick;

    // ---------------- Step 2a: NNLS problem table ----------------
    slots.assign(FK, FitSlot());
    last_F = F;
    std::vector<NnlsProblem>& h_np = nnls.h_prob;
    h_np.clear();
    std::vector<int> h_np_slot, h_row_off(FK, 0), h_kpad(FK, 0), h_sel_off(FK, 0);
    int64_t rows_total = 0, q_total = 0, sel_total = 0;
    for (int64_t fk = 0; fk < FK; ++fk) {
        FitSlot& sl = slots[fk];
        sl.s = h_nsel[fk];
        if (sl.s == 0) continue;
        const int kp = (int)round_up(sl.s, 16);
        sl.row_off = (int)rows_total; sl.sel_off = (int)sel_total;
        h_row_off[fk] = sl.row_off; h_kpad[fk] = kp; h_sel_off[fk] = sl.sel_off;
        h_np.push_back(make_nnls_problem(sl.s, (int)sel_total, q_total, rows_total));
        if (use_quick) {          // columns = the fit's target subset; R recycles z1_target_sds_tmp over ITS length (:1580-1588)
            NnlsProblem& np = h_np.back();
            const int f = (int)(fk / K);
            np.ncols = h_fcnt[f]; np.xcol_off = h_foff[f]; np.scale_off = h_foff[f]; np.scale_len = h_fcnt[f];
        }
        h_np_slot.push_back((int)fk);
        rows_total += kp; q_total += (int64_t)sl.s * sl.s; sel_total += sl.s;
    }
    if (rows_total > 0x7fffff00LL) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_fit_cycle: too many selected regulators in one batch");
    const int nnp = (int)h_np.size();
    const int64_t rows_alloc = std::max<int64_t>(rows_total, 16);
    nnls.nprob = nnp; nnls.m = (int)T; nnls.rows = rows_total;
    nnls.prob.ensure(std::max(nnp, 1)); nnls.sel.ensure(std::max<int64_t>(sel_total, 1)); nnls.sign.ensure(std::max<int64_t>(sel_total, 1));
    nnls.Q.ensure(std::max<int64_t>(q_total, 1)); nnls.dvec.ensure(rows_alloc);
    nnls.grad0.reshape(rows_alloc, T); nnls.beta.reshape(rows_alloc, T);
    nnls.iters.ensure(std::max<int64_t>((int64_t)nnp * T, 1));
    row_src.ensure(rows_alloc); row_off_d.ensure(FK); kpad_d.ensure(FK); sel_off_d.ensure(FK);
    if (nnp > 0) SCRC_CUDA(cudaMemcpyAsync(nnls.prob.p, h_np.data(), nnp * sizeof(NnlsProblem), cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(row_off_d.p, h_row_off.data(), FK * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(kpad_d.p, h_kpad.data(), FK * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(sel_off_d.p, h_sel_off.data(), FK * 4, cudaMemcpyHostToDevice, st));
    if (nnp > 0) {
        nnls_fill_kernel<<<(unsigned)FK, 256, 0, st>>>(models_d.p, signs_d.p, (int)P, sel_off_d.p, row_off_d.p, kpad_d.p,
                                                       nnls.sel.p, nnls.sign.p, row_src.p);
        SCRC_LAUNCHED();
    }

    // target slice of this rank (whole range with a prior: the sweep is sequential)
    int64_t ts0 = 0, ts1 = T;
    const bool sliced = (W > 1) && !use_prior && !aggregate;   // (the aggregate sweep is tiny: replicated like the prior sweep)
    if (sliced) shard_slice(T, RV_BN, R, W, &ts0, &ts1);
    const int t0 = (int)ts0, t1 = (int)ts1;
    const bool want_nnls_it = out && out->nnls_iterations;
    nnls.col0 = t0; nnls.col1 = t1; nnls.use_active = false;
    pt.mark(phase_s, 3);
    if (rows_total == 0) SCRC_CUDA(cudaMemsetAsync(nnls.beta.data(), 0, nnls.beta.bytes(), st));
    if (sliced && want_nnls_it) SCRC_CUDA(cudaMemsetAsync(nnls.iters.p, 0, (size_t)std::max<int64_t>((int64_t)nnp * T, 1) * 4, st));
    if (nnp > 0) {
        nnls_setup(nnls, G_rr.data(), G_rr.ld, G_rt.data(), G_rt.ld, inv_sds_t.p, st);
        // out scale: z1_target_sds recycled over T -- in quick mode the gathered z1_target_sds_tmp of each fit,
        // recycled over the fit's own subset length (per-problem window, scregclust.R:1580-1588)
        if (use_quick) nnls_solve(nnls, opt.tol_nnls, opt.max_optim_iter, q_sds.p, 1, dev.num_sms, st);
        else nnls_solve(nnls, opt.tol_nnls, opt.max_optim_iter, sds_t.p, (int)T, dev.num_sms, st);
    }

    pt.mark(phase_s, 4);
    // ---------------- Step 2a/2b: residuals, SSE, variances, votes ----------------
    const int CRte = resid_vote_pick_cr((int)n2, t1 - t0, F, dev.num_sms);
    const size_t fkt = (size_t)F * K * T, fk1t = (size_t)F * (K + 1) * T, ft = (size_t)F * T;
    sse_train.ensure(fk1t); sse_test.ensure(fk1t); var.ensure(fkt); two_var.ensure(fkt); inv_two_var.ensure(fkt); half_log.ensure(fkt);
    r2.ensure(fkt); best_r2.ensure(ft); best_idx.ensure(ft); k_new.ensure(ft);
    votes.ensure(fkt);
    const bool vote = !skip_alloc;
    if (sliced) {   // merged outputs: zero outside the slice
        SCRC_CUDA(cudaMemsetAsync(best_r2.p, 0, ft * 8, st));
        if (vote) SCRC_CUDA(cudaMemsetAsync(k_new.p, 0, ft * 4, st));
        if (out && out->r2_test) SCRC_CUDA(cudaMemsetAsync(r2.p, 0, fkt * 8, st));
        if (out && out->resid_var) SCRC_CUDA(cudaMemsetAsync(var.p, 0, fkt * 8, st));
        if (out && out->best_r2_idx) SCRC_CUDA(cudaMemsetAsync(best_idx.p, 0, ft * 4, st));
    }

    ResidVoteArgs ra;
    ra.beta = nnls.beta.data(); ra.ldb = nnls.beta.ld; ra.rows_total = rows_total;
    ra.T = (int)T; ra.K = K; ra.L = F; ra.row_off = row_off_d.p; ra.kpad = kpad_d.p; ra.nsel = nsel_d.p;
    ra.two_var = two_var.p; ra.inv_two_var = inv_two_var.p; ra.half_log = half_log.p; ra.votes = votes.p;
    ra.t_begin = t0; ra.t_end = t1;
    if (use_quick) { ra.col_index = nnls.cols.p; ra.col_off = q_off.p; ra.col_cnt = q_cnt.p; }
    // Both splits: sums of squared residuals from the split's Gram matrices (no pass over the cells) -- the training
    // ones give the variances (scregclust.R:1635-1638), the test ones R2 (scregclust.R:1628-1632); slot K of the
    // test table is the "no model" sum of squares (scregclust.R:1505-1515), the denominator of R2.
    sse_train_gram(G_rr.data(), G_rr.ld, G_rt.data(), G_rt.ld, gtt.p, nnls.beta.data(), nnls.beta.ld, row_src.p,
                   row_off_d.p, nsel_d.p, (int)T, K, F, sse_train.p, st, t0, t1, ra.col_index, ra.col_off, ra.col_cnt);
    make_resid_var(sse_train.p, nsel_d.p, (int)n1, (int)T, K, F, var.p, two_var.p, inv_two_var.p, half_log.p, st, t0, t1);
    sse_train_gram(G2_rr.data(), G2_rr.ld, G2_rt.data(), G2_rt.ld, gtt2.p, nnls.beta.data(), nnls.beta.ld, row_src.p,
                   row_off_d.p, nsel_d.p, (int)T, K, F, sse_test.p, st, t0, t1, ra.col_index, ra.col_off, ra.col_cnt,
                   /*pseudo=*/true);
    // per-cell votes (test split): the only pass over cells of a cycle; skipped in a last cycle
    const bool cell_votes = vote && !aggregate;
    if (cell_votes) {
        ZselT.reshape(rows_alloc, n2);
        if (rows_total == 0) SCRC_CUDA(cudaMemsetAsync(ZselT.data(), 0, ZselT.bytes(), st));
        gather_transpose(Z2r.data(), Z2r.ld, (int)n2, row_src.p, rows_total, ZselT.data(), ZselT.ld, st);
        ra.ZselT = ZselT.data(); ra.ldzs = ZselT.ld; ra.Z = Z2t.data(); ra.ldz = Z2t.ld; ra.ncells = (int)n2; ra.CR = CRte;
    }
    if (aggregate) {
        // allocate_per_obs = FALSE (R/scregclust.R:1678-1728): reassignment from the T x K test sums of squares and
        // residual variances of this cycle; no pass over the cells, no votes
        finish_alloc(sse_test.p, nullptr, (int)T, K, F, r2.p, best_r2.p, best_idx.p, nullptr, st);
        DevBuf<int> d_kin((size_t)ft), d_off(agg_prior ? T + 1 : 1), d_idx(1), d_order(agg_prior ? (size_t)ft : 1);
        DevBuf<double> d_mlp(fkt);
        SCRC_CUDA(cudaMemcpyAsync(d_kin.p, k_in, ft * 4, cudaMemcpyHostToDevice, st));
        if (agg_prior) {
            const int nidx = prior->prior_off[T];
            d_idx.ensure((size_t)std::max(nidx, 1));
            SCRC_CUDA(cudaMemcpyAsync(d_off.p, prior->prior_off, (T + 1) * 4, cudaMemcpyHostToDevice, st));
            if (nidx > 0) SCRC_CUDA(cudaMemcpyAsync(d_idx.p, prior->prior_idx, (size_t)nidx * 4, cudaMemcpyHostToDevice, st));
            SCRC_CUDA(cudaMemcpyAsync(d_order.p, prior->update_order, ft * 4, cudaMemcpyHostToDevice, st));
        }
        alloc_aggregate_device(sse_test.p, var.p, K, T, F, (double)n2, d_kin.p, agg_prior ? d_off.p : nullptr, d_idx.p,
                               agg_prior ? d_order.p : nullptr, prior->baseline, prior->weight, d_mlp.p, k_new.p, st);
        SCRC_CUDA(cudaStreamSynchronize(st));   // the staging buffers above are released with this scope
    } else if (!use_prior) {
        if (vote) {
            SCRC_CUDA(cudaMemsetAsync(votes.p, 0, fkt * 4, st));
            resid_vote(ra, 1, dev.num_sms, st);
        }
        finish_alloc(sse_test.p, vote ? votes.p : nullptr, (int)T, K, F, r2.p, best_r2.p, best_idx.p, k_new.p, st, t0, t1);
    } else {
        // Network prior (allocation.cpp:39-56,74-77): the sweep is sequential in update_order because every
        // reassignment changes the prior of the genes that follow, so each fit materialises its K x n2 x T
        // squared residuals on the device (never on the host) and one CTA walks the genes in order.
        const size_t need = (size_t)K * n2 * T;
        size_t free_b = 0, total_b = 0;
        SCRC_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (sq_arr.n < need && need * 8 + (1ull << 30) > free_b + sq_arr.n * 8)
            throw Error(SCRC_ERR_UNSUPPORTED, "network prior: the K x n2 x T residual array does not fit in device memory");
        sq_arr.ensure(need);
        const int nidx = prior->prior_off[T];
        DevBuf<int> d_off(T + 1), d_idx(std::max(nidx, 1)), d_order(T), d_k(T);
        SCRC_CUDA(cudaMemcpyAsync(d_off.p, prior->prior_off, (T + 1) * 4, cudaMemcpyHostToDevice, st));
        if (nidx > 0) SCRC_CUDA(cudaMemcpyAsync(d_idx.p, prior->prior_idx, (size_t)nidx * 4, cudaMemcpyHostToDevice, st));
        std::vector<int> k0(T);
        ResidVoteArgs r1 = ra;
        r1.L = 1; r1.sq_out = sq_arr.p;
        for (int f = 0; f < F; ++f) {
            r1.row_off = row_off_d.p + (size_t)f * K; r1.kpad = kpad_d.p + (size_t)f * K;
            r1.nsel = nsel_d.p + (size_t)f * K;
            resid_vote(r1, 2, dev.num_sms, st);
            for (int64_t t = 0; t < T; ++t) k0[t] = k_in[(int64_t)f * T + t] - 1;     // allocation.cpp:23
            SCRC_CUDA(cudaMemcpyAsync(d_k.p, k0.data(), T * 4, cudaMemcpyHostToDevice, st));
            SCRC_CUDA(cudaMemcpyAsync(d_order.p, prior->update_order + (int64_t)f * T, T * 4, cudaMemcpyHostToDevice, st));
            alloc_sequential_device(sq_arr.p, K, n2, T, var.p + (size_t)f * K * T, d_off.p, d_idx.p, d_k.p, d_order.p,
                                    prior->baseline, prior->weight, votes.p + (size_t)f * T * K, st);
            SCRC_CUDA(cudaMemcpyAsync(k_new.p + (size_t)f * T, d_k.p, T * 4, cudaMemcpyDeviceToDevice, st));
            SCRC_CUDA(cudaStreamSynchronize(st));   // k0 is reused by the next fit
            check_cancel();                          // allocation.cpp:96 polls per gene; here per fit
        }
        // r2 / best_r2 from the SSEs; k_new (0-based from the sweep) is shifted to 1-based below
        finish_alloc(sse_test.p, nullptr, (int)T, K, F, r2.p, best_r2.p, best_idx.p, nullptr, st);
        add_one_kernel<<<(unsigned)ceil_div((int64_t)F * T, 256), 256, 0, st>>>(k_new.p, (int64_t)F * T);   // allocation.cpp:99
        SCRC_LAUNCHED();
    }

    pt.mark(phase_s, 5);
    // ---------------- merge 2: per-gene results of the target slices ----------------
    if (sliced) {
        comm->group_begin();
        comm->allreduce_sum_f64(best_r2.p, ft, st);
        if (vote) comm->allreduce_sum_i32(k_new.p, ft, st);
        if (out && out->r2_test) comm->allreduce_sum_f64(r2.p, fkt, st);
        if (out && out->resid_var) comm->allreduce_sum_f64(var.p, fkt, st);
        if (out && out->best_r2_idx) comm->allreduce_sum_i32(best_idx.p, ft, st);
        if (vote && out && out->votes) comm->allreduce_sum_i32(votes.p, fkt, st);
        if (want_nnls_it && nnp > 0) comm->allreduce_sum_i32(nnls.iters.p, (size_t)nnp * T, st);
        if (skip_alloc && rows_total > 0)   // coefficient columns: every rank sends its slice (ragged all-gather)
            for (int r = 0; r < W; ++r) {
                int64_t a = 0, b = 0;
                shard_slice(T, RV_BN, r, W, &a, &b);
                comm->broadcast_bytes(nnls.beta.data() + nnls.beta.ld * a, (size_t)(nnls.beta.ld * (b - a)) * 8, r, st);
            }
        comm->group_end();
    }

    // quick mode: subset column space -> all targets, with the reference's values for the skipped targets
    double *o_r2 = r2.p, *o_var = var.p, *o_br = best_r2.p;
    int *o_bi = best_idx.p, *o_kn = k_new.p, *o_votes = votes.p;
    if (use_quick) {
        q_r2.ensure(fkt); q_var.ensure(fkt); q_br.ensure(ft); q_bi.ensure(ft); q_kn.ensure(ft);
        const bool wv = vote && out && out->votes;
        if (wv) q_votes.ensure(fkt);
        quick_scatter_kernel<<<(unsigned)ceil_div((int64_t)ft, 256), 256, 0, st>>>(
            F, K, (int)T, q_inv.p, (int)n2, skip_alloc ? 1 : 0, r2.p, var.p, best_r2.p, best_idx.p, vote ? k_new.p : nullptr,
            wv ? votes.p : nullptr, q_r2.p, q_var.p, q_br.p, q_bi.p, vote ? q_kn.p : nullptr, wv ? q_votes.p : nullptr);
        SCRC_LAUNCHED();
        o_r2 = q_r2.p; o_var = q_var.p; o_br = q_br.p; o_bi = q_bi.p; o_kn = q_kn.p; o_votes = q_votes.p;
    }
    pt.mark(phase_s, 6);
    // ---------------- outputs ----------------
    if (out) {
        auto d2h = [&](void* dst, const void* src, size_t bytes) {
            if (dst && bytes) SCRC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st));
        };
        if (vote) { d2h(out->k_new, o_kn, ft * 4); if (cell_votes) d2h(out->votes, o_votes, fkt * 4); }
        if (aggregate && out->votes) std::memset(out->votes, 0, fkt * 4);   // the aggregate branch casts no votes
        d2h(out->models, models_d.p, FK * P);
        d2h(out->weights, weights_d.p, FK * P * 8);
        d2h(out->signs, signs_d.p, FK * P * 8);
        d2h(out->r2_test, o_r2, fkt * 8);
        d2h(out->resid_var, o_var, fkt * 8);
        d2h(out->best_r2, o_br, ft * 8);
        d2h(out->best_r2_idx, o_bi, ft * 4);
        if (out->admm_iterations) std::memcpy(out->admm_iterations, h_admm_it, FK * 4);
        if (out->n_selected) std::memcpy(out->n_selected, h_nsel, FK * 4);
        if (out->admm_sweeps) *out->admm_sweeps = sweeps;
        if (out->nnls_iterations) {
            std::memset(out->nnls_iterations, 0, fkt * 4);
            if (!use_quick) {
                for (int q = 0; q < nnp; ++q)
                    d2h(out->nnls_iterations + (size_t)h_np_slot[q] * T, nnls.iters.p + (size_t)q * T, (size_t)T * 4);
            } else if (nnp > 0) {           // subset column space -> targets
                std::vector<int> loc((size_t)nnp * T);
                SCRC_CUDA(cudaMemcpyAsync(loc.data(), nnls.iters.p, loc.size() * 4, cudaMemcpyDeviceToHost, st));
                SCRC_CUDA(cudaStreamSynchronize(st));
                for (int q = 0; q < nnp; ++q) {
                    const int f = h_np_slot[q] / K;
                    for (int c = 0; c < h_fcnt[f]; ++c)
                        out->nnls_iterations[(size_t)h_np_slot[q] * T + h_filt[h_foff[f] + c]] = loc[(size_t)q * T + c];
                }
            }
        }
    }
    SCRC_CUDA(cudaStreamSynchronize(st));
    pt.mark(phase_s, 7);
}

// out[base[r] + a[r] + s[r] * t] = beta[src[r] + ld * t]: drops the 16-row padding of the stacked layout and
// lays every module's s x T block out contiguously, so that one linear copy brings a whole fit to the host.
// inv (optional, quick mode): column of `beta` that holds target t, or -1 -- the reference returns zero coefficients
// for the targets Step 2 skipped (scregclust.R:1606-1610).
__global__ void coeff_pack_kernel(const double* __restrict__ beta, int64_t ld, int T, int nrows,
                                  const int* __restrict__ src, const int* __restrict__ s_of, const int* __restrict__ a_of,
                                  const long long* __restrict__ base, const int* __restrict__ inv, double* out) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (r >= nrows || t >= T) return;
    const int c = inv ? inv[t] : t;
    out[base[r] + a_of[r] + (long long)s_of[r] * t] = c >= 0 ? beta[src[r] + ld * c] : 0.0;
}

// Packs the K coefficient matrices of fit f (s_k x T each, one after the other) into `out_dev` on the device.
void Ctx::pack_coeffs_fit(int f, double* out_dev) {
    if (f < 0 || f >= last_F) throw Error(SCRC_ERR_ARG, "scrc_ctx_get_coeffs_fit: index out of range");
    SCRC_CUDA(cudaSetDevice(dev.id));
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    PhaseTimer pt(st);
    std::vector<int> src, s_of, a_of;
    std::vector<long long> base;
    long long off = 0;
    for (int k = 0; k < K; ++k) {
        const FitSlot& sl = slots[(size_t)f * K + k];
        for (int a = 0; a < sl.s; ++a) { src.push_back(sl.row_off + a); s_of.push_back(sl.s); a_of.push_back(a); base.push_back(off); }
        off += (long long)sl.s * T;
    }
    if (src.empty()) return;
    const int nrows = (int)src.size();
    pack_src.ensure(nrows); pack_s.ensure(nrows); pack_a.ensure(nrows); pack_base.ensure(nrows);
    SCRC_CUDA(cudaMemcpyAsync(pack_src.p, src.data(), nrows * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(pack_s.p, s_of.data(), nrows * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(pack_a.p, a_of.data(), nrows * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(pack_base.p, base.data(), nrows * 8, cudaMemcpyHostToDevice, st));
    coeff_pack_kernel<<<dim3((unsigned)ceil_div(nrows, 128), (unsigned)T), 128, 0, st>>>(
        nnls.beta.data(), nnls.beta.ld, (int)T, nrows, pack_src.p, pack_s.p, pack_a.p, pack_base.p,
        last_quick ? q_inv.p + (size_t)f * T : nullptr, out_dev);
    SCRC_LAUNCHED();
    SCRC_CUDA(cudaStreamSynchronize(st));
    pt.mark(phase_s, 12);
}

void Ctx::get_coeffs_fit(int f, double* coeffs_out, int* sel_out) {
    if (f < 0 || f >= last_F) throw Error(SCRC_ERR_ARG, "scrc_ctx_get_coeffs_fit: index out of range");
    SCRC_CUDA(cudaSetDevice(dev.id));
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    int nsel_fit = 0, first = -1;
    for (int k = 0; k < K; ++k) {
        const FitSlot& sl = slots[(size_t)f * K + k];
        if (sl.s > 0 && first < 0) first = sl.sel_off;
        nsel_fit += sl.s;
    }
    if (nsel_fit == 0) return;
    // the selections of a fit are contiguous in the stacked list (fits and modules in order)
    if (sel_out) SCRC_CUDA(cudaMemcpyAsync(sel_out, nnls.sel.p + first, (size_t)nsel_fit * 4, cudaMemcpyDeviceToHost, st));
    if (coeffs_out) {
        pack_out.ensure((size_t)nsel_fit * T);
        pack_coeffs_fit(f, pack_out.p);
        SCRC_CUDA(cudaMemcpyAsync(coeffs_out, pack_out.p, (size_t)nsel_fit * T * 8, cudaMemcpyDeviceToHost, st));
    }
    SCRC_CUDA(cudaStreamSynchronize(st));
}

// ============================================================================ importance (next row f1)
// R/scregclust.R:1823-1912.  For every final configuration f and module j with >1 regulators and
// >0 targets: refit the sign-constrained NNLS on the module's targets with each regulator left out
// in turn (r = 0: none), measure the test SSE per target, and derive
//   r2_removed_vec[r] = 1 - sum_t ssq[t, r] / sum(scale(z2[, targets], center = TRUE)^2)
//   r2_module[j] = r2_removed_vec[0],  importance[reg_r, j] = 1 - r2_removed_vec[r] / r2_module[j].
// All (f, j, r) problems run in one batched NNLS launch and one batched residual launch; the s + 1
// problems of a module share one gathered regulator block (the removed regulator's coefficient row
// is kept as an exact zero row) and differ only in their coefficient rows.
// Multi-GPU: every rank builds the same tables, solves the problems of the modules dealt to it
// (shard.hpp) and the per-target SSE table is merged with one all-reduce(SUM).
void Ctx::importance(int F, const int* k_final, const unsigned char* models, const double* signs,
                     const scrc_options& opt, double* r2_module, double* importance_out, double* r2_removed,
                     int64_t r2_removed_capacity) {
    if (F <= 0) throw Error(SCRC_ERR_ARG, "scrc_importance: F must be positive");
    SCRC_CUDA(cudaSetDevice(dev.id));
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    const int64_t FK = (int64_t)F * K;
    const int W = world(), R = rank();
    if (W == 1) check_cancel();
    PhaseTimer pt(st);
    const double nan_v = std::nan("");
    for (int64_t i = 0; i < FK; ++i) r2_module[i] = nan_v;
    for (int64_t i = 0; i < FK * P; ++i) importance_out[i] = nan_v;

    struct Group { int f, j, s, spad, m; int arow; int col_off; std::vector<int> regs; int first_prob; };
    std::vector<Group> groups;
    NnlsBatch& nb = imp_nb;
    std::vector<NnlsProblem>& h_np = nb.h_prob;
    h_np.clear();
    std::vector<int> h_sel, h_pos, h_cols, h_row_src, h_arow, h_brow, h_kpad, h_coff, h_ccnt;
    std::vector<double> h_sign, g_cost;
    int64_t rows_a = 0, rows_b = 0, q_total = 0;
    int Tmax = 0;
    for (int f = 0; f < F; ++f) {
        const int* k = k_final + (int64_t)f * T;
        // targets of every module in increasing index (one pass over the configuration)
        std::vector<std::vector<int>> cols_of(K);
        for (int64_t t = 0; t < T; ++t) { const int j = k[t]; if (j >= 1 && j <= K) cols_of[j - 1].push_back((int)t); }
        for (int j = 0; j < K; ++j) {
            Group g;
            g.f = f; g.j = j;
            const unsigned char* mk = models + ((int64_t)f * K + j) * P;
            for (int64_t i = 0; i < P; ++i) if (mk[i]) g.regs.push_back((int)i);
            g.s = (int)g.regs.size();
            g.m = (int)cols_of[j].size();
            if (!(g.m > 0 && g.s > 1)) continue;                                    // scregclust.R:1839
            g.col_off = (int)h_cols.size();
            h_cols.insert(h_cols.end(), cols_of[j].begin(), cols_of[j].end());
            g.spad = (int)round_up(g.s, 16);
            g.arow = (int)rows_a;
            g.first_prob = (int)h_np.size();
            for (int a = 0; a < g.s; ++a) h_row_src.push_back(g.regs[a]);
            for (int a = g.s; a < g.spad; ++a) h_row_src.push_back(-1);
            rows_a += g.spad;
            Tmax = std::max(Tmax, g.m);
            const double* sg = signs + ((int64_t)f * K + j) * P;
            for (int r = 0; r <= g.s; ++r) {                                        // scregclust.R:1844-1847
                const int sp = g.s - (r > 0 ? 1 : 0);
                NnlsProblem np = make_nnls_problem(sp, (int)h_sel.size(), q_total, rows_b);
                np.ncols = g.m; np.xcol_off = g.col_off; np.pos_off = (int)h_sel.size();
                np.scale_off = g.col_off; np.scale_len = g.m;
                for (int a = 0; a < g.s; ++a) {
                    if (r > 0 && a == r - 1) continue;
                    h_sel.push_back(g.regs[a]); h_sign.push_back(sg[g.regs[a]]); h_pos.push_back(a);
                }
                h_np.push_back(np);
                h_arow.push_back(g.arow); h_brow.push_back((int)rows_b); h_kpad.push_back(g.spad);
                h_coff.push_back(g.col_off); h_ccnt.push_back(g.m);
                rows_b += g.spad; q_total += (int64_t)sp * sp;
            }
            g_cost.push_back((double)g.m * (g.s + 1.0) * g.s * (g.s + 16.0));
            groups.push_back(std::move(g));
        }
    }
    if (groups.empty()) return;
    if (rows_b > 0x7fffff00LL) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_importance: batch too large; pass fewer configurations per call");
    const int nprob = (int)h_np.size();
    // modules dealt to the ranks; the problems of the others are skipped here (zero columns / not in `active`)
    std::vector<int> owner(groups.size());
    shard_assign(g_cost.data(), (int)groups.size(), W, owner.data());
    nb.use_active = (W > 1);
    nb.active.clear();
    if (W > 1)
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            const Group& g = groups[gi];
            for (int r = 0; r <= g.s; ++r) {
                if (owner[gi] == R) nb.active.push_back(g.first_prob + r);
                else h_ccnt[g.first_prob + r] = 0;
            }
        }

    nb.nprob = nprob; nb.m = Tmax; nb.rows = rows_b; nb.col0 = 0; nb.col1 = -1;
    nb.prob.ensure(nprob); nb.sel.ensure(h_sel.size()); nb.sign.ensure(h_sign.size()); nb.pos.ensure(h_pos.size());
    nb.cols.ensure(h_cols.size()); nb.Q.ensure(q_total); nb.dvec.ensure(rows_b);
    nb.grad0.reshape(rows_b, Tmax); nb.beta.reshape(rows_b, Tmax); nb.iters.ensure((size_t)nprob * Tmax);
    imp_scale.ensure(h_cols.size());
    if (h_sds.empty()) {
        h_sds.resize(T);
        SCRC_CUDA(cudaMemcpyAsync(h_sds.data(), sds_t.p, T * 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<double> h_scale(h_cols.size());
    for (size_t i = 0; i < h_cols.size(); ++i) h_scale[i] = h_sds[h_cols[i]];
    SCRC_CUDA(cudaMemcpyAsync(nb.prob.p, h_np.data(), nprob * sizeof(NnlsProblem), cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(nb.sel.p, h_sel.data(), h_sel.size() * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(nb.sign.p, h_sign.data(), h_sign.size() * 8, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(nb.pos.p, h_pos.data(), h_pos.size() * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(nb.cols.p, h_cols.data(), h_cols.size() * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_scale.p, h_scale.data(), h_scale.size() * 8, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemsetAsync(nb.grad0.data(), 0, nb.grad0.bytes(), st));
    SCRC_CUDA(cudaMemsetAsync(nb.beta.data(), 0, nb.beta.bytes(), st));   // left-out rows and padding stay exact zeros
    pt.mark(phase_s, 8);
    nnls_setup(nb, G_rr.data(), G_rr.ld, G_rt.data(), G_rt.ld, inv_sds_t.p, st);
    nnls_solve(nb, opt.tol_nnls, opt.max_optim_iter, imp_scale.p, 1, dev.num_sms, st);   // scregclust.R:1865-1873

    pt.mark(phase_s, 9);
    // test-split SSE per (problem, target of the module) (scregclust.R:1875-1878), from the test split's Gram
    // matrices: the s + 1 problems of a module share its regulator list (rows imp_arow of imp_row_src), their
    // coefficient blocks (rows imp_brow of nb.beta) carry an exact zero row for the regulator left out
    std::vector<int> h_ns(nprob);
    for (const Group& g : groups) for (int r = 0; r <= g.s; ++r) h_ns[g.first_prob + r] = g.s;
    imp_row_src.ensure(rows_a); imp_arow.ensure(nprob); imp_brow.ensure(nprob); imp_kpad.ensure(nprob);
    imp_coff.ensure(nprob); imp_ccnt.ensure(nprob);
    SCRC_CUDA(cudaMemcpyAsync(imp_row_src.p, h_row_src.data(), rows_a * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_arow.p, h_arow.data(), nprob * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_brow.p, h_brow.data(), nprob * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_kpad.p, h_ns.data(), nprob * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_coff.p, h_coff.data(), nprob * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(imp_ccnt.p, h_ccnt.data(), nprob * 4, cudaMemcpyHostToDevice, st));
    imp_sse.ensure((size_t)nprob * 2 * Tmax);
    if (W > 1) SCRC_CUDA(cudaMemsetAsync(imp_sse.p, 0, (size_t)nprob * 2 * Tmax * 8, st));
    sse_train_gram(G2_rr.data(), G2_rr.ld, G2_rt.data(), G2_rt.ld, gtt2.p, nb.beta.data(), nb.beta.ld, imp_row_src.p,
                   imp_brow.p, imp_kpad.p, Tmax, 1, nprob, imp_sse.p, st, 0, Tmax, nb.cols.p, imp_coff.p, imp_ccnt.p,
                   /*pseudo=*/false, /*sel_row_off=*/imp_arow.p);
    if (W > 1) comm->allreduce_sum_f64(imp_sse.p, (size_t)nprob * 2 * Tmax, st);
    if (z2t_css.n == 0) { z2t_css.alloc(T); col_centered_ss(Z2t.data(), Z2t.ld, n2, T, z2t_css.p, st); }
    std::vector<double>& h_sse = imp_h_sse;
    h_sse.resize((size_t)nprob * 2 * Tmax);
    std::vector<double> h_css(T);
    SCRC_CUDA(cudaMemcpyAsync(h_sse.data(), imp_sse.p, h_sse.size() * 8, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaMemcpyAsync(h_css.data(), z2t_css.p, T * 8, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    pt.mark(phase_s, 10);

    int64_t rr_off = 0;
    for (const Group& g : groups) {
        if (r2_removed) {
            // r2_removed[[m]][[j]] = 1 - t(t(ssq) / colSums(scale(z2...)^2))   (scregclust.R:1890-1896): the
            // (s+1) x m matrix t(ssq) is divided by a length-m vector, which R recycles down ITS columns, so
            // entry (target c, removal r) is divided by css[(r + c (s+1)) mod m] -- reproduced as is.
            const int64_t need = (int64_t)g.m * (g.s + 1);
            if (rr_off + need > r2_removed_capacity) throw Error(SCRC_ERR_ARG, "scrc_importance_ex: r2_removed buffer too small");
            for (int r = 0; r <= g.s; ++r) {
                const double* q = h_sse.data() + ((size_t)(g.first_prob + r) * 2 + 0) * Tmax;
                for (int c = 0; c < g.m; ++c) {
                    const int64_t flat = (int64_t)r + (int64_t)c * (g.s + 1);
                    const double css = h_css[h_cols[g.col_off + (int)(flat % g.m)]];
                    r2_removed[rr_off + c + (int64_t)r * g.m] = 1.0 - q[c] / css;
                }
            }
            rr_off += need;
        }
        long double denom = 0.0L;
        for (int c = 0; c < g.m; ++c) denom += h_css[h_cols[g.col_off + c]];            // scregclust.R:1883-1887
        std::vector<double> vec(g.s + 1);
        for (int r = 0; r <= g.s; ++r) {
            const double* q = h_sse.data() + ((size_t)(g.first_prob + r) * 2 + 0) * Tmax;
            long double tot = 0.0L;
            for (int c = 0; c < g.m; ++c) tot += q[c];
            vec[r] = 1.0 - (double)(tot / denom);                                       // scregclust.R:1882
        }
        r2_module[(int64_t)g.f * K + g.j] = vec[0];                                     // scregclust.R:1901
        for (int r = 1; r <= g.s; ++r)
            importance_out[((int64_t)g.f * K + g.j) * P + g.regs[r - 1]] = 1.0 - vec[r] / vec[0];   // scregclust.R:1904-1909
    }
    pt.mark(phase_s, 11);
}

// ============================================================================ drop-ins
void dropin_crossprod(Device& dev, const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    DMat X(n, px), C;
    X.upload(x, n, st);
    if (y == x && px == py) {
        gram(X, X, C, true, dev);
    } else {
        DMat Y(n, py);
        Y.upload(y, n, st);
        gram(X, Y, C, false, dev);
        C.download(out, px, st);
        SCRC_CUDA(cudaStreamSynchronize(st));
        return;
    }
    C.download(out, px, st);
    SCRC_CUDA(cudaStreamSynchronize(st));
}

void dropin_fast_cor(Device& dev, const double* x, const double* y, int64_t n, int64_t px, int64_t py, double* out) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    DMat X(n, px), Y(n, py), C(px, py);
    X.upload(x, n, st); Y.upload(y, n, st);
    DevBuf<double> xss(px), yss(py);
    col_center_ss(X.data(), X.ld, n, px, xss.p, st);    // utils.R:534-537
    col_center_ss(Y.data(), Y.ld, n, py, yss.p, st);
    DenseArgs a{X.data(), X.ld, Y.data(), Y.ld, C.data(), C.ld, (int)px, (int)py, (int)n};
    a.rvec = xss.p; a.cvec = yss.p;
    gemm_tn_dense(EPI_COR, a, dev.num_sms, st);          // utils.R:538-540
    C.download(out, px, st);
    SCRC_CUDA(cudaStreamSynchronize(st));
}

void dropin_ols(Device& dev, const double* y, const double* x, int64_t n, int64_t p, int64_t m, double lambda, bool ridge,
                double* beta_out) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    DMat X(n, p), Y(n, m), G, Gy, B, tmp;
    X.upload(x, n, st); Y.upload(y, n, st);
    gram(X, X, G, true, dev);
    gram(X, Y, Gy, false, dev);
    EigBasis eig;
    build_eig(G, 1.0, eig, dev);
    double shift = 0.0, cut = 0.0;
    if (ridge) {
        shift = lambda;
    } else if (n < p) {
        // Moore-Penrose inverse of X'X: drop eigenvalues <= eps * p * max(d)   (utils.R:20-30)
        double lmax = 0.0;
        SCRC_CUDA(cudaMemcpyAsync(&lmax, eig.lam.p + (p - 1), 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        cut = 2.220446049250313e-16 * (double)p * lmax;
    }
    eig_solve(eig, 1.0, shift, cut, Gy, B, tmp, dev);
    B.download(beta_out, p, st);
    SCRC_CUDA(cudaStreamSynchronize(st));
}

__global__ void set_warm_start_kernel(const double* b0, int64_t ld0, double* beta, double* zeta, double* beta_old,
                                      double* zeta_old, int64_t ld, int64_t P, int64_t m) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P * m) return;
    const int64_t r = i % P, c = i / P;
    const double v = b0[r + c * ld0];
    const int64_t o = r + c * ld;
    beta[o] = v; zeta[o] = v; beta_old[o] = v; zeta_old[o] = v;      // optim.cpp:84-112
}

void dropin_coop_lasso(Device& dev, const double* y, int64_t n, int64_t m, const double* x, int64_t p, double lambda,
                       const double* weights, const double* beta0, const AdmmOpts& o, double* beta_out, int* iters) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    DMat X(n, p), Y(n, m), G, Gy;
    X.upload(x, n, st); Y.upload(y, n, st);
    gram(X, X, G, true, dev);       // optim.cpp:97
    gram(X, Y, Gy, false, dev);     // optim.cpp:99
    EigBasis eig;
    build_eig(G, 1.0, eig, dev);
    AdmmBatch b;
    b.resize((int)p, (int)m, 1);
    b.h_cols.assign(1, (int)m);
    std::vector<int> cp(m, 0);
    const int c0 = 0, c1 = (int)m;
    SCRC_CUDA(cudaMemcpyAsync(b.col_prob.p, cp.data(), m * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(b.prob_c0.p, &c0, 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(b.prob_c1.p, &c1, 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(b.prob_lambda.p, &lambda, 8, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpy2DAsync(b.w.data(), b.w.ld * 8, weights, p * 8, p * 8, 1, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpy2DAsync(b.xty.data(), b.xty.ld * 8, Gy.data(), Gy.ld * 8, p * 8, m, cudaMemcpyDeviceToDevice, st));
    bool cold = true;
    DMat B0;
    if (beta0) {
        cold = false;
        B0.alloc(p, m);
        B0.upload(beta0, p, st);
        set_warm_start_kernel<<<(unsigned)ceil_div(p * m, 256), 256, 0, st>>>(B0.data(), B0.ld, b.beta.data(), b.zeta.data(), b.beta_old.data(), b.zeta_old.data(), b.beta.ld, p, m);
        SCRC_LAUNCHED();
    }
    SCRC_CUDA(cudaStreamSynchronize(st));   // host temporaries above must outlive the copies
    admm_solve(b, eig, o, cold, dev.num_sms, st, dev.h_pinned);
    b.beta.download(beta_out, p, st);
    SCRC_CUDA(cudaMemcpyAsync(iters, b.prob_iters.p, 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
}

void dropin_nnls(Device& dev, const double* x, int64_t nobs, int64_t nvar, const double* y, int64_t m, double eps,
                 int max_iter, double* beta_out, int* iters) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    DMat X(nobs, nvar), Y(nobs, m), G, Gy;
    X.upload(x, nobs, st); Y.upload(y, nobs, st);
    gram(X, X, G, true, dev);       // optim.cpp:414
    gram(X, Y, Gy, false, dev);     // optim.cpp:421
    NnlsBatch b;
    b.nprob = 1; b.m = (int)m;
    const int64_t rows = round_up(nvar, 16);
    b.rows = rows;
    NnlsProblem np = make_nnls_problem((int)nvar, 0, 0, 0);
    std::vector<int> sel(nvar);
    std::iota(sel.begin(), sel.end(), 0);
    std::vector<double> sg(nvar, 1.0);
    b.prob.alloc(1); b.sel.alloc(nvar); b.sign.alloc(nvar); b.Q.alloc((size_t)nvar * nvar); b.dvec.alloc(rows);
    b.grad0.alloc(rows, m); b.beta.alloc(rows, m); b.iters.alloc(m);
    b.h_prob.assign(1, np);
    SCRC_CUDA(cudaMemcpyAsync(b.prob.p, &np, sizeof(np), cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(b.sel.p, sel.data(), nvar * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemcpyAsync(b.sign.p, sg.data(), nvar * 8, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    nnls_setup(b, G.data(), G.ld, Gy.data(), Gy.ld, nullptr, st);
    nnls_solve(b, eps, max_iter, nullptr, 1, dev.num_sms, st);
    SCRC_CUDA(cudaMemcpy2DAsync(beta_out, nvar * 8, b.beta.data(), b.beta.ld * 8, nvar * 8, m, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaMemcpyAsync(iters, b.iters.p, m * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
}

// ---------------------------------------------------------------- allocate_into_modules on a materialised array
// One CTA per gene: per-cell Gaussian log-likelihood per module, first-max argmax, vote histogram
// (allocation.cpp:59-94).  plp (K values, may be nullptr) is added with weight w.
__device__ void gene_votes(const double* __restrict__ resid, int K, int64_t n_obs, const double* __restrict__ var_row,
                           int64_t var_ld, double w, const double* plp, int* s_votes) {
    for (int64_t c = threadIdx.x; c < n_obs; c += blockDim.x) {
        const double* rc = resid + c * K;
        double best = -INFINITY;
        int bi = 0;
        for (int k = 0; k < K; ++k) {
            const double v = var_row[(int64_t)k * var_ld];
            double sc = (-rc[k]) / (2.0 * v) - 0.5 * log((2.0 * M_PI) * v);
            if (plp) sc = (1.0 - w) * sc + w * plp[k];                     // allocation.cpp:74-77
            if (sc > best) { best = sc; bi = k; }
        }
        atomicAdd(&s_votes[bi], 1);
    }
}

__global__ void __launch_bounds__(256) alloc_parallel_kernel(const double* __restrict__ arr, int K, int64_t n_obs,
                                                             int64_t g0, int64_t ng, const double* __restrict__ resid_var,
                                                             int64_t T, int* k_out, int* votes_out) {
    __shared__ int s_votes[RV_MAX_K];
    const int64_t g = blockIdx.x;
    if (g >= ng) return;
    for (int k = threadIdx.x; k < K; k += blockDim.x) s_votes[k] = 0;
    __syncthreads();
    gene_votes(arr + g * K * n_obs, K, n_obs, resid_var + (g0 + g), T, 0.0, nullptr, s_votes);
    __syncthreads();
    if (threadIdx.x == 0) {
        int bv = -1, bk = 0;
        for (int k = 0; k < K; ++k) {
            if (s_votes[k] > bv) { bv = s_votes[k]; bk = k; }
            if (votes_out) votes_out[(g0 + g) * K + k] = s_votes[k];
        }
        k_out[g0 + g] = bk + 1;
    }
}

// Sequential Gauss-Seidel sweep with the network prior (allocation.cpp:33-97): one CTA walks
// the genes in `order`; the module totals and k live in shared / global memory.
__global__ void __launch_bounds__(1024) alloc_sequential_kernel(const double* __restrict__ arr, int K, int64_t n_obs,
                                                                int64_t T, const double* __restrict__ resid_var,
                                                                const int* __restrict__ prior_off,
                                                                const int* __restrict__ prior_idx, int* k /* 0-based, -2 noise */,
                                                                const int* __restrict__ order, double baseline, double w,
                                                                int* votes_out) {
    __shared__ int s_votes[RV_MAX_K];
    __shared__ double s_tot[RV_MAX_K], s_plp[RV_MAX_K];
    if (threadIdx.x < K) s_tot[threadIdx.x] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int64_t t = 0; t < T; ++t) if (k[t] != -2) s_tot[k[t]] += 1.0;   // allocation.cpp:25-30
    __syncthreads();
    for (int64_t oi = 0; oi < T; ++oi) {
        const int64_t j = order[oi];
        if (threadIdx.x < K) s_votes[threadIdx.x] = 0;
        if (threadIdx.x == 0) {
            double frac[RV_MAX_K];
            for (int q = 0; q < K; ++q) frac[q] = 0.0;
            if (prior_off)
                for (int e = prior_off[j]; e < prior_off[j + 1]; ++e) {
                    const int kk = k[prior_idx[e]];
                    if (kk != -2) frac[kk] += 1.0;                            // allocation.cpp:41-46
                }
            double sum = 0.0;
            for (int q = 0; q < K; ++q) {
                if (s_tot[q] > 0.0) frac[q] /= s_tot[q];
                frac[q] += baseline;
                sum += frac[q];
            }
            const double ls = log(sum);
            for (int q = 0; q < K; ++q) s_plp[q] = log(frac[q]) - ls;        // allocation.cpp:56
        }
        __syncthreads();
        gene_votes(arr + j * K * n_obs, K, n_obs, resid_var + j, T, w, s_plp, s_votes);
        __syncthreads();
        if (threadIdx.x == 0) {
            int bv = -1, bk = 0;
            for (int q = 0; q < K; ++q) {
                if (s_votes[q] > bv) { bv = s_votes[q]; bk = q; }
                if (votes_out) votes_out[j * K + q] = s_votes[q];
            }
            if (k[j] != -2) s_tot[k[j]] -= 1.0;
            s_tot[bk] += 1.0;
            k[j] = bk;                                                        // allocation.cpp:90-94
        }
        __syncthreads();
    }
}

void alloc_sequential_device(const double* d_arr, int K, int64_t n2, int64_t T, const double* d_resid_var,
                             const int* d_prior_off, const int* d_prior_idx, int* d_k, const int* d_order,
                             double baseline, double weight, int* d_votes, cudaStream_t st) {
    alloc_sequential_kernel<<<1, 1024, 0, st>>>(d_arr, K, n2, T, d_resid_var, d_prior_off, d_prior_idx, d_k, d_order,
                                                baseline, weight, d_votes);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

// ---------------------------------------------------------------- aggregate reallocation (allocate_per_obs = FALSE)
// R/scregclust.R:1678-1728, restated as written: scores n2 log(2 pi s2) + SS / s2 normalised like log-probabilities
// (:1685-1698), mixed with the log prior fractions (:1706-1725) and minimised (which.min, :1727); module_totals
// comes from the incoming configuration and is not updated while genes move (:1682).  One CTA per fit.
// rowSums() / sum() accumulate in long double in R; a compensated double sum stands in for it here.
__device__ __forceinline__ void acc2(double& hi, double& lo, double v) {
    const double t = hi + v;
    const double bp = t - hi;
    lo += (hi - (t - bp)) + (v - bp);
    hi = t;
}
__global__ void __launch_bounds__(256) alloc_aggregate_kernel(const double* __restrict__ sse_test, const double* __restrict__ var,
                                                              int K, int64_t T, double n2, const int* __restrict__ k_in,
                                                              const int* __restrict__ prior_off, const int* __restrict__ prior_idx,
                                                              const int* __restrict__ order, double baseline, double w,
                                                              double* __restrict__ mlp, int* __restrict__ k_out) {
    __shared__ double s_tot[RV_MAX_K];
    const int f = blockIdx.x;
    const double* ss = sse_test + (int64_t)f * (K + 1) * T;
    const double* vr = var + (int64_t)f * K * T;
    const int* kin = k_in + (int64_t)f * T;
    int* idx = k_out + (int64_t)f * T;
    double* lp = mlp + (int64_t)f * K * T;            // [t][j]
    for (int j = threadIdx.x; j < K; j += blockDim.x) s_tot[j] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0)
        for (int64_t t = 0; t < T; ++t) if (kin[t] >= 1 && kin[t] <= K) s_tot[kin[t] - 1] += 1.0;      // :1682
    auto mll = [&](int64_t t, int j) {
        const double v = vr[(int64_t)j * T + t];
        return n2 * log((2.0 * M_PI) * v) + ss[(int64_t)j * T + t] / v;                                // :1685-1688
    };
    for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
        idx[t] = kin[t];                                                                                // idx <- k (:1681)
        double mx = -INFINITY;
        for (int j = 0; j < K; ++j) { const double m = mll(t, j); if (m > mx) mx = m; }   // (a NaN score leaves the gene where it is)
        double hi = 0.0, lo = 0.0;
        for (int j = 0; j < K; ++j) acc2(hi, lo, exp(mll(t, j) - mx));
        const double lrs = log(hi + lo);
        for (int j = 0; j < K; ++j) lp[t * K + j] = (mll(t, j) - mx) - lrs;                             // :1691-1698
    }
    __syncthreads();
    auto pick = [&](int64_t g, const double* plp) {
        int bi = -1; double best = 0.0;
        for (int j = 0; j < K; ++j) {
            const double tot = (1.0 - w) * lp[g * K + j] + w * plp[j];                                  // :1722-1725
            if (tot == tot && (bi < 0 || tot < best)) { best = tot; bi = j; }                           // which.min
        }
        if (bi >= 0) idx[g] = bi + 1;
    };
    if (!prior_off) {
        // no prior entries: every gene sees the same prior fractions, so the order does not matter
        __shared__ double s_plp[RV_MAX_K];
        if (threadIdx.x == 0) {
            double hi = 0.0, lo = 0.0;
            for (int j = 0; j < K; ++j) acc2(hi, lo, 0.0 / (s_tot[j] > 0.0 ? s_tot[j] : 1.0) + baseline);
            const double ls = log(hi + lo);
            for (int j = 0; j < K; ++j) s_plp[j] = log(0.0 + baseline) - ls;
        }
        __syncthreads();
        for (int64_t t = threadIdx.x; t < T; t += blockDim.x) pick(t, s_plp);
        return;
    }
    if (threadIdx.x != 0) return;
    const int* ord = order + (int64_t)f * T;
    double frac[RV_MAX_K], plp[RV_MAX_K];
    for (int64_t oi = 0; oi < T; ++oi) {
        const int64_t g = ord[oi];
        for (int j = 0; j < K; ++j) frac[j] = 0.0;
        for (int e = prior_off[g]; e < prior_off[g + 1]; ++e) {                                         // :1706-1714
            const int m = idx[prior_idx[e]];
            if (m >= 1 && m <= K) frac[m - 1] += 1.0;
        }
        double hi = 0.0, lo = 0.0;
        for (int j = 0; j < K; ++j) {
            if (s_tot[j] > 0.0) frac[j] /= s_tot[j];                                                    // :1716-1718
            frac[j] += baseline;
            acc2(hi, lo, frac[j]);
        }
        const double ls = log(hi + lo);
        for (int j = 0; j < K; ++j) plp[j] = log(frac[j]) - ls;                                         // :1720
        pick(g, plp);
    }
}

void alloc_aggregate_device(const double* d_sse_test, const double* d_var, int K, int64_t T, int F, double n2,
                            const int* d_k_in, const int* d_prior_off, const int* d_prior_idx, const int* d_order,
                            double baseline, double weight, double* d_mlp, int* d_k_out, cudaStream_t st) {
    if (K > RV_MAX_K) throw Error(SCRC_ERR_UNSUPPORTED, "aggregate allocation: at most 64 modules are supported");
    alloc_aggregate_kernel<<<F, 256, 0, st>>>(d_sse_test, d_var, K, T, n2, d_k_in, d_prior_off, d_prior_idx, d_order,
                                              baseline, weight, d_mlp, d_k_out);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

void dropin_allocate(Device& dev, const double* sq_resid, int K, int64_t n2, int64_t T, const double* resid_var,
                     const int* prior_off, const int* prior_idx, const int* k_in, const int* order, double baseline,
                     double weight, int* k_out, int* votes_out) {
    cudaStream_t st = dev.stream;
    AllocScope as(st);
    if (K > RV_MAX_K) throw Error(SCRC_ERR_UNSUPPORTED, "allocate_into_modules: at most 64 modules are supported");
    if (weight != 0.0) validate_prior_inputs("allocate_into_modules", k_in, T, K, prior_off, prior_idx, order);
    DevBuf<double> d_var((size_t)T * K);
    SCRC_CUDA(cudaMemcpyAsync(d_var.p, resid_var, (size_t)T * K * 8, cudaMemcpyHostToDevice, st));
    DevBuf<int> d_k(T), d_votes(votes_out ? (size_t)T * K : 0);
    const int64_t per_gene = (int64_t)K * n2;
    const bool has_prior = (weight != 0.0);
    if (!has_prior) {
        // genes are independent (allocation.cpp:39-56 contributes nothing): stream the array in chunks
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(T, (int64_t)(1ull << 28) / std::max<int64_t>(per_gene, 1)));
        DevBuf<double> d_arr[2];
        d_arr[0].alloc((size_t)chunk * per_gene);
        d_arr[1].alloc((size_t)chunk * per_gene);
        cudaEvent_t done[2];
        SCRC_CUDA(cudaEventCreateWithFlags(&done[0], cudaEventDisableTiming));
        SCRC_CUDA(cudaEventCreateWithFlags(&done[1], cudaEventDisableTiming));
        int buf = 0;
        for (int64_t g0 = 0; g0 < T; g0 += chunk, buf ^= 1) {
            const int64_t ng = std::min(chunk, T - g0);
            SCRC_CUDA(cudaEventSynchronize(done[buf]));
            SCRC_CUDA(cudaMemcpyAsync(d_arr[buf].p, sq_resid + g0 * per_gene, (size_t)ng * per_gene * 8, cudaMemcpyHostToDevice, st));
            alloc_parallel_kernel<<<(unsigned)ng, 256, 0, st>>>(d_arr[buf].p, K, n2, g0, ng, d_var.p, T, d_k.p, d_votes.p);
            SCRC_LAUNCHED();
            SCRC_CUDA(cudaEventRecord(done[buf], st));
        }
        SCRC_CUDA(cudaGetLastError());
        SCRC_CUDA(cudaStreamSynchronize(st));
        cudaEventDestroy(done[0]); cudaEventDestroy(done[1]);
    } else {
        size_t free_b = 0, total_b = 0;
        SCRC_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = (size_t)T * per_gene * 8;
        if (need + (1ull << 30) > free_b)
            throw Error(SCRC_ERR_UNSUPPORTED, "allocate_into_modules with a prior needs the K x n2 x T array resident on the device; use the session API");
        DevBuf<double> d_arr((size_t)T * per_gene);
        SCRC_CUDA(cudaMemcpyAsync(d_arr.p, sq_resid, need, cudaMemcpyHostToDevice, st));
        std::vector<int> k0(T);
        for (int64_t t = 0; t < T; ++t) k0[t] = k_in[t] - 1;                 // allocation.cpp:23
        SCRC_CUDA(cudaMemcpyAsync(d_k.p, k0.data(), T * 4, cudaMemcpyHostToDevice, st));
        DevBuf<int> d_order(T), d_off, d_idx;
        SCRC_CUDA(cudaMemcpyAsync(d_order.p, order, T * 4, cudaMemcpyHostToDevice, st));
        if (prior_off) {
            d_off.alloc(T + 1);
            SCRC_CUDA(cudaMemcpyAsync(d_off.p, prior_off, (T + 1) * 4, cudaMemcpyHostToDevice, st));
            const int nidx = prior_off[T];
            d_idx.alloc(std::max(nidx, 1));
            if (nidx > 0) SCRC_CUDA(cudaMemcpyAsync(d_idx.p, prior_idx, (size_t)nidx * 4, cudaMemcpyHostToDevice, st));
        }
        alloc_sequential_kernel<<<1, 1024, 0, st>>>(d_arr.p, K, n2, T, d_var.p, prior_off ? d_off.p : nullptr,
                                                    prior_off ? d_idx.p : nullptr, d_k.p, d_order.p, baseline, weight,
                                                    d_votes.p);
        SCRC_LAUNCHED();
        SCRC_CUDA(cudaGetLastError());
        SCRC_CUDA(cudaStreamSynchronize(st));
    }
    SCRC_CUDA(cudaMemcpyAsync(k_out, d_k.p, T * 4, cudaMemcpyDeviceToHost, st));
    if (votes_out) SCRC_CUDA(cudaMemcpyAsync(votes_out, d_votes.p, (size_t)T * K * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    if (has_prior) for (int64_t t = 0; t < T; ++t) k_out[t] += 1;             // allocation.cpp:99
}

}  // namespace scrc
