// Initial clustering of scregclust(): kmeanspp() (R/utils.R:351-375) = k-means++ seeding (kmeanspp_init,
// R/utils.R:275-322) followed by stats::kmeans (Hartigan & Wong 1979, AS 136) for every initialisation, the
// clustering with the smallest tot.withinss wins.  SURVEY.md section 8 (f3).
//
// What runs where.  R's RNG stays with the caller: it supplies, per initialisation, the value of sample(n, 1) and the
// K - 1 unif_rand() draws that the K - 1 sample.int(prob =) calls consume (INTEGRATION.md shows the two R lines).  The
// device computes the distances of every point to the newest centre of every initialisation (sequential sums in
// R_euclidean's order), the host turns them into R's sampling decision (log / exp / long-double sum / FixupProb /
// revsort / cumulative walk of ProbSampleNoReplace -- O(n log n) integer-and-compare work per draw).  Hartigan-Wong is
// a sequential point-transfer algorithm: every decision uses the centres as the previous point left them, and its
// distances are sequential sums whose order decides near-ties.  It is therefore run as published -- one CTA per
// initialisation (they are independent: 50 of them keep 50 SMs busy), points in order, the K candidate distances of a
// point on K lanes (each lane a sequential sum over the coordinates, read from a transposed copy so that the point is
// contiguous and the centres are coalesced across lanes), centre updates spread over the CTA.
#include <algorithm>
#include <cmath>
#include <vector>

#include "../../include/scregclust_b200.h"
#include "dmat.cuh"
#include "misc.cuh"

namespace scrc {

constexpr int KM_MAX_K = 64;
constexpr int KM_TPB = 128;
constexpr double KM_BIG = 1.0e30;

// mind[r][i] = min(mind[r][i], dist(i, newest[r])) with dist = sqrt(sum_j (x[i,j] - x[c,j])^2), j increasing
// (R's dist(): R_euclidean).  Consecutive threads read consecutive points of a column.
__global__ void __launch_bounds__(256) kpp_dist_kernel(const double* __restrict__ x, int64_t ld, int n, int p,
                                                       const int* __restrict__ newest, double* mind) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int r = blockIdx.y;
    if (i >= n) return;
    const int c = newest[r];
    double s = 0.0;
    for (int j = 0; j < p; ++j) {
        const double dev = x[i + (int64_t)j * ld] - x[c + (int64_t)j * ld];
        s += dev * dev;
    }
    const double d = sqrt(s);
    double* mp = mind + (int64_t)r * n + i;
    if (d < *mp) *mp = d;
}
__global__ void fill_inf_kernel(double* x, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) x[i] = INFINITY;
}

struct KmParams {
    const double* x;    int64_t ldx;    // m x p, column-major (point index contiguous)
    const double* xT;   int64_t ldt;    // p x m, coordinate contiguous per point
    int m, p, K, iter_max;
    const int* init;                    // [R][K] 0-based row indices of the initial centres
    double* C;                          // [R][p][K] centres (cluster index contiguous)
    int* ic1; int* ic2;                 // [R][m]
    double* D;                          // [R][m]
    double* wss;                        // [R][K]
    int* ifault;                        // [R]
};

// Hartigan & Wong (1979), AS 136: one CTA per initialisation.  Every thread executes the scalar control flow (the state
// it branches on lives in shared memory or in registers that all threads compute identically), lanes 0..K-1 own one
// cluster each for the distance evaluations, all threads share the centre updates.
__global__ void __launch_bounds__(KM_TPB) kmeans_hw_kernel(KmParams P) {
    extern __shared__ double a_s[];                 // coordinates of the current point
    __shared__ double sd[KM_MAX_K], an1[KM_MAX_K], an2[KM_MAX_K];
    __shared__ int nc[KM_MAX_K], ncp[KM_MAX_K], itran[KM_MAX_K], live[KM_MAX_K];
    const int tid = threadIdx.x, r = blockIdx.x;
    const int m = P.m, p = P.p, K = P.K;
    double* C = P.C + (int64_t)r * p * K;
    int* ic1 = P.ic1 + (int64_t)r * m;
    int* ic2 = P.ic2 + (int64_t)r * m;
    double* D = P.D + (int64_t)r * m;

    auto load_point = [&](int i) {
        for (int j = tid; j < p; j += KM_TPB) a_s[j] = P.xT[j + (int64_t)i * P.ldt];
        __syncthreads();
    };
    // squared distance of the loaded point to cluster `tid`, coordinates in increasing order (the early exits of the
    // published loops only skip candidates whose partial sum already exceeds the bound: same decisions)
    auto dist_lane = [&]() {
        double acc = 0.0;
        for (int j = 0; j < p; ++j) {
            const double t = a_s[j] - C[(int64_t)j * K + tid];
            acc += t * t;
        }
        return acc;
    };
    auto transfer = [&](int i, int l1, int l2) {     // move point i from l1 to l2: centres, counts, AN1 / AN2
        const double al1 = (double)nc[l1], alw = al1 - 1.0, al2 = (double)nc[l2], alt = al2 + 1.0;
        for (int j = tid; j < p; j += KM_TPB) {
            const double a = a_s[j];
            C[(int64_t)j * K + l1] = (C[(int64_t)j * K + l1] * al1 - a) / alw;
            C[(int64_t)j * K + l2] = (C[(int64_t)j * K + l2] * al2 + a) / alt;
        }
        __syncthreads();
        if (tid == 0) {
            nc[l1] -= 1; nc[l2] += 1;
            an2[l1] = alw / al1;
            an1[l1] = alw > 1.0 ? alw / (alw - 1.0) : KM_BIG;
            an1[l2] = alt / al2;
            an2[l2] = alt / (alt + 1.0);
            ic1[i] = l2; ic2[i] = l1;
        }
        __syncthreads();
    };

    // ---- initial centres = the seeded points
    for (int e = tid; e < p * K; e += KM_TPB) {
        const int j = e / K, l = e % K;
        C[e] = P.xT[j + (int64_t)P.init[r * K + l] * P.ldt];
    }
    __syncthreads();
    // ---- for each point its two closest centres
    for (int i = 0; i < m; ++i) {
        load_point(i);
        if (tid < K) sd[tid] = dist_lane();
        __syncthreads();
        if (tid == 0) {
            int i1 = 0, i2 = 1;
            double d1 = sd[0], d2 = sd[1];
            if (d1 > d2) { i1 = 1; i2 = 0; const double t = d1; d1 = d2; d2 = t; }
            for (int l = 2; l < K; ++l) {
                const double db = sd[l];
                if (db >= d2) continue;
                if (db >= d1) { d2 = db; i2 = l; }
                else { d2 = d1; i2 = i1; d1 = db; i1 = l; }
            }
            ic1[i] = i1; ic2[i] = i2;
        }
        __syncthreads();
    }
    // ---- centres = means of their points (each coordinate summed over the points in increasing order)
    if (tid < K) nc[tid] = 0;
    __syncthreads();
    if (tid == 0) for (int i = 0; i < m; ++i) nc[ic1[i]] += 1;
    for (int j = tid; j < p; j += KM_TPB) {
        for (int l = 0; l < K; ++l) C[(int64_t)j * K + l] = 0.0;
        for (int i = 0; i < m; ++i) C[(int64_t)j * K + ic1[i]] += P.x[i + (int64_t)j * P.ldx];
    }
    __syncthreads();
    bool empty = false;
    for (int l = 0; l < K; ++l) empty |= (nc[l] == 0);
    if (empty) {                                    // IFAULT = 1: R stops with "empty cluster"
        if (tid == 0) P.ifault[r] = 1;
        return;
    }
    for (int j = tid; j < p; j += KM_TPB)
        for (int l = 0; l < K; ++l) C[(int64_t)j * K + l] /= (double)nc[l];
    if (tid < K) {
        const double aa = (double)nc[tid];
        an2[tid] = aa / (aa + 1.0);
        an1[tid] = aa > 1.0 ? aa / (aa - 1.0) : KM_BIG;
        itran[tid] = 1;
        ncp[tid] = -1;
    }
    __syncthreads();

    int indx = 0, ifault = 2;
    const long long imaxqtr = 50LL * m < 2147483647LL ? 50LL * m : 2147483647LL;
    for (int it = 0; it < P.iter_max; ++it) {
        // ================= optimal-transfer stage
        if (tid < K && itran[tid] == 1) live[tid] = m;
        __syncthreads();
        bool stop = false;
        for (int i = 0; i < m; ++i) {
            indx += 1;
            const int l1 = ic1[i];
            if (nc[l1] != 1) {
                load_point(i);
                const int ll = ic2[i];
                const bool l1_dead = (i + 1 >= live[l1]);
                const bool redo_l1 = (ncp[l1] != 0);
                if (tid < K) {
                    const bool need = (tid == l1) ? redo_l1 : (tid == ll || !(l1_dead && i + 1 >= live[tid]));
                    if (need) sd[tid] = dist_lane();
                }
                __syncthreads();
                // scalar decision, replicated by every thread (reads shared state only)
                const double di = redo_l1 ? sd[l1] * an1[l1] : D[i];
                int l2 = ll;
                double r2 = sd[ll] * an2[ll];
                for (int l = 0; l < K; ++l) {
                    if ((l1_dead && i + 1 >= live[l]) || l == l1 || l == ll) continue;
                    const double rr = r2 / an2[l];
                    const double dc = sd[l];
                    if (dc >= rr) continue;
                    r2 = dc * an2[l]; l2 = l;
                }
                __syncthreads();                    // everyone has read sd / D before they change
                if (tid == 0 && redo_l1) D[i] = di;
                if (r2 >= di) {
                    if (tid == 0) ic2[i] = l2;
                    __syncthreads();
                } else {
                    indx = 0;
                    if (tid == 0) { live[l1] = m + i + 1; live[l2] = m + i + 1; ncp[l1] = i + 1; ncp[l2] = i + 1; }
                    transfer(i, l1, l2);
                }
            }
            if (indx == m) { stop = true; break; }
        }
        if (!stop) {
            __syncthreads();
            if (tid < K) { itran[tid] = 0; live[tid] -= m; }
            __syncthreads();
        }
        if (indx == m) { ifault = 0; break; }
        // ================= quick-transfer stage
        int icoun = 0;
        long long istep = 0;
        bool capped = false, done = false;
        while (!done) {
            for (int i = 0; i < m; ++i) {
                icoun += 1; istep += 1;
                if (istep >= imaxqtr) { capped = true; done = true; break; }
                const int l1 = ic1[i], l2 = ic2[i];
                if (nc[l1] != 1) {
                    const bool redo_l1 = (istep <= (long long)ncp[l1]);
                    const bool try_l2 = (istep < (long long)ncp[l1] || istep < (long long)ncp[l2]);
                    if (redo_l1 || try_l2) {
                        load_point(i);
                        if (tid < K && ((tid == l1 && redo_l1) || (tid == l2 && try_l2))) sd[tid] = dist_lane();
                        __syncthreads();
                        const double di = redo_l1 ? sd[l1] * an1[l1] : D[i];
                        const bool move = try_l2 && (sd[l2] < di / an2[l2]);
                        __syncthreads();
                        if (tid == 0 && redo_l1) D[i] = di;
                        if (move) {
                            icoun = 0; indx = 0;
                            if (tid == 0) {
                                itran[l1] = 1; itran[l2] = 1;
                                ncp[l1] = (int)(istep + m); ncp[l2] = (int)(istep + m);
                            }
                            transfer(i, l1, l2);
                        } else {
                            __syncthreads();
                        }
                    }
                }
                if (icoun == m) { done = true; break; }
            }
        }
        if (capped) { ifault = 4; break; }
        if (K == 2) { ifault = 0; break; }
        __syncthreads();
        if (tid < K) ncp[tid] = 0;
        __syncthreads();
    }
    // ---- within-cluster sums of squares about the recomputed means (coordinates outer, points inner)
    __syncthreads();
    for (int j = tid; j < p; j += KM_TPB) {
        for (int l = 0; l < K; ++l) C[(int64_t)j * K + l] = 0.0;
        for (int i = 0; i < m; ++i) C[(int64_t)j * K + ic1[i]] += P.x[i + (int64_t)j * P.ldx];
        for (int l = 0; l < K; ++l) C[(int64_t)j * K + l] /= (double)nc[l];
    }
    __syncthreads();
    if (tid < K) {
        double w = 0.0;
        for (int j = 0; j < p; ++j) {
            const double c = C[(int64_t)j * K + tid];
            for (int i = 0; i < m; ++i)
                if (ic1[i] == tid) { const double da = P.x[i + (int64_t)j * P.ldx] - c; w += da * da; }
        }
        P.wss[r * K + tid] = w;
    }
    if (tid == 0) P.ifault[r] = ifault;
}

// ---------------------------------------------------------------- host side of the seeding: R's sampling decision
static void r_revsort(std::vector<double>& a, std::vector<int>& ib) {     // src/main/sort.c: revsort
    const int n = (int)a.size();
    if (n <= 1) return;
    double* A = a.data() - 1;
    int* B = ib.data() - 1;
    int l = (n >> 1) + 1, ir = n;
    for (;;) {
        double ra; int ii;
        if (l > 1) {
            l -= 1; ra = A[l]; ii = B[l];
        } else {
            ra = A[ir]; ii = B[ir];
            A[ir] = A[1]; B[ir] = B[1];
            if (--ir == 1) { A[1] = ra; B[1] = ii; return; }
        }
        int i = l, j = l << 1;
        while (j <= ir) {
            if (j < ir && A[j] > A[j + 1]) ++j;
            if (ra > A[j]) { A[i] = A[j]; B[i] = B[j]; j += (i = j); }
            else j = ir + 1;
        }
        A[i] = ra; B[i] = ii;
    }
}
// sample.int(length(ps), 1, prob = ps) given the unif_rand() it draws: FixupProb + ProbSampleNoReplace (random.c)
static int r_sample_prob_1(std::vector<double>& p, double u) {
    double sum = 0.0;
    for (double v : p) if (v > 0.0) sum += v;
    for (double& v : p) v /= sum;
    std::vector<int> perm(p.size());
    for (size_t i = 0; i < p.size(); ++i) perm[i] = (int)i + 1;
    r_revsort(p, perm);
    const double rT = 1.0 * u;
    double mass = 0.0;
    const int n1 = (int)p.size() - 1;
    int j = 0;
    for (; j < n1; ++j) { mass += p[j]; if (rT <= mass) break; }
    return perm[j];
}

void kmeanspp_device(cudaStream_t st, const double* x_dev, int64_t ldx, int64_t n, int64_t p, int K, int n_init,
                     int n_max_iter, const int* first_centers, const double* uniforms, int* cluster_out,
                     double* tot_withinss_out, int* ifault_out) {
    if (K < 2 || K > KM_MAX_K) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_kmeanspp: n_cluster must be in 2..64");
    if (n < K || p <= 0 || n_init <= 0 || n_max_iter <= 0) throw Error(SCRC_ERR_ARG, "scrc_kmeanspp: invalid dimensions");
    if (n > 0x7fffffffLL / 64 || p > 0x7fffffffLL / KM_MAX_K) throw Error(SCRC_ERR_UNSUPPORTED, "scrc_kmeanspp: matrix too large");
    const int R = n_init;
    for (int r = 0; r < R; ++r)
        if (first_centers[r] < 1 || first_centers[r] > n) throw Error(SCRC_ERR_ARG, "scrc_kmeanspp: first_centers must be in 1..n");
    // ---- seeding (R/utils.R:275-322)
    std::vector<int> centers((size_t)R * K, -1), newest(R);
    for (int r = 0; r < R; ++r) { centers[(size_t)r * K] = first_centers[r] - 1; newest[r] = first_centers[r] - 1; }
    DevBuf<double> mind((size_t)R * n);
    DevBuf<int> d_newest(R);
    fill_inf_kernel<<<(unsigned)ceil_div((int64_t)R * n, 256), 256, 0, st>>>(mind.p, (int64_t)R * n);
    SCRC_LAUNCHED();
    std::vector<double> h_mind((size_t)R * n), ps;
    std::vector<int> remaining;
    std::vector<char> is_center((size_t)n);
    for (int round = 1; round < K; ++round) {
        SCRC_CUDA(cudaMemcpyAsync(d_newest.p, newest.data(), R * 4, cudaMemcpyHostToDevice, st));
        kpp_dist_kernel<<<dim3((unsigned)ceil_div(n, 256), R), 256, 0, st>>>(x_dev, ldx, (int)n, (int)p, d_newest.p, mind.p);
        SCRC_LAUNCHED();
        SCRC_CUDA(cudaMemcpyAsync(h_mind.data(), mind.p, h_mind.size() * 8, cudaMemcpyDeviceToHost, st));
        SCRC_CUDA(cudaStreamSynchronize(st));
        for (int r = 0; r < R; ++r) {
            std::fill(is_center.begin(), is_center.end(), 0);
            for (int q = 0; q < round; ++q) is_center[centers[(size_t)r * K + q]] = 1;
            remaining.clear(); ps.clear();
            double mx = -INFINITY;
            for (int64_t i = 0; i < n; ++i) {
                if (is_center[i]) continue;
                const double d = h_mind[(size_t)r * n + i];
                const double lw = std::log(d * d);                      // log(apply(..., 1, min)^2)
                remaining.push_back((int)i); ps.push_back(lw);
                if (lw > mx) mx = lw;
            }
            long double tot = 0.0L;                                     // R's sum(): long double accumulation
            for (double& v : ps) { v = std::exp(v - mx); tot += (long double)v; }
            const double totd = (double)tot;
            for (double& v : ps) v /= totd;
            const int pick = r_sample_prob_1(ps, uniforms[(size_t)r * (K - 1) + (round - 1)]);
            centers[(size_t)r * K + round] = remaining[pick - 1];
            newest[r] = remaining[pick - 1];
        }
    }
    // ---- Hartigan-Wong for every initialisation
    DMat XT(p, n);
    if (XT.ld != p) SCRC_CUDA(cudaMemsetAsync(XT.data(), 0, XT.bytes(), st));
    transpose(x_dev, ldx, n, p, XT.data(), XT.ld, st);
    DevBuf<int> d_init((size_t)R * K), ic1((size_t)R * n), ic2((size_t)R * n), d_fault(R);
    DevBuf<double> C((size_t)R * p * K), D((size_t)R * n), wss((size_t)R * K);
    SCRC_CUDA(cudaMemcpyAsync(d_init.p, centers.data(), (size_t)R * K * 4, cudaMemcpyHostToDevice, st));
    SCRC_CUDA(cudaMemsetAsync(D.p, 0, (size_t)R * n * 8, st));
    SCRC_CUDA(cudaMemsetAsync(wss.p, 0, (size_t)R * K * 8, st));
    KmParams P;
    P.x = x_dev; P.ldx = ldx; P.xT = XT.data(); P.ldt = XT.ld; P.m = (int)n; P.p = (int)p; P.K = K; P.iter_max = n_max_iter;
    P.init = d_init.p; P.C = C.p; P.ic1 = ic1.p; P.ic2 = ic2.p; P.D = D.p; P.wss = wss.p; P.ifault = d_fault.p;
    const size_t smem = (size_t)p * 8;
    static SmemAttr attr;
    attr.ensure(kmeans_hw_kernel, smem);
    kmeans_hw_kernel<<<R, KM_TPB, smem, st>>>(P);
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
    std::vector<double> h_wss((size_t)R * K);
    std::vector<int> h_fault(R);
    SCRC_CUDA(cudaMemcpyAsync(h_wss.data(), wss.p, h_wss.size() * 8, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaMemcpyAsync(h_fault.data(), d_fault.p, R * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    int best = -1;
    double best_tot = INFINITY;
    for (int r = 0; r < R; ++r) {
        if (ifault_out) ifault_out[r] = h_fault[r];
        if (h_fault[r] == 1)      // stats::kmeans stops here: "empty cluster: try a better set of initial centers"
            throw Error(SCRC_ERR_ARG, "scrc_kmeanspp: empty cluster in initialisation " + std::to_string(r + 1) +
                                          ": try a better set of initial centers");
        long double t = 0.0L;
        for (int l = 0; l < K; ++l) t += (long double)h_wss[(size_t)r * K + l];       // tot.withinss = sum(wss)
        if (tot_withinss_out) tot_withinss_out[r] = (double)t;
        if ((double)t < best_tot) { best_tot = (double)t; best = r; }                  // which.min: first minimum
    }
    if (best < 0) throw Error(SCRC_ERR_INTERNAL, "scrc_kmeanspp: no finite objective");
    std::vector<int> h_cl(n);
    SCRC_CUDA(cudaMemcpyAsync(h_cl.data(), ic1.p + (size_t)best * n, n * 4, cudaMemcpyDeviceToHost, st));
    SCRC_CUDA(cudaStreamSynchronize(st));
    for (int64_t i = 0; i < n; ++i) cluster_out[i] = h_cl[i] + 1;
}

}  // namespace scrc
