// Host side of the FP64 TN GEMM: TMA descriptors + DensePolicy launches.
#include "gemm_tn.cuh"
#include "prof.cuh"

#include <cstdlib>
#include <mutex>

namespace scrc {

std::atomic<unsigned long long> g_launch_count{0ull};
thread_local cudaStream_t g_alloc_stream = nullptr;
thread_local bool g_alloc_async = false;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    if (!fn) throw Error(-90, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

CUtensorMap make_tmap_f64(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_cols) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 8) % 16 != 0)
        throw Error(-91, "TMA operand must be 16-byte aligned with an even leading dimension");
    if (rows <= 0 || cols <= 0) throw Error(-92, "TMA operand with empty extent");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)GEMM_BK, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims,
                              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(-93, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
    return m;
}

CUtensorMap make_tmap_f64_plain(const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows,
                                int box_cols) {
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (ld * 8) % 16 != 0 || (box_rows * 8) % 16 != 0)
        throw Error(-91, "TMA operand must be 16-byte aligned with an even leading dimension");
    if (rows <= 0 || cols <= 0) throw Error(-92, "TMA operand with empty extent");
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box,
                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        throw Error(-93, "cuTensorMapEncodeTiled (plain) failed with CUresult " + std::to_string((int)r));
    return m;
}

template <int BM, int BN, int STAGES, int MODE>
static void launch_dense(const DenseArgs& a, int num_sms, cudaStream_t stream) {
    using Pol = DensePolicy<BM, BN, STAGES, MODE>;
    typename Pol::Params p;
    p.ta = make_tmap_f64(a.A, a.K, a.M, a.lda, BM);
    p.tb = make_tmap_f64(a.B, a.K, a.N, a.ldb, BN);
    p.C = a.C; p.ldc = a.ldc; p.M = a.M; p.N = a.N; p.K = a.K; p.alpha = a.alpha;
    p.rvec = a.rvec; p.cvec = a.cvec; p.D = a.D; p.ldd = a.ldd; p.shift = a.shift; p.cut = a.cut;
    constexpr int smem = GemmSmem<BM, BN, STAGES>::TOTAL;
    static SmemAttr attr;
    attr.ensure(gemm_tn_kernel<Pol>, smem);
    const int64_t tiles = ceil_div(a.M, BM) * ceil_div(a.N, BN);
    int grid = (int)std::min<int64_t>(tiles, num_sms);
    if (const char* e = getenv("SCRC_DEBUG_GEMM_GRID")) grid = (int)std::max<int64_t>(1, std::min<int64_t>(tiles, atoi(e)));
    {
        const double flops = (MODE == EPI_SYM_LOWER) ? (double)a.K * a.M * (a.M + 1.0) : 2.0 * a.K * (double)a.M * a.N;
        ProfScope ps(PROF_GEMM_DENSE, stream, flops);
        gemm_tn_kernel<Pol><<<grid, GEMM_THREADS, smem, stream>>>(p);
    }
    SCRC_CUDA(cudaGetLastError());
    SCRC_LAUNCHED();
}

template <int MODE>
static void dispatch_tile(const DenseArgs& a, int num_sms, cudaStream_t stream) {
    const int64_t big_tiles = ceil_div(a.M, 128) * ceil_div(a.N, 128);
    if (big_tiles >= num_sms) launch_dense<128, 128, 6, MODE>(a, num_sms, stream);
    else launch_dense<64, 64, 8, MODE>(a, num_sms, stream);
}

void gemm_tn_dense(int mode, const DenseArgs& a, int num_sms, cudaStream_t stream) {
    if (a.M <= 0 || a.N <= 0) return;
    if (a.K <= 0) throw Error(-94, "gemm_tn_dense: empty contraction");
    switch (mode) {
        case EPI_PLAIN: dispatch_tile<EPI_PLAIN>(a, num_sms, stream); break;
        case EPI_SYM_LOWER: dispatch_tile<EPI_SYM_LOWER>(a, num_sms, stream); break;
        case EPI_COR: dispatch_tile<EPI_COR>(a, num_sms, stream); break;
        case EPI_ROWDIV: dispatch_tile<EPI_ROWDIV>(a, num_sms, stream); break;
        case EPI_SUBFROM: dispatch_tile<EPI_SUBFROM>(a, num_sms, stream); break;
        default: throw Error(-95, "gemm_tn_dense: unknown epilogue mode");
    }
}

}  // namespace scrc
