// Launchers of the float64 DMMA contractions (gemm_dmma.cuh).  Compiled for double only.
#include "gemm_dmma.cuh"
#include "host_common.cuh"

namespace cagp_host {

namespace {

constexpr size_t kSyrkSmem = (size_t)cagp::kGemmStages * 2 * cagp::kGemmRowKElems * sizeof(double);
constexpr size_t kWmSmem = (size_t)cagp::kGemmStages * (cagp::kGemmRowKElems + cagp::kGemmKColElems) * sizeof(double);

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int tiles(int k) { return (k + cagp::kGemmTile - 1) / cagp::kGemmTile; }

}  // namespace

// number of m-splits of the Gram kernel: about 8 CTAs per SM over all lower-triangular tiles, each split at least
// 512 columns long and a multiple of the k-chunk
int gram_splits(int64_t m, int k, int64_t* chunk_out) {
    const int T = tiles(k), ntiles = T * (T + 1) / 2;
    int64_t splits = ((int64_t)num_sms() * 9 + ntiles / 2) / ntiles;
    const int64_t max_splits = (m + 511) / 512;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int64_t chunk = (m + splits - 1) / splits;
    chunk = (chunk + cagp::kGemmKC - 1) / cagp::kGemmKC * cagp::kGemmKC;
    splits = (m + chunk - 1) / chunk;
    if (chunk_out) *chunk_out = chunk;
    return (int)splits;
}

size_t gram_ws_bytes_f64(int64_t m, int k) { return (size_t)gram_splits(m, k, nullptr) * k * k * sizeof(double); }

int gram_mtm_f64(const double* Mt, int64_t ldm, int64_t m, int k, double* B, void* ws, size_t ws_bytes, cudaStream_t stream) {
    cagp::SyrkArgs a;
    a.Mt = Mt; a.ldm = ldm; a.m = m; a.k = k; a.ws = (double*)ws;
    const int splits = gram_splits(m, k, &a.chunk);
    if (!ws || ws_bytes < (size_t)splits * k * k * sizeof(double)) return fail(CAGP_ERR_WORKSPACE, "gram_mtm: workspace too small");
    const int T = tiles(k);
    const dim3 grid((unsigned)(T * (T + 1) / 2), (unsigned)splits);
    const bool v16 = aligned16(Mt) && (ldm % 2 == 0);
    auto kern = v16 ? cagp::gram_syrk_kernel<true> : cagp::gram_syrk_kernel<false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSyrkSmem) != cudaSuccess)
        return fail(CAGP_ERR_CUDA, "gram_syrk_kernel: shared memory opt-in failed");
    kern<<<grid, cagp::kGemmThreads, kSyrkSmem, stream>>>(a);
    if (int rc = check_launch("gram_syrk_kernel")) return rc;
    cagp::gram_reduce_kernel<<<dim3((k + 31) / 32, (k + 7) / 8), dim3(32, 8), 0, stream>>>((const double*)ws, splits, k, B);
    return check_launch("gram_reduce_kernel");
}

static int launch_wm(int epi, const cagp::GemmWmArgs& a, bool v16, cudaStream_t stream) {
    const int64_t nit = (a.m + cagp::kGemmTile - 1) / cagp::kGemmTile;
    const int64_t nblocks = nit * a.nbt;
    if (nblocks > 0x7fffffffLL) return fail(CAGP_ERR_UNSUPPORTED, "gemm_wm: too many tiles");
    void (*kern)(const cagp::GemmWmArgs);
    if (epi == cagp::GEMM_COT) kern = v16 ? cagp::gemm_wm_kernel<cagp::GEMM_COT, true> : cagp::gemm_wm_kernel<cagp::GEMM_COT, false>;
    else if (epi == cagp::GEMM_COT_PEER) kern = v16 ? cagp::gemm_wm_kernel<cagp::GEMM_COT_PEER, true> : cagp::gemm_wm_kernel<cagp::GEMM_COT_PEER, false>;
    else kern = v16 ? cagp::gemm_wm_kernel<cagp::GEMM_MOM, true> : cagp::gemm_wm_kernel<cagp::GEMM_MOM, false>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWmSmem) != cudaSuccess)
        return fail(CAGP_ERR_CUDA, "gemm_wm_kernel: shared memory opt-in failed");
    kern<<<(unsigned)nblocks, cagp::kGemmThreads, kWmSmem, stream>>>(a);
    return check_launch("gemm_wm_kernel");
}

int assemble_cotangent_f64(const double* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total, int k,
                           const double* W, const double* Abar, const double* cbar, const double* s, const double* yt,
                           double* Gt, int64_t ldg, cudaStream_t stream) {
    cagp::GemmWmArgs a{};
    a.W = W; a.k = k; a.Mt = Mt; a.ldm = ldm; a.m = m; a.nbt = tiles(k);
    a.Gt = Gt; a.ldg = ldg; a.Abar = Abar; a.cbar = cbar; a.s = s; a.yt = yt; a.row_offset = row_offset;
    a.lay = cagp::BlockLayout{n_total, k, n_total / k};
    const bool v16 = aligned16(Mt) && aligned16(W) && aligned16(Gt) && ldm % 2 == 0 && ldg % 2 == 0 && k % 2 == 0;
    return launch_wm(cagp::GEMM_COT, a, v16, stream);
}

int assemble_cotangent_peer_f64(const double* Mt, int64_t ldm, int64_t m, int64_t row_offset, int64_t n_total, int k,
                                const double* W, const double* Abar, const double* cbar, const double* s, const double* yt,
                                void* const* buffers, const int32_t* block_begin, int world, int me, int64_t ld,
                                cudaStream_t stream) {
    cagp::GemmWmArgs a{};
    a.W = W; a.k = k; a.Mt = Mt; a.ldm = ldm; a.m = m; a.nbt = tiles(k);
    a.Abar = Abar; a.cbar = cbar; a.s = s; a.yt = yt; a.row_offset = row_offset;
    a.lay = cagp::BlockLayout{n_total, k, n_total / k};
    a.peer_ld = ld; a.world = world; a.me = me;
    bool v16 = aligned16(Mt) && aligned16(W) && ldm % 2 == 0 && ld % 2 == 0 && k % 2 == 0 && row_offset % 2 == 0;
    for (int h = 0; h < world; ++h) {
        a.peer[h] = (double*)buffers[h];
        a.block_begin[h] = block_begin[h];
        v16 = v16 && aligned16(buffers[h]);
    }
    a.block_begin[world] = block_begin[world];
    return launch_wm(cagp::GEMM_COT_PEER, a, v16, stream);
}

size_t predict_ws_bytes_f64(int64_t m, int k) { return (size_t)2 * tiles(k) * m * sizeof(double); }

int predict_moments_f64(const double* Mt, int64_t ldm, int64_t m, int k, const double* w, const double* Pinv,
                        const double* scalars, double* mean, double* var, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (ws_bytes < predict_ws_bytes_f64(m, k)) return fail(CAGP_ERR_WORKSPACE, "predict_moments: workspace too small");
    cagp::GemmWmArgs a{};
    a.W = Pinv; a.k = k; a.Mt = Mt; a.ldm = ldm; a.m = m; a.nbt = tiles(k);
    a.w = w; a.pv = (double*)ws; a.pm = (double*)ws + (size_t)a.nbt * m;
    const bool v16 = aligned16(Mt) && aligned16(Pinv) && ldm % 2 == 0 && k % 2 == 0;
    if (int rc = launch_wm(cagp::GEMM_MOM, a, v16, stream)) return rc;
    cagp::predict_finalize_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(a.pv, a.pm, a.nbt, m, scalars, mean, var);
    return check_launch("predict_finalize_kernel");
}

}  // namespace cagp_host
