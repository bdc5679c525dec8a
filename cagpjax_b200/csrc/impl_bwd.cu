// Launcher of the rectangular backward kernel (one object per floating-point type).
#include "aux_kernels.cuh"
#include "host_common.cuh"
#include "ks_blocksparse.cuh"

namespace cagp_host {

namespace {
struct BwdPlan {
    int rows, tiles_main, tiles_last, n_ichunks;
    int64_t ichunk, gx;
};
constexpr int kBwdThreads = 128;

BwdPlan plan_bwd(int64_t m, int64_t n, int k, int dp) {
    cagp::BlockLayout lay{n, k, n / k};
    BwdPlan pl;
    pl.rows = choose_rows(lay.bs, kBwdThreads, max_rows(dp));
    const int64_t cap = (int64_t)kBwdThreads * pl.rows;
    pl.tiles_main = (int)((lay.bs + cap - 1) / cap);
    if (pl.tiles_main < 1) pl.tiles_main = 1;
    pl.tiles_last = (int)((lay.len(k - 1) + cap - 1) / cap);
    if (pl.tiles_last < 1) pl.tiles_last = 1;
    pl.gx = (int64_t)(k - 1) * pl.tiles_main + pl.tiles_last;
    const int64_t target = (int64_t)num_sms() * 4 * 32;
    int64_t gy = (target + pl.gx - 1) / pl.gx;
    const int64_t max_chunks = (m + 2047) / 2048;  // at least 2048 streamed points per CTA
    if (gy > max_chunks) gy = max_chunks;
    if (gy > 64) gy = 64;
    if (gy < 1) gy = 1;
    pl.ichunk = (m + gy - 1) / gy;
    pl.n_ichunks = (int)((m + pl.ichunk - 1) / pl.ichunk);
    return pl;
}
}  // namespace

template <typename T>
size_t bwd_ws_bytes(int64_t m, int64_t n, int d, int k, bool ard) {
    const int dp = cagp_padded_dim(d);
    BwdPlan pl = plan_bwd(m, n, k, dp);
    const size_t nls = ard ? dp : 1;
    return ((size_t)pl.n_ichunks * (size_t)n + (size_t)pl.n_ichunks * (size_t)pl.gx * nls) * sizeof(T) + 256;
}

template <typename T>
int bwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
             int64_t ld2, int d, const void* s, int k, const void* Gt, int64_t ldg, void* s_bar, void* ls_bar,
             void* ws, size_t ws_bytes, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const bool ard = kd->n_lengthscale > 1;
    const BwdPlan pl = plan_bwd(m, n, k, dp);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            return dispatch_rows<D>(pl.rows, [&](auto rc) {
                constexpr int R = decltype(rc)::value;
                constexpr int NT = kBwdThreads;
                constexpr int JC = ColsPerStage<D>::value;
                const int nls = ard ? D : 1;
                const size_t need = ((size_t)pl.n_ichunks * n + (size_t)pl.n_ichunks * pl.gx * nls) * sizeof(T);
                if (ws_bytes < need) return fail(CAGP_ERR_WORKSPACE, "ks bwd workspace: need %zu bytes, got %zu", need, ws_bytes);
                cagp::BwdArgs<T> a;
                a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
                a.xs2 = (const T*)xs2; a.ld2 = ld2;
                a.s = (const T*)s;
                a.lay = cagp::BlockLayout{n, k, n / k};
                a.Gt = (const T*)Gt; a.ldg = ldg;
                a.sbar_part = (T*)ws;
                a.ls_part = (T*)ws + (size_t)pl.n_ichunks * n;
                a.tiles_main = pl.tiles_main;
                a.ichunk = pl.ichunk;
                a.bbox = nullptr;
                const size_t smem = (cagp::ExpTab<T>::kSmemElems + (size_t)JC * cagp::EntrySize<D>::value) * sizeof(T);
                dim3 grid((unsigned)pl.gx, (unsigned)pl.n_ichunks);
                if (ard) {
                    auto kern = cagp::ks_bwd_kernel<T, D, FAM, true, R, NT, JC>;
                    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<grid, NT, smem, stream>>>(a);
                } else {
                    auto kern = cagp::ks_bwd_kernel<T, D, FAM, false, R, NT, JC>;
                    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<grid, NT, smem, stream>>>(a);
                }
                if (int rc2 = check_launch("ks_bwd_kernel")) return rc2;
                // deterministic reduction of the partials
                const int64_t nparts = (int64_t)pl.n_ichunks * pl.gx;
                cagp::bwd_finalize_kernel<T><<<dim3((unsigned)((n + 255) / 256) + 1), 256, 0, stream>>>(
                    a.sbar_part, pl.n_ichunks, n, (T*)s_bar, a.ls_part, nparts, nls, kd->n_lengthscale,
                    (const T*)kd->lengthscale, (T)cagp::family_grad_const(kd->family), (T*)ls_bar);
                return check_launch("bwd_finalize_kernel");
            });
        });
    });
}

template size_t bwd_ws_bytes<CAGP_REAL>(int64_t, int64_t, int, int, bool);
template int bwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t, int64_t,
                                 int, const void*, int, const void*, int64_t, void*, void*, void*, size_t, cudaStream_t);

}  // namespace cagp_host
