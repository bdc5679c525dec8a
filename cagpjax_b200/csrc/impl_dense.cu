// Launchers of the dense right-hand-side kernels (one object per floating-point type).
#include "aux_kernels.cuh"
#include "host_common.cuh"
#include "ks_dense.cuh"

namespace cagp_host {

template <typename T>
int dense_fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, const void* V, int p, int64_t ldv, void* Y, int64_t ldy, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            cagp::DenseFwdArgs<T> a;
            a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
            a.xs2 = (const T*)xs2; a.ld2 = ld2; a.n = n;
            a.V = (const T*)V; a.ldv = ldv; a.p = p;
            a.Y = (T*)Y; a.ldy = ldy;
            const int64_t gx = (m + cagp::kDenseRows - 1) / cagp::kDenseRows;
            const int64_t gz = (p + cagp::kDenseCols - 1) / cagp::kDenseCols;
            const int64_t target = (int64_t)num_sms() * 2 * 4;
            int64_t gy = (target + gx * gz - 1) / (gx * gz);
            const int64_t max_split = (n + 2047) / 2048;
            if (gy > max_split) gy = max_split;
            if (gy < 1) gy = 1;
            if (gy > 65535) gy = 65535;
            a.jchunk = ((n + gy - 1) / gy + cagp::kDenseJC - 1) / cagp::kDenseJC * cagp::kDenseJC;
            gy = (n + a.jchunk - 1) / a.jchunk;
            if (gy > 1) {
                if (cudaMemset2DAsync(Y, (size_t)ldy * sizeof(T), 0, (size_t)p * sizeof(T), (size_t)m, stream) != cudaSuccess)
                    return fail(CAGP_ERR_CUDA, "memset Y failed");
            }
            const size_t smem = (cagp::ExpTab<T>::kSmemElems + (size_t)cagp::kDenseJC * (2 * D + cagp::kDenseKsStride + 2 * cagp::kDenseVsStride)) * sizeof(T);
            auto kern = cagp::ks_dense_fwd_kernel<T, D, FAM>;
            if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), cagp::kDenseThreads, smem, stream>>>(a);
            return check_launch("ks_dense_fwd_kernel");
        });
    });
}

namespace {
struct LsPlan {
    int64_t gx, gy, jchunk;
};
LsPlan plan_ls(int64_t m, int64_t n) {
    LsPlan pl;
    pl.gx = (m + 63) / 64;
    const int64_t target = (int64_t)num_sms() * 4 * 4;
    int64_t gy = (target + pl.gx - 1) / pl.gx;
    const int64_t max_split = (n + 63) / 64;
    if (gy > max_split) gy = max_split;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    pl.jchunk = ((n + gy - 1) / gy + 63) / 64 * 64;
    pl.gy = (n + pl.jchunk - 1) / pl.jchunk;
    return pl;
}
}  // namespace

template <typename T>
size_t dense_ls_ws_bytes(int64_t m, int64_t n, int d, bool ard) {
    const LsPlan pl = plan_ls(m, n);
    return (size_t)pl.gx * pl.gy * (ard ? cagp_padded_dim(d) : 1) * sizeof(T) + 256;
}

template <typename T>
int dense_ls_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                  int64_t ld2, int d, const void* V, int64_t ldv, const void* Ybar, int64_t ldyb, int p, void* ls_bar,
                  void* ws, size_t ws_bytes, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const bool ard = kd->n_lengthscale > 1;
    const LsPlan pl = plan_ls(m, n);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            auto launch = [&](auto ardc) {
                constexpr bool ARD = decltype(ardc)::value;
                constexpr int NLS = ARD ? D : 1;
                const int64_t nparts = pl.gx * pl.gy;
                if (ws_bytes < (size_t)nparts * NLS * sizeof(T)) return fail(CAGP_ERR_WORKSPACE, "dense ls workspace too small");
                cagp::DenseLsArgs<T> a;
                a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
                a.xs2 = (const T*)xs2; a.ld2 = ld2; a.n = n;
                a.Ybt = (const T*)Ybar; a.ldyb = ldyb;
                a.Vt = (const T*)V; a.ldv = ldv; a.p = p;
                a.ls_part = (T*)ws;
                a.jchunk = pl.jchunk;
                const size_t smem = (cagp::ExpTab<T>::kSmemElems + (size_t)32 * 2 * cagp::kDenseKsStride) * sizeof(T);
                auto kern = cagp::ks_dense_bwd_ls_kernel<T, D, FAM, ARD>;
                if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                kern<<<dim3((unsigned)pl.gx, (unsigned)pl.gy), 128, smem, stream>>>(a);
                if (int rc = check_launch("ks_dense_bwd_ls_kernel")) return rc;
                cagp::bwd_finalize_kernel<T><<<1, 256, 0, stream>>>(nullptr, 0, 0, nullptr, a.ls_part, nparts, NLS,
                                                                     kd->n_lengthscale, (const T*)kd->lengthscale,
                                                                     (T)cagp::family_grad_const(kd->family), (T*)ls_bar);
                return check_launch("bwd_finalize_kernel");
            };
            return ard ? launch(std::true_type{}) : launch(std::false_type{});
        });
    });
}

template int dense_fwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t,
                                       int64_t, int, const void*, int, int64_t, void*, int64_t, cudaStream_t);
template size_t dense_ls_ws_bytes<CAGP_REAL>(int64_t, int64_t, int, bool);
template int dense_ls_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t,
                                      int64_t, int, const void*, int64_t, const void*, int64_t, int, void*, void*,
                                      size_t, cudaStream_t);

}  // namespace cagp_host
