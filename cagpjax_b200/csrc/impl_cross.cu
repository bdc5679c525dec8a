// Launchers of the materialised cross-covariance kernels (one object per floating-point type).
#include "aux_kernels.cuh"
#include "cross_cov.cuh"
#include "host_common.cuh"

namespace cagp_host {

namespace {
struct CrossPlan {
    int64_t gx, gy, rows_per_cta;
};
CrossPlan plan_cross(int64_t m, int64_t n) {
    CrossPlan pl;
    pl.gx = (n + cagp::kCrossThreads - 1) / cagp::kCrossThreads;
    const int64_t target = (int64_t)num_sms() * 8;  // 8 CTAs of 128 threads per SM
    int64_t gy = (target + pl.gx - 1) / pl.gx;
    const int64_t max_split = (m + cagp::kCrossRowTile - 1) / cagp::kCrossRowTile;
    if (gy > max_split) gy = max_split;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    pl.rows_per_cta = ((m + gy - 1) / gy + cagp::kCrossRowTile - 1) / cagp::kCrossRowTile * cagp::kCrossRowTile;
    pl.gy = (m + pl.rows_per_cta - 1) / pl.rows_per_cta;
    return pl;
}
template <typename T, int D>
size_t cross_smem() {
    return (cagp::ExpTab<T>::kSmemElems + (size_t)D * cagp::kCrossRowTile) * sizeof(T);
}
}  // namespace

template <typename T>
int cross_fwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, void* C, int64_t ldc, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const CrossPlan pl = plan_cross(m, n);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            cagp::CrossArgs<T> a{};
            a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
            a.xs2 = (const T*)xs2; a.ld2 = ld2; a.n = n;
            a.C = (T*)C; a.ldc = ldc; a.rows_per_cta = pl.rows_per_cta;
            cagp::cross_cov_fwd_kernel<T, D, FAM><<<dim3((unsigned)pl.gx, (unsigned)pl.gy), cagp::kCrossThreads,
                                                     cross_smem<T, D>(), stream>>>(a);
            return check_launch("cross_cov_fwd_kernel");
        });
    });
}

template <typename T>
size_t cross_bwd_ws_bytes(int64_t m, int64_t n, int d, bool ard) {
    const CrossPlan pl = plan_cross(m, n);
    const int dp = cagp_padded_dim(d);
    return ((size_t)pl.gy * dp * n + (size_t)pl.gx * pl.gy * (ard ? dp : 1)) * sizeof(T) + 256;
}

template <typename T>
int cross_bwd_impl(const cagp_kernel_desc* kd, const void* xs1, int64_t m, int64_t ld1, const void* xs2, int64_t n,
                   int64_t ld2, int d, const void* Cbar, int64_t ldc, void* x2s_bar, void* ls_bar, void* ws,
                   size_t ws_bytes, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const bool ard = kd->n_lengthscale > 1;
    const CrossPlan pl = plan_cross(m, n);
    if (ws_bytes < cross_bwd_ws_bytes<T>(m, n, d, ard)) return fail(CAGP_ERR_WORKSPACE, "cross_cov bwd workspace too small");
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            auto launch = [&](auto ardc) {
                constexpr bool ARD = decltype(ardc)::value;
                constexpr int NLS = ARD ? D : 1;
                cagp::CrossArgs<T> a{};
                a.xs1 = (const T*)xs1; a.ld1 = ld1; a.m = m;
                a.xs2 = (const T*)xs2; a.ld2 = ld2; a.n = n;
                a.C = (T*)const_cast<void*>(Cbar); a.ldc = ldc; a.rows_per_cta = pl.rows_per_cta;
                a.x2_part = (T*)ws; a.ldb = n;
                a.ls_part = a.x2_part + (size_t)pl.gy * D * n;
                a.gc = (T)cagp::family_grad_const(kd->family);
                cagp::cross_cov_bwd_kernel<T, D, FAM, ARD><<<dim3((unsigned)pl.gx, (unsigned)pl.gy), cagp::kCrossThreads,
                                                              cross_smem<T, D>(), stream>>>(a);
                if (int rc = check_launch("cross_cov_bwd_kernel")) return rc;
                const int64_t tot = (int64_t)D * n;
                cagp::bwd_finalize_kernel<T><<<(unsigned)((tot + 255) / 256 + 1), 256, 0, stream>>>(
                    a.x2_part, (int)pl.gy, tot, (T*)x2s_bar, a.ls_part, pl.gx * pl.gy, NLS, kd->n_lengthscale,
                    (const T*)kd->lengthscale, a.gc, (T*)ls_bar);
                return check_launch("bwd_finalize_kernel");
            };
            return ard ? launch(std::true_type{}) : launch(std::false_type{});
        });
    });
}

template int cross_fwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t,
                                       int64_t, int, void*, int64_t, cudaStream_t);
template size_t cross_bwd_ws_bytes<CAGP_REAL>(int64_t, int64_t, int, bool);
template int cross_bwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, const void*, int64_t,
                                       int64_t, int, const void*, int64_t, void*, void*, void*, size_t, cudaStream_t);

}  // namespace cagp_host
