// Launcher of the symmetric backward kernel (one object per floating-point type).
#include "aux_kernels.cuh"
#include "sym_tiling.cuh"

namespace cagp_host {

template <typename T>
size_t sym_bwd_ws_bytes(int64_t n, int d, int k, bool ard, int a_begin, int a_end) {
    const int dp = cagp_padded_dim(d);
    auto til = make_sym_tiling(n, k, a_begin, a_end, sym_shape(n, k, dp, ard, true));
    const int64_t gx = sym_grid_x(til);
    const int gy = sym_choose_groups(til, gx > 0 ? gx : 1);
    size_t bytes = (size_t)(gx > 0 ? gx : 1) * gy * (ard ? dp : 1) * sizeof(T) + 256;
    // the fast path (ks_symx.cuh) appends its own per-CTA partials
    if (std::is_same<T, double>::value && !ard && dp <= 4) bytes += (size_t)symx_bwd_nparts(n, k, a_begin, a_end) * sizeof(T);
    return bytes;
}

template <typename T>
int sym_bwd_impl(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int d, const void* s, int k,
                 int a_begin, int a_end, const void* Gt, int64_t ldg, void* s_bar, void* ls_bar, void* ws,
                 size_t ws_bytes, cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const bool ard = kd->n_lengthscale > 1;
    const SymShape sh = sym_shape(n, k, dp, ard, true);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            return dispatch_rows<D>(sh.rows, [&](auto rc) {
                constexpr int R = decltype(rc)::value;
                constexpr int JC = ColsPerStage<D>::value;
                auto launch = [&](auto ardc, auto sp) {
                    constexpr bool ARD = decltype(ardc)::value;
                    constexpr bool SPREAD = decltype(sp)::value;
                    constexpr int NLS = ARD ? D : 1;
                    constexpr int NC = SPREAD ? R : 1;
                    cagp::SymBwdArgs<T> a;
                    a.xs = (const T*)xs; a.ld = ld; a.s = (const T*)s; a.Gt = (const T*)Gt; a.ldg = ldg;
                    a.sbar = (T*)s_bar; a.ls_part = (T*)ws;
                    a.til = make_sym_tiling(n, k, a_begin, a_end, sh);
                    a.bbox = (const T*)kd->bbox;
                    a.defer_expandable = 0;
                    const int64_t gx = sym_grid_x(a.til);
                    if (cudaMemsetAsync(s_bar, 0, (size_t)n * sizeof(T), stream) != cudaSuccess) return fail(CAGP_ERR_CUDA, "memset s_bar failed");
                    int64_t nparts = 0;
                    // fast path (ks_symx.cuh): both kernels are launched and the bounding box decides on the device
                    // which one does the work; the idle one leaves its (pre-zeroed) partials untouched
                    int64_t nparts_fast = 0;
                    if constexpr (std::is_same<T, double>::value && FAM == cagp::RBF && D <= 4 && !ARD && !SPREAD) {
                        if (symx_applicable(kd, n, k, dp, true)) {
                            nparts_fast = symx_bwd_nparts(n, k, a_begin, a_end);
                            a.defer_expandable = nparts_fast > 0 ? 1 : 0;
                        }
                    }
                    if (gx >= 1) {
                        const int gy = sym_choose_groups(a.til, gx);
                        nparts = gx * gy;
                        if (ws_bytes < (size_t)(nparts * NLS + nparts_fast) * sizeof(T)) return fail(CAGP_ERR_WORKSPACE, "ks sym bwd workspace too small");
                        if (nparts_fast > 0 && cudaMemsetAsync(ws, 0, (size_t)(nparts * NLS + nparts_fast) * sizeof(T), stream) != cudaSuccess)
                            return fail(CAGP_ERR_CUDA, "memset ws failed");
                        const size_t smem = (cagp::SymCanExpand<T, FAM, SPREAD>::kTabElems + (size_t)JC * (D + 2 + NC * 9)) * sizeof(T);
                        auto kern = cagp::ks_sym_bwd_kernel<T, D, FAM, ARD, R, JC, SPREAD>;
                        if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                        kern<<<dim3((unsigned)gx, (unsigned)gy), a.til.nthreads, smem, stream>>>(a);
                        if (int rc2 = check_launch("ks_sym_bwd_kernel")) return rc2;
                    }
                    if constexpr (std::is_same<T, double>::value && FAM == cagp::RBF && D <= 4 && !ARD && !SPREAD) {
                        if (nparts_fast > 0) {
                            if (int rc2 = symx_bwd_launch_any<T>(dp, a, a.ls_part + nparts /* NLS == 1 */, n, k, a_begin, a_end, stream)) return rc2;
                            nparts += nparts_fast;
                        }
                    }
                    cagp::bwd_finalize_kernel<T><<<1, 256, 0, stream>>>(nullptr, 0, 0, nullptr, a.ls_part, nparts, NLS,
                                                                         kd->n_lengthscale, (const T*)kd->lengthscale,
                                                                         (T)cagp::family_grad_const(kd->family), (T*)ls_bar);
                    return check_launch("bwd_finalize_kernel");
                };
                // per-dimension accumulators: only R <= MaxRows/2 is instantiated; SPREAD needs R >= 2
                constexpr int RARD = MaxRows<D>::value > 1 ? MaxRows<D>::value / 2 : 1;
                if (ard) {
                    if constexpr (R <= RARD) {
                        if (sh.spread) {
                            if constexpr (R >= 2) return launch(std::true_type{}, std::true_type{});
                            else return fail(CAGP_ERR_UNSUPPORTED, "spread shape needs R >= 2");
                        }
                        return launch(std::true_type{}, std::false_type{});
                    } else {
                        return fail(CAGP_ERR_UNSUPPORTED, "ARD backward with %d rows per thread not instantiated", R);
                    }
                }
                if (sh.spread) {
                    if constexpr (R >= 2) return launch(std::false_type{}, std::true_type{});
                    else return fail(CAGP_ERR_UNSUPPORTED, "spread shape needs R >= 2");
                }
                return launch(std::false_type{}, std::false_type{});
            });
        });
    });
}

template size_t sym_bwd_ws_bytes<CAGP_REAL>(int64_t, int, int, bool, int, int);
template int sym_bwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, int, const void*, int, int,
                                     int, const void*, int64_t, void*, void*, void*, size_t, cudaStream_t);

}  // namespace cagp_host
