// Launcher of the symmetric forward kernel (one object per floating-point type).
#include "sym_tiling.cuh"

namespace cagp_host {

template <typename T>
int sym_fwd_impl(const cagp_kernel_desc* kd, const void* xs, int64_t n, int64_t ld, int d, const void* s, int k,
                 int a_begin, int a_end, void* const* panels, const int64_t* row_begin, int world, int me, int64_t ldm,
                 cudaStream_t stream) {
    const int dp = cagp_padded_dim(d);
    const SymShape sh = sym_shape(n, k, dp, false);
    return dispatch_dim(dp, [&](auto dc) {
        return dispatch_family(kd->family, [&](auto fc) {
            constexpr int D = decltype(dc)::value;
            constexpr int FAM = decltype(fc)::value;
            return dispatch_rows<D>(sh.rows, [&](auto rc) {
                constexpr int R = decltype(rc)::value;
                constexpr int JC = ColsPerStage<D>::value;
                cagp::SymFwdArgs<T> a;
                a.xs = (const T*)xs; a.ld = ld; a.s = (const T*)s; a.ldm = ldm;
                a.world = world; a.me = me;
                for (int r = 0; r < cagp::kMaxPeers; ++r) a.panel[r] = r < world ? (T*)panels[r] : nullptr;
                for (int r = 0; r <= cagp::kMaxPeers; ++r) a.row_begin[r] = r <= world ? row_begin[r] : n;
                void* Mt = panels[me];
                const int64_t own_rows = row_begin[me + 1] - row_begin[me];
                a.til = make_sym_tiling(n, k, a_begin, a_end, sh);
                a.bbox = (const T*)kd->bbox;
                a.defer_expandable = 0;
                const int64_t gx = sym_grid_x(a.til);
                if (gx < 1) return (int)CAGP_OK;
                const int gy = sym_choose_groups(a.til, gx);
                // fast path (ks_symx.cuh): both kernels are launched, the bounding box decides on the device which
                // of the two does the work (the other returns at once) - no host synchronisation
                bool fast = false;
                int tiles_main = a.til.tiles_main, tiles_last = a.til.tiles_last;
                if constexpr (std::is_same<T, double>::value && FAM == cagp::RBF && D <= 4) {
                    fast = !sh.spread && symx_applicable(kd, n, k, dp, false);
                    if (fast) {
                        const cagp::SymTiling tf = make_sym_tiling(n, k, a_begin, a_end, symx_shape(false));
                        tiles_main = tf.tiles_main > tiles_main ? tf.tiles_main : tiles_main;
                        tiles_last = tf.tiles_last > tiles_last ? tf.tiles_last : tiles_last;
                        a.defer_expandable = 1;
                    }
                }
                // column sums of a block that is split over several CTAs are accumulated atomically
                // (single device only; in peer mode the caller zeroes the panels before the cross-rank barrier)
                if (world == 1) {
                    if (tiles_main > 1) {
                        if (cudaMemsetAsync(Mt, 0, (size_t)k * ldm * sizeof(T), stream) != cudaSuccess) return fail(CAGP_ERR_CUDA, "memset Mt failed");
                    } else if (tiles_last > 1) {
                        if (cudaMemsetAsync((T*)Mt + (size_t)(k - 1) * ldm, 0, (size_t)own_rows * sizeof(T), stream) != cudaSuccess) return fail(CAGP_ERR_CUDA, "memset Mt failed");
                    }
                }
                auto launch = [&](auto sp) {
                    constexpr bool SPREAD = decltype(sp)::value;
                    constexpr int NC = SPREAD ? R : 1;
                    const size_t smem = (cagp::SymCanExpand<T, FAM, SPREAD>::kTabElems + (size_t)JC * (D + 2 + NC * 8)) * sizeof(T);
                    auto kern = cagp::ks_sym_fwd_kernel<T, D, FAM, R, JC, SPREAD>;
                    if (smem > 40 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                    kern<<<dim3((unsigned)gx, (unsigned)gy), a.til.nthreads, smem, stream>>>(a);
                    return check_launch("ks_sym_fwd_kernel");
                };
                if (sh.spread) {
                    if constexpr (R >= 2) return launch(std::true_type{});
                    else return fail(CAGP_ERR_UNSUPPORTED, "spread shape needs R >= 2");
                }
                if (int rc2 = launch(std::false_type{})) return rc2;
                if constexpr (std::is_same<T, double>::value && FAM == cagp::RBF && D <= 4) {
                    if (fast) return symx_fwd_launch_any<T>(dp, a, n, k, a_begin, a_end, stream);
                }
                return (int)CAGP_OK;
            });
        });
    });
}

template int sym_fwd_impl<CAGP_REAL>(const cagp_kernel_desc*, const void*, int64_t, int64_t, int, const void*, int, int,
                                     int, void* const*, const int64_t*, int, int, int64_t, cudaStream_t);

}  // namespace cagp_host
