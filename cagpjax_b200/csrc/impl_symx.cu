// Launchers of the fast-path symmetric kernels (ks_symx.cuh): float64, RBF, isotropic lengthscale, SINGLE shape.
// Compiled for double only.
#include <cstdlib>
#include <cstring>

#include "ks_symx.cuh"
#include "sym_tiling.cuh"

namespace cagp_host {

namespace {

// (R, NT, JC, MINB, OPT) of the instantiated variants.  Variant 0 is the production choice (instantiated for D = 1..4);
// the others exist for D = 3 only, for A/B measurements and the in-situ ablations of tools/symx_sweep.py (forward 2-6,
// backward 2: WRONG RESULTS by construction), and are selected with CAGP_SYMX_FWD / CAGP_SYMX_BWD = index ("off"
// disables the fast path).  The sweeps behind the choice: profiles/r02_pair_kernels.md.
#define SYMX_FWD_VARIANTS(X) \
    X(0, 8, 128, 128, 3, 8)   \
    X(1, 4, 256, 128, 2, 0)   \
    X(2, 8, 128, 128, 3, 24)  \
    X(3, 8, 128, 128, 3, 40)  \
    X(4, 8, 128, 128, 3, 72)  \
    X(5, 8, 128, 128, 3, 136) \
    X(6, 8, 128, 128, 3, 264)

#define SYMX_BWD_VARIANTS(X) \
    X(0, 8, 128, 128, 2, 8)   \
    X(1, 4, 256, 128, 2, 2056) \
    X(2, 8, 128, 128, 2, 12)

int env_variant(const char* name, int dflt) {
    const char* v = std::getenv(name);
    if (!v || !*v) return dflt;
    if (!std::strcmp(v, "off")) return -1;
    return std::atoi(v);
}

struct Shape { int rows, nt; };
Shape fwd_shape(int variant) {
    switch (variant) {
#define X(ID, R, NT, JC, MINB, OPT) case ID: return Shape{R, NT};
        SYMX_FWD_VARIANTS(X)
#undef X
    }
    return Shape{0, 0};
}
Shape bwd_shape(int variant) {
    switch (variant) {
#define X(ID, R, NT, JC, MINB, OPT) case ID: return Shape{R, NT};
        SYMX_BWD_VARIANTS(X)
#undef X
    }
    return Shape{0, 0};
}

}  // namespace

int symx_fwd_variant() { return env_variant("CAGP_SYMX_FWD", 0); }
int symx_bwd_variant() { return env_variant("CAGP_SYMX_BWD", 0); }

// The fast path applies to: double, RBF, isotropic, SINGLE shape of the general kernel, D <= 4, a bounding box.
bool symx_applicable(const cagp_kernel_desc* kd, int64_t n, int k, int dp, bool bwd) {
    // the forward sees only the scaled inputs, so ARD lengthscales qualify; the backward's lengthscale cotangent uses
    // the isotropic identity sum_t (a_t - b_t)^2 = q and needs the per-dimension differences for ARD
    if (kd->dtype != CAGP_F64 || kd->family != CAGP_RBF || (bwd && kd->n_lengthscale != 1) || !kd->bbox || dp > 4) return false;
    if ((bwd ? symx_bwd_variant() : symx_fwd_variant()) < 0) return false;
    const Shape s = bwd ? bwd_shape(symx_bwd_variant()) : fwd_shape(symx_fwd_variant());
    if (s.rows == 0) return false;
    return !sym_shape(n, k, dp, false, bwd).spread;
}

SymShape symx_shape(bool bwd) {
    const Shape s = bwd ? bwd_shape(symx_bwd_variant()) : fwd_shape(symx_fwd_variant());
    return SymShape{s.rows, 0, s.nt};
}

template <int D, int R, int NT, int JC, int MINB, int OPT>
static int launch_fwd(const cagp::SymFwdArgs<double>& a, int64_t gx, int gy, cudaStream_t stream) {
    auto kern = cagp::ks_symx_fwd_kernel<D, R, NT, JC, MINB, OPT>;
    constexpr size_t smem = cagp::symx_smem_bytes<D, false, JC, NT>();
    if (smem > 40 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(CAGP_ERR_CUDA, "ks_symx_fwd_kernel: shared memory opt-in failed");
    kern<<<dim3((unsigned)gx, (unsigned)gy), NT, smem, stream>>>(a);
    return check_launch("ks_symx_fwd_kernel");
}

int symx_fwd_launch(int dp, cagp::SymFwdArgs<double> a, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream) {
    const int variant = symx_fwd_variant();
    a.til = make_sym_tiling(n, k, a_begin, a_end, symx_shape(false));
    a.defer_expandable = 0;
    const int64_t gx = sym_grid_x(a.til);
    if (gx < 1) return CAGP_OK;
    const int gy = sym_choose_groups(a.til, gx);
    return dispatch_dim(dp, [&](auto dc) -> int {
        constexpr int D = decltype(dc)::value;
        if constexpr (D <= 4) {
            switch (variant) {
#define X(ID, R, NT, JC, MINB, OPT) \
    case ID:                        \
        if constexpr (ID == 0 || D == 3) return launch_fwd<D, R, NT, JC, MINB, OPT>(a, gx, gy, stream); \
        break;
                SYMX_FWD_VARIANTS(X)
#undef X
            }
        }
        return fail(CAGP_ERR_UNSUPPORTED, "ks_symx_fwd: variant %d / dimension %d not instantiated", variant, D);
    });
}

template <int D, int R, int NT, int JC, int MINB, int OPT>
static int launch_bwd(const cagp::SymBwdArgs<double>& a, int64_t gx, int gy, cudaStream_t stream) {
    auto kern = cagp::ks_symx_bwd_kernel<D, R, NT, JC, MINB, OPT>;
    constexpr size_t smem = cagp::symx_smem_bytes<D, true, JC, NT>();
    if (smem > 40 * 1024 && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return fail(CAGP_ERR_CUDA, "ks_symx_bwd_kernel: shared memory opt-in failed");
    kern<<<dim3((unsigned)gx, (unsigned)gy), NT, smem, stream>>>(a);
    return check_launch("ks_symx_bwd_kernel");
}

// number of lengthscale partials the fast backward writes (one per CTA)
int64_t symx_bwd_nparts(int64_t n, int k, int a_begin, int a_end) {
    if (symx_shape(true).rows == 0) return 0;  // fast path disabled
    auto til = make_sym_tiling(n, k, a_begin, a_end, symx_shape(true));
    const int64_t gx = sym_grid_x(til);
    if (gx < 1) return 0;
    return gx * sym_choose_groups(til, gx);
}

int symx_bwd_launch(int dp, cagp::SymBwdArgs<double> a, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream) {
    const int variant = symx_bwd_variant();
    a.til = make_sym_tiling(n, k, a_begin, a_end, symx_shape(true));
    a.defer_expandable = 0;
    const int64_t gx = sym_grid_x(a.til);
    if (gx < 1) return CAGP_OK;
    const int gy = sym_choose_groups(a.til, gx);
    return dispatch_dim(dp, [&](auto dc) -> int {
        constexpr int D = decltype(dc)::value;
        if constexpr (D <= 4) {
            switch (variant) {
#define X(ID, R, NT, JC, MINB, OPT)                                                                  \
    case ID:                                                                                         \
        if constexpr (ID == 0 || D == 3) return launch_bwd<D, R, NT, JC, MINB, OPT>(a, gx, gy, stream); \
        break;
                SYMX_BWD_VARIANTS(X)
#undef X
            }
        }
        return fail(CAGP_ERR_UNSUPPORTED, "ks_symx_bwd: variant %d / dimension %d not instantiated", variant, D);
    });
}

}  // namespace cagp_host
