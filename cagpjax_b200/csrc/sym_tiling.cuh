// Tiling of the symmetric (Gram) kernels, shared by the forward and backward launchers.
#pragma once
#include "host_common.cuh"
#include "ks_sym.cuh"

namespace cagp_host {

constexpr int kSymThreads = 256;

// CTA shape for the symmetric kernels (see csrc/ks_sym.cuh): SPREAD for blocks of at most 256 rows when at
// least two rows per thread fit in registers, SINGLE otherwise.
struct SymShape {
    int rows, spread, nthreads;
};
inline SymShape sym_shape(int64_t n, int k, int dp, bool ard_bwd, bool bwd = false) {
    int rmax = max_rows(dp);
    if (ard_bwd && rmax > 1) rmax /= 2;  // per-dimension accumulators double the register need
    const int64_t bs = n / k;
    if (bs >= 1 && bs <= kSymThreads && rmax >= 2 && k >= 2) {
        // the spread backward kernel with 4 row blocks per thread needs 174 registers (1 CTA/SM): use 2
        const int r = (bwd && rmax > 2) ? 2 : rmax;
        return SymShape{r, 1, (int)((bs + 31) / 32 * 32)};
    }
    return SymShape{choose_rows(bs, kSymThreads, rmax), 0, kSymThreads};
}

inline cagp::SymTiling make_sym_tiling(int64_t n, int k, int a_begin, int a_end, SymShape sh) {
    cagp::SymTiling t;
    t.lay = cagp::BlockLayout{n, k, n / k};
    t.spread = sh.spread;
    t.rows = sh.rows;
    t.nthreads = sh.nthreads;
    const int64_t cap = sh.spread ? (int64_t)sh.nthreads : (int64_t)sh.nthreads * sh.rows;
    t.tiles_main = sh.spread ? 1 : (int)((t.lay.bs + cap - 1) / cap);
    if (t.tiles_main < 1) t.tiles_main = 1;
    t.tiles_last = (int)((t.lay.len(k - 1) + cap - 1) / cap);
    if (t.tiles_last < 1) t.tiles_last = 1;
    t.a_begin = a_begin;
    t.a_end = a_end;
    t.n_off = k / 2 + 1;
    t.offs_per_cta = 1;
    return t;
}

inline int64_t sym_grid_x(const cagp::SymTiling& t) {
    return t.main_ctas() + (t.a_end == t.lay.k ? t.tiles_last : 0);
}

inline int sym_choose_groups(cagp::SymTiling& t, int64_t gx) {
    const int n_off0 = t.n_off0();
    const int64_t target = (int64_t)num_sms() * 2 * 128;
    int64_t gy = (target + gx - 1) / gx;
    if (gy > n_off0) gy = n_off0;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    t.offs_per_cta = (int)((n_off0 + gy - 1) / gy);
    return (n_off0 + t.offs_per_cta - 1) / t.offs_per_cta;
}

// ---- fast path (ks_symx.cuh, impl_symx.cu): float64 RBF, isotropic, SINGLE shape, D <= 4, bounding box given ----
bool symx_applicable(const cagp_kernel_desc* kd, int64_t n, int k, int dp, bool bwd);
SymShape symx_shape(bool bwd);
int symx_fwd_launch(int dp, cagp::SymFwdArgs<double> a, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream);
int64_t symx_bwd_nparts(int64_t n, int k, int a_begin, int a_end);
int symx_bwd_launch(int dp, cagp::SymBwdArgs<double> a, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream);
// type-generic shims (the fast path exists for double only)
template <typename T>
int symx_fwd_launch_any(int dp, const cagp::SymFwdArgs<T>& a, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream) {
    if constexpr (std::is_same<T, double>::value) return symx_fwd_launch(dp, a, n, k, a_begin, a_end, stream);
    else return CAGP_OK;
}
template <typename T>
int symx_bwd_launch_any(int dp, const cagp::SymBwdArgs<T>& a, T* ls_part, int64_t n, int k, int a_begin, int a_end, cudaStream_t stream) {
    if constexpr (std::is_same<T, double>::value) {
        cagp::SymBwdArgs<double> af = a;
        af.ls_part = ls_part;
        return symx_bwd_launch(dp, af, n, k, a_begin, a_end, stream);
    } else {
        return CAGP_OK;
    }
}

}  // namespace cagp_host
