// Tiling of the symmetric (Gram) kernels, shared by the forward and backward launchers.
#pragma once
#include "host_common.cuh"
#include "ks_sym.cuh"

namespace cagp_host {

constexpr int kSymThreads = 256;

inline int sym_rows(int64_t n, int k, int dp, bool ard_bwd) {
    int rmax = max_rows(dp);
    if (ard_bwd && rmax > 1) rmax /= 2;  // per-dimension accumulators double the register need
    return choose_rows(n / k, kSymThreads, rmax);
}

inline cagp::SymTiling make_sym_tiling(int64_t n, int k, int a_begin, int a_end, int rows) {
    cagp::SymTiling t;
    t.lay = cagp::BlockLayout{n, k, n / k};
    const int64_t cap = (int64_t)kSymThreads * rows;
    t.tiles_main = (int)((t.lay.bs + cap - 1) / cap);
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
    const int k = t.lay.k;
    const int n_main = (t.a_end == k ? t.a_end - 1 : t.a_end) - t.a_begin;
    return (int64_t)(n_main > 0 ? n_main : 0) * t.tiles_main + (t.a_end == k ? t.tiles_last : 0);
}

inline int sym_choose_groups(cagp::SymTiling& t, int64_t gx) {
    const int64_t target = (int64_t)num_sms() * 2 * 24;
    int64_t gy = (target + gx - 1) / gx;
    if (gy > t.n_off) gy = t.n_off;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    t.offs_per_cta = (int)((t.n_off + gy - 1) / gy);
    return (t.n_off + t.offs_per_cta - 1) / t.offs_per_cta;
}

}  // namespace cagp_host
