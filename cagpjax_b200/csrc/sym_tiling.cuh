// Tiling of the symmetric (Gram) kernels, shared by the forward and backward launchers.
#pragma once
#include "host_common.cuh"
#include "ks_sym.cuh"

namespace cagp_host {

constexpr int kSymThreads = 256;

// rows per thread R and warps per row block wpb for the main blocks: the smallest capacity
// 32 * wpb * R that covers a block, preferring large R (column loads are amortised over R rows)
struct SymShape {
    int rows, wpb;
};
inline SymShape sym_shape(int64_t n, int k, int dp, bool ard_bwd) {
    int rmax = max_rows(dp);
    if (ard_bwd && rmax > 1) rmax /= 2;  // per-dimension accumulators double the register need
    const int64_t bs = n / k;
    SymShape best{rmax, 8};
    int64_t best_cap = -1;
    for (int r = rmax; r >= 1; r /= 2) {
        for (int w = 1; w <= 8; w *= 2) {
            const int64_t cap = (int64_t)32 * w * r;
            if (cap >= bs && (best_cap < 0 || cap < best_cap)) {
                best_cap = cap;
                best = SymShape{r, w};
            }
        }
    }
    return best;  // blocks larger than 256 * rmax: {rmax, 8} with several row sub-tiles
}

inline cagp::SymTiling make_sym_tiling(int64_t n, int k, int a_begin, int a_end, SymShape sh) {
    cagp::SymTiling t;
    t.lay = cagp::BlockLayout{n, k, n / k};
    t.wpb = sh.wpb;
    const int64_t cap = (int64_t)kSymThreads * sh.rows;
    t.tiles_main = sh.wpb == 8 ? (int)((t.lay.bs + cap - 1) / cap) : 1;
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
    const int n_off0 = t.n_off + t.groups() - 1;
    const int64_t target = (int64_t)num_sms() * 2 * 24;
    int64_t gy = (target + gx - 1) / gx;
    if (gy > n_off0) gy = n_off0;
    if (gy < 1) gy = 1;
    if (gy > 65535) gy = 65535;
    t.offs_per_cta = (int)((n_off0 + gy - 1) / gy);
    return (n_off0 + t.offs_per_cta - 1) / t.offs_per_cta;
}

}  // namespace cagp_host
