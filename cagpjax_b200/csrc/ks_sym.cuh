// Symmetric (Gram, X1 == X2) block-sparse kernels: every K'(x_i, x_j) is evaluated ONCE per
// unordered pair of action blocks {a, b} and used for both M'[i, b] and M'[j, a].  sm_100a.
//
// Work unit: tile (a, b) = rows of action block a x columns of action block b, with
// b = (a + off) mod k for off = 0 .. floor(k/2) (off = 0 is the diagonal tile, computed as a full
// square with row sums only; for even k the offset k/2 is taken by a < k/2 only).  This covers every
// unordered block pair exactly once and gives every a the same number of tiles.
//
// Inside a tile a thread owns R rows (coordinates in registers).  A warp walks a 32-column group in
// 32 steps; at step t lane l handles column (l + t) mod 32, so the 32 lanes always read 32 DIFFERENT
// columns (conflict-free SoA smem reads) and after the 32 steps every (row, column) pair of the
// 32R x 32 sub-tile has been visited once.  Row sums stay in registers.  The partial COLUMN sum
// travels with the column: after each step a lane hands its running column accumulator to the lane
// that handles this column next (one 64-bit shuffle), so after 32 steps lane l holds the warp-wide
// sum for column l.  Warps then combine through shared memory.
//
//   forward : row out  Mt[b][i in a] = sum_{j in b} K'_ij s_j          (plain store)
//             col out  Mt[a][j in b] = sum_{i in a} K'_ij s_i          (store, or atomicAdd when block a
//                                                                      is split over several CTAs)
//   backward: s_bar_i += sum_j K'_ij Gt[a][j],  s_bar_j += sum_i K'_ij Gt[b][i]      (atomicAdd)
//             ls_bar  += sum_ij H_ij q_ij (Gt[b][i] s_j + Gt[a][j] s_i)              (per-CTA partials)
#pragma once
#include "ks_blocksparse.cuh"

namespace cagp {

template <typename T>
__device__ __forceinline__ T shfl_next(T v, int src_lane);
template <>
__device__ __forceinline__ double shfl_next<double>(double v, int src_lane) {
    return __shfl_sync(0xffffffffu, v, src_lane);
}
template <>
__device__ __forceinline__ float shfl_next<float>(float v, int src_lane) {
    return __shfl_sync(0xffffffffu, v, src_lane);
}

// A CTA has NW = 8 warps.  Blocks with at most 32*R*wpb rows are served by `wpb` warps each, so one CTA
// hosts G = NW / wpb consecutive row blocks ("sub-groups") that all face the SAME column block b: the
// column stage in shared memory is shared, every warp keeps R rows of ONE block in registers.  Sub-group g
// (block a0 + g) walks the offsets off0 - g, so that a + off = a0 + off0 = b for all of them.  Large
// blocks use wpb = 8 (G = 1) and, if they exceed 256*R rows, several row sub-tiles with atomics.  The
// last block (which absorbs n % k and may be up to twice as long) always gets its own CTAs.
struct SymTiling {
    BlockLayout lay;
    int wpb;            // warps per row block for the main blocks (1, 2, 4 or 8)
    int tiles_main;     // row sub-tiles per main block (> 1 only when wpb == 8)
    int tiles_last;     // row sub-tiles of the last block
    int a_begin, a_end; // action blocks whose tiles this launch computes
    int n_off;          // tile offsets 0 .. n_off-1  (n_off = k/2 + 1)
    int offs_per_cta;   // CTA-level offsets off0 per CTA (range [0, n_off + G - 1))

    __host__ __device__ int groups() const { return 8 / wpb; }
    __host__ __device__ int main_end() const { return a_end == lay.k ? a_end - 1 : a_end; }
    __host__ __device__ int64_t main_ctas() const {
        const int nb = main_end() - a_begin;
        if (nb <= 0) return 0;
        const int G = groups();
        return (int64_t)((nb + G - 1) / G) * tiles_main;
    }
    // tile partner of a at offset off; returns false if (a, off) is not a tile of the schedule
    __host__ __device__ bool partner(int a, int off, int& b) const {
        if (off < 0 || off >= n_off) return false;
        if (off == 0) { b = a; return true; }
        if (2 * off == lay.k && a >= off) return false;
        b = a + off;
        if (b >= lay.k) b -= lay.k;
        return true;
    }
};

// Per-CTA / per-warp placement shared by the forward and the backward kernel.
struct SymPlace {
    int a0;          // first row block of the CTA
    int G;           // sub-groups (row blocks) in this CTA
    int wpb;         // warps per sub-group
    int64_t sub_off; // row offset of this CTA's sub-tile inside its block(s)
    bool split;      // several CTAs share a row block -> column sums are accumulated atomically
    int a_limit;     // row blocks >= a_limit do not exist for this CTA
    int rows_stride; // row stride between a thread's R rows (= 32 * wpb)

    // MULTI == false: the tiling has wpb == 8 (one row block per CTA); constants fold away
    template <bool MULTI>
    __device__ static SymPlace make(const SymTiling& t, int bx, int R) {
        SymPlace p;
        const int64_t nmain = t.main_ctas();
        if (bx < nmain) {
            p.wpb = MULTI ? t.wpb : 8;
            p.G = MULTI ? t.groups() : 1;
            p.a0 = t.a_begin + (int)(bx / t.tiles_main) * p.G;
            p.sub_off = (int64_t)(bx % t.tiles_main) * (256 * R);
            p.split = t.tiles_main > 1;
            p.a_limit = t.main_end();
        } else {
            p.wpb = 8;
            p.G = 1;
            p.a0 = t.lay.k - 1;
            p.sub_off = (int64_t)(bx - nmain) * (256 * R);
            p.split = t.tiles_last > 1;
            p.a_limit = t.lay.k;
        }
        p.rows_stride = 32 * p.wpb;
        return p;
    }
};

template <typename T>
struct SymFwdArgs {
    const T* xs; int64_t ld;
    const T* s;
    T* Mt; int64_t ldm;
    SymTiling til;
};

template <typename T, int D, int FAM, int R, int NT, int JC, bool MULTI>
__global__ void __launch_bounds__(NT) ks_sym_fwd_kernel(const SymFwdArgs<T> p) {
    static_assert(NT == 256, "the symmetric kernels assume 8 warps per CTA");
    constexpr int NW = NT / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc = tab + ExpTab<T>::kSmemElems;  // [D][JC]
    T* sc = xc + D * JC;                  // [JC]
    T* colacc = sc + JC;                  // [NW][JC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, NT);
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymPlace pl = SymPlace::make<MULTI>(til, blockIdx.x, R);

    const int kG = MULTI ? pl.G : 1, kWpb = MULTI ? pl.wpb : 8;
    const int g = MULTI ? warp / kWpb : 0; // sub-group of this warp
    const int a = pl.a0 + g;               // its row block
    const bool a_ok = a < pl.a_limit;
    const int64_t i0 = a_ok ? lay.start(a) : 0, len_a = a_ok ? lay.len(a) : 0;
    T ar[R][D], srow[R];
    int64_t irow[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t o = pl.sub_off + (warp % kWpb) * 32 + lane + (int64_t)r * (32 * kWpb);
        valid[r] = o < len_a;
        irow[r] = i0 + (valid[r] ? o : (len_a > 0 ? len_a - 1 : 0));
#pragma unroll
        for (int t = 0; t < D; ++t) ar[r][t] = p.xs[t * p.ld + irow[r]];
        srow[r] = valid[r] ? p.s[irow[r]] : T(0);
    }
    const int n_off0 = til.n_off + kG - 1;
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(n_off0, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        // the column block is common to all sub-groups; a sub-group takes part if (a, off0 - g) is a tile
        bool any = false;
        for (int gg = 0; gg < kG; ++gg) {
            int bb;
            any = any || (pl.a0 + gg < pl.a_limit && til.partner(pl.a0 + gg, off0 - gg, bb));
        }
        if (!any) continue;
        int b = pl.a0 + off0;
        b = b % lay.k;
        int bw;
        const bool mine = a_ok && til.partner(a, off0 - g, bw);  // warp-uniform
        const bool diag = (off0 - g == 0);
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        T racc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) racc[r] = T(0);
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            for (int idx = tid; idx < clp; idx += NT) {
                const bool in = idx < cl;
                const int64_t j = j0 + c0 + (in ? idx : 0);
#pragma unroll
                for (int t = 0; t < D; ++t) xc[t * JC + idx] = in ? p.xs[t * p.ld + j] : T(0);
                sc[idx] = in ? p.s[j] : T(0);
            }
            __syncthreads();
            if (mine) {
                for (int sub = 0; sub < clp; sub += 32) {
                    T cacc = T(0);
#pragma unroll 2
                    for (int t = 0; t < 32; ++t) {
                        const int cidx = sub + ((lane + t) & 31);
                        T xb[D];
#pragma unroll
                        for (int tt = 0; tt < D; ++tt) xb[tt] = xc[tt * JC + cidx];
                        const T w = sc[cidx];
                        T cp = cacc;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const T q = sqdist<T, D>(ar[r], xb);
                            const T kv = kernel_value<T, FAM>(q, tab_s);
                            racc[r] = fma(kv, w, racc[r]);
                            cp = fma(kv, srow[r], cp);
                        }
                        cacc = shfl_next<T>(cp, next_lane);
                    }
                    colacc[warp * JC + sub + lane] = cacc;
                }
            }
            __syncthreads();
            // column sums: combine the warps of every sub-group that owns a symmetric tile
            for (int gg = 0; gg < kG; ++gg) {
                int bb;
                const int ag = pl.a0 + gg;
                if (!(ag < pl.a_limit && til.partner(ag, off0 - gg, bb)) || off0 - gg == 0) continue;
                for (int idx = tid; idx < cl; idx += NT) {
                    T v = T(0);
                    for (int w = 0; w < kWpb; ++w) v += colacc[(gg * kWpb + w) * JC + idx];
                    T* dst = p.Mt + (int64_t)ag * p.ldm + j0 + c0 + idx;
                    if (pl.split) atomicAdd(dst, v);
                    else *dst = v;
                }
            }
        }
        if (mine) {
#pragma unroll
            for (int r = 0; r < R; ++r)
                if (valid[r]) p.Mt[(int64_t)b * p.ldm + irow[r]] = racc[r];
        }
        (void)diag;
    }
}

template <typename T>
struct SymBwdArgs {
    const T* xs; int64_t ld;
    const T* s;
    const T* Gt; int64_t ldg;
    T* sbar;     // (n), zero-filled by the caller; accumulated with atomics
    T* ls_part;  // [gridDim.y * gridDim.x][NLS]
    SymTiling til;
};

template <typename T, int D, int FAM, bool ARD, int R, int NT, int JC, bool MULTI>
__global__ void __launch_bounds__(NT) ks_sym_bwd_kernel(const SymBwdArgs<T> p) {
    static_assert(NT == 256, "the symmetric kernels assume 8 warps per CTA");
    constexpr int NW = NT / 32;
    constexpr int NLS = ARD ? D : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc = tab + ExpTab<T>::kSmemElems;  // [D][JC]
    T* sc = xc + D * JC;                  // [JC]
    T* gc = sc + JC;                      // [G][JC]  G[j, a_g] for the columns, one row per sub-group (G <= 8)
    T* colacc = gc + NW * JC;             // [NW][JC]
    __shared__ T red[NW][NLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, NT);
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymPlace pl = SymPlace::make<MULTI>(til, blockIdx.x, R);

    const int kG = MULTI ? pl.G : 1, kWpb = MULTI ? pl.wpb : 8;
    const int g = MULTI ? warp / kWpb : 0;
    const int a = pl.a0 + g;
    const bool a_ok = a < pl.a_limit;
    const int64_t i0 = a_ok ? lay.start(a) : 0, len_a = a_ok ? lay.len(a) : 0;
    T ar[R][D], srow[R];
    int64_t irow[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t o = pl.sub_off + (warp % kWpb) * 32 + lane + (int64_t)r * (32 * kWpb);
        valid[r] = o < len_a;
        irow[r] = i0 + (valid[r] ? o : (len_a > 0 ? len_a - 1 : 0));
#pragma unroll
        for (int t = 0; t < D; ++t) ar[r][t] = p.xs[t * p.ld + irow[r]];
        srow[r] = valid[r] ? p.s[irow[r]] : T(0);
    }
    T sacc[R], l2[R][NLS], ltot[NLS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sacc[r] = T(0);
#pragma unroll
        for (int t = 0; t < NLS; ++t) l2[r][t] = T(0);
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) ltot[t] = T(0);
    const int n_off0 = til.n_off + kG - 1;
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(n_off0, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        bool any = false;
        for (int gg = 0; gg < kG; ++gg) {
            int bb;
            any = any || (pl.a0 + gg < pl.a_limit && til.partner(pl.a0 + gg, off0 - gg, bb));
        }
        if (!any) continue;
        const int b = (pl.a0 + off0) % lay.k;
        int bw;
        const bool mine = a_ok && til.partner(a, off0 - g, bw);
        const bool diag = (off0 - g == 0);
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        T grow[R], l1[R][NLS];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            grow[r] = (mine && valid[r]) ? p.Gt[(int64_t)b * p.ldg + irow[r]] : T(0);  // G[i, b]
#pragma unroll
            for (int t = 0; t < NLS; ++t) l1[r][t] = T(0);
        }
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            for (int idx = tid; idx < clp; idx += NT) {
                const bool in = idx < cl;
                const int64_t j = j0 + c0 + (in ? idx : 0);
#pragma unroll
                for (int t = 0; t < D; ++t) xc[t * JC + idx] = in ? p.xs[t * p.ld + j] : T(0);
                sc[idx] = in ? p.s[j] : T(0);
            }
            for (int gg = 0; gg < kG; ++gg) {
                const int ag = pl.a0 + gg;
                const T* __restrict__ ga = p.Gt + (int64_t)(ag < pl.a_limit ? ag : 0) * p.ldg;  // G[j, a_g]
                for (int idx = tid; idx < clp; idx += NT)
                    gc[gg * JC + idx] = (idx < cl && ag < pl.a_limit) ? ga[j0 + c0 + idx] : T(0);
            }
            __syncthreads();
            if (mine) {
                const T* __restrict__ gcg = gc + g * JC;
                for (int sub = 0; sub < clp; sub += 32) {
                    T cacc = T(0);
#pragma unroll 2
                    for (int t = 0; t < 32; ++t) {
                        const int cidx = sub + ((lane + t) & 31);
                        T xb[D];
#pragma unroll
                        for (int tt = 0; tt < D; ++tt) xb[tt] = xc[tt * JC + cidx];
                        const T scol = sc[cidx];
                        const T gcol_row = gcg[cidx];
                        const T gcol = diag ? T(0) : gcol_row;  // diagonal tile: ordered pairs, row part only
                        T cp = cacc;
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            T q = T(0);
                            T qt[D];
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) {
                                const T df = ar[r][tt] - xb[tt];
                                if (ARD) {
                                    qt[tt] = df * df;
                                    q += qt[tt];
                                } else {
                                    q = fma(df, df, q);
                                }
                            }
                            T kv, h;
                            kernel_value_grad<T, FAM>(q, tab_s, kv, h);
                            sacc[r] = fma(kv, gcol_row, sacc[r]);
                            cp = fma(kv, grow[r], cp);
                            if (ARD) {
#pragma unroll
                                for (int tt = 0; tt < D; ++tt) {
                                    const T hq = h * qt[tt];
                                    l1[r][tt] = fma(hq, scol, l1[r][tt]);
                                    l2[r][tt] = fma(hq, gcol, l2[r][tt]);
                                }
                            } else {
                                const T hq = h * q;
                                l1[r][0] = fma(hq, scol, l1[r][0]);
                                l2[r][0] = fma(hq, gcol, l2[r][0]);
                            }
                        }
                        cacc = shfl_next<T>(cp, next_lane);
                    }
                    colacc[warp * JC + sub + lane] = cacc;
                }
            }
            __syncthreads();
            for (int gg = 0; gg < kG; ++gg) {
                int bb;
                const int ag = pl.a0 + gg;
                if (!(ag < pl.a_limit && til.partner(ag, off0 - gg, bb)) || off0 - gg == 0) continue;
                for (int idx = tid; idx < cl; idx += NT) {
                    T v = T(0);
                    for (int w = 0; w < kWpb; ++w) v += colacc[(gg * kWpb + w) * JC + idx];
                    atomicAdd(p.sbar + j0 + c0 + idx, v);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int t = 0; t < NLS; ++t) ltot[t] = fma(grow[r], l1[r][t], ltot[t]);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (valid[r]) atomicAdd(p.sbar + irow[r], sacc[r]);
#pragma unroll
        for (int t = 0; t < NLS; ++t) ltot[t] = fma(srow[r], l2[r][t], ltot[t]);
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) {
        T v = ltot[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (tid < NLS) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < NW; ++w) v += red[w][tid];
        p.ls_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NLS + tid] = v;
    }
}

}  // namespace cagp
