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
#include <type_traits>

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

// Two CTA shapes.
//  SINGLE: 256 threads, each with R rows of ONE row block a (rows o = sub*256R + r*256 + tid).  Blocks longer
//          than 256 R rows take several row sub-tiles (column sums then go through atomics).
//  SPREAD: for blocks of at most 256 rows.  blockDim = 32 * ceil(bs / 32) threads, thread t holds row t of R
//          DIFFERENT consecutive row blocks a0 .. a0+R-1, which all face the same column block b: row block
//          a0 + r walks the offsets off0 - r, so that a_r + off_r = a0 + off0 = b.  The column stage in shared
//          memory and the column loads of the inner loop are shared by the R rows, lanes are padded to a
//          multiple of 32 only (87 % occupancy at 195-row blocks instead of 76 % with power-of-two warp groups).
// The last block (which absorbs n % k and may be up to twice as long) always gets its own CTAs.
struct SymTiling {
    BlockLayout lay;
    int spread;         // 1: SPREAD shape, 0: SINGLE
    int rows;           // R of the launched kernel
    int nthreads;       // threads per CTA
    int tiles_main;     // SINGLE: row sub-tiles per main block
    int tiles_last;     // row sub-tiles of the last block
    int a_begin, a_end; // action blocks whose tiles this launch computes
    int n_off;          // tile offsets 0 .. n_off-1  (n_off = k/2 + 1)
    int offs_per_cta;   // CTA-level offsets off0 per CTA

    __host__ __device__ int main_end() const { return a_end == lay.k ? a_end - 1 : a_end; }
    __host__ __device__ int64_t main_ctas() const {
        const int nb = main_end() - a_begin;
        if (nb <= 0) return 0;
        return spread ? (int64_t)((nb + rows - 1) / rows) : (int64_t)nb * tiles_main;
    }
    __host__ __device__ int n_off0() const { return spread ? n_off + rows - 1 : n_off; }
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

// The expanded-distance RBF form (kernel_math.cuh: launch_is_expandable, exp2neg16) is compiled into the fp64 RBF
// kernels of the SINGLE shape only: its replicated exp2 table needs 32 KB of shared memory instead of 16 KB, which
// would cost the SPREAD shape (large column accumulators) a resident CTA.
template <typename T, int FAM, bool SPREAD>
struct SymCanExpand {
    static constexpr bool value = sizeof(T) == 8 && FAM == RBF && !SPREAD;
    static constexpr int kTabElems = value ? kExpTab16Elems : ExpTab<T>::kSmemElems;
};

constexpr int kMaxPeers = 16;
// column tails of at most this many columns use the broadcast formulation instead of a 32-step ring
constexpr int kSymTailBroadcast = 20;

// Output placement.  Single device: one panel Mt (k, ldm) holding all n columns (world = 1, row_begin =
// {0, n}).  Peer mode (one process per GPU, NVLink peer memory): rank h owns the rows [row_begin[h],
// row_begin[h+1]) - aligned to action blocks - and a panel (k, ldm) for them; a tile's row sums land in
// the computing rank's own panel, its column sums are stored DIRECTLY into the panel of the rank that owns
// the column block (remote st.global / red.global over NVLink), so the forward needs no collective.
template <typename T>
struct SymFwdArgs {
    const T* xs; int64_t ld;
    const T* s;
    T* panel[kMaxPeers]; int64_t ldm;
    int64_t row_begin[kMaxPeers + 1];
    int world, me;
    const T* bbox;  // bounding box of the scaled inputs (2 * D values) or null
    SymTiling til;
    int defer_expandable;  // 1: return at once when the expanded form applies (ks_symx.cuh computes the launch)
};

template <typename T>
struct SymBwdArgs {
    const T* xs; int64_t ld;
    const T* s;
    const T* Gt; int64_t ldg;
    T* sbar;     // (n), zero-filled by the caller; accumulated with atomics
    T* ls_part;  // [gridDim.y * gridDim.x][NLS]
    const T* bbox;  // bounding box of the scaled inputs (2 * D values) or null
    SymTiling til;
    int defer_expandable;  // see SymFwdArgs
};

// Row placement of a CTA.  SINGLE: all R rows in block a0 (a_of(r) == a0).  SPREAD: row r in block a0 + r.
template <int R, bool SPREAD>
struct SymRows {
    int a0;        // first (SPREAD) / only (SINGLE) row block
    int nvalid;    // SPREAD: number of existing row blocks among a0 .. a0+R-1 (SINGLE: R)
    bool split;    // several CTAs share a row block -> column sums go through atomics
    int64_t o[R];  // row offset inside the block

    __device__ int a_of(int r) const { return SPREAD ? a0 + r : a0; }
    __device__ bool block_ok(int r) const { return SPREAD ? r < nvalid : true; }

    __device__ static SymRows make(const SymTiling& t, int bx, int tid, int nthreads) {
        SymRows p;
        const int64_t nmain = t.main_ctas();
        if (bx < nmain) {
            if (SPREAD) {
                p.a0 = t.a_begin + bx * R;
                p.nvalid = min(R, t.main_end() - p.a0);
                p.split = false;
#pragma unroll
                for (int r = 0; r < R; ++r) p.o[r] = tid;
            } else {
                p.a0 = t.a_begin + (int)(bx / t.tiles_main);
                p.nvalid = R;
                p.split = t.tiles_main > 1;
                const int64_t sub = (int64_t)(bx % t.tiles_main) * (nthreads * R);
#pragma unroll
                for (int r = 0; r < R; ++r) p.o[r] = sub + (int64_t)r * nthreads + tid;
            }
        } else {  // the last block: one row block per CTA in both shapes
            p.a0 = t.lay.k - 1;
            p.split = t.tiles_last > 1;
            const int64_t s = bx - nmain;
            if (SPREAD) {
                p.nvalid = 1;
#pragma unroll
                for (int r = 0; r < R; ++r) p.o[r] = s * nthreads + tid;
            } else {
                p.nvalid = R;
#pragma unroll
                for (int r = 0; r < R; ++r) p.o[r] = s * (nthreads * R) + (int64_t)r * nthreads + tid;
            }
        }
        return p;
    }
};

template <typename T, int D, int FAM, int R, int JC, bool SPREAD>
__global__ void __launch_bounds__(256) ks_sym_fwd_kernel(const SymFwdArgs<T> p) {
    constexpr int NWMAX = 8;
    constexpr int NC = SPREAD ? R : 1;  // independent column accumulators (one per row block)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool CAN_EXPAND = SymCanExpand<T, FAM, SPREAD>::value;
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc = tab + SymCanExpand<T, FAM, SPREAD>::kTabElems;  // [D][JC]
    T* sc = xc + D * JC;                  // [JC]
    T* nbc = sc + JC;                     // [JC]  |x_j|^2 of the staged columns (expanded form only)
    T* colacc = nbc + JC;                 // [NC][NWMAX][JC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nw = nthreads >> 5;
    const bool safe = launch_is_safe<T, D, FAM>(p.bbox);  // uniform: no clamp needed in exp2neg
    // uniform: expanded distance form with the replicated exp2 table (kernel_math.cuh)
    const bool expand = CAN_EXPAND && launch_is_expandable<T, D, FAM>(p.bbox);
    // uniform: the fast-path kernel (ks_symx.cuh) does this launch (it also covers |x|^2 < 900 with a clamp)
    if (p.defer_expandable && launch_expand_level<T, D, FAM>(p.bbox) > 0) return;
    // uniform: expanded distance q = |a|^2 + (|b|^2 - 2 a.b) for the smooth Matern families (kernel_math.cuh)
    constexpr bool CAN_XDIST = sizeof(T) == 8 && (FAM == MATERN32 || FAM == MATERN52);
    const bool xdist = CAN_XDIST && safe && launch_is_xdist<T, D, FAM>(p.bbox);
    unsigned tab_s;
    if constexpr (CAN_EXPAND) {
        tab_s = expand ? exptab16_load(tab, tid, nthreads) : ExpTab<T>::load(tab, tid, nthreads);
    } else {
        tab_s = ExpTab<T>::load(tab, tid, nthreads);
    }
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymRows<R, SPREAD> pl = SymRows<R, SPREAD>::make(til, blockIdx.x, tid, nthreads);

    T ar[R][D], srow[R], erow[R], narow[R];
    int64_t irow[R];
    bool valid[R];
    __syncthreads();  // the exp2 table is used below (expanded form)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool bok = pl.block_ok(r);
        const int a = bok ? pl.a_of(r) : 0;
        const int64_t len_a = bok ? lay.len(a) : 0;
        valid[r] = pl.o[r] < len_a;
        irow[r] = lay.start(a) + (valid[r] ? pl.o[r] : (len_a > 0 ? len_a - 1 : 0));
        if (irow[r] >= lay.n) irow[r] = lay.n - 1;
#pragma unroll
        for (int t = 0; t < D; ++t) ar[r][t] = p.xs[t * p.ld + irow[r]];
        srow[r] = valid[r] ? p.s[irow[r]] : T(0);
        erow[r] = T(1);
        narow[r] = T(0);
        if constexpr (CAN_EXPAND) {
            if (expand) {  // row factor 2^(-|a|^2); the column partials take s_i 2^(-|a|^2)
                T na = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) na = fma(ar[r][t], ar[r][t], na);
                erow[r] = exp2neg16(na, tab_s);
                srow[r] *= erow[r];
            }
        }
        if constexpr (CAN_XDIST) {
            if (xdist) {
#pragma unroll
                for (int t = 0; t < D; ++t) narow[r] = fma(ar[r][t], ar[r][t], narow[r]);
            }
        }
    }
    const int n_off0 = til.n_off0();
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(n_off0, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        // CTA-uniform tile flags: row block of r takes part if (a_r, off_r) is a tile of the schedule
        bool mine[NC], any = false;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int bb;
            mine[c] = pl.block_ok(c) && til.partner(pl.a_of(c), SPREAD ? off0 - c : off0, bb);
            any = any || mine[c];
        }
        if (!any) continue;
        const int b = (pl.a0 + off0) % lay.k;
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        int owner = 0;  // rank that owns the rows of column block b
        while (owner + 1 < p.world && j0 >= p.row_begin[owner + 1]) ++owner;
        T* __restrict__ col_panel = p.panel[owner] - p.row_begin[owner];
        T* __restrict__ row_panel = p.panel[p.me] - p.row_begin[p.me];
        T racc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) racc[r] = T(0);
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            for (int idx = tid; idx < clp; idx += nthreads) {
                const bool in = idx < cl;
                const int64_t j = j0 + c0 + (in ? idx : 0);
                T nb = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T v = in ? p.xs[t * p.ld + j] : T(0);
                    nb = fma(v, v, nb);
                    xc[t * JC + idx] = (expand || xdist) ? T(-2) * v : v;
                }
                nbc[idx] = nb;
                sc[idx] = in ? p.s[j] : T(0);
            }
            __syncthreads();
            auto sweep = [&](auto mode_c) {
                // 0 general, 1 safe range, 2 safe + expanded RBF form, 3 safe + expanded distance (Matern-3/2, -5/2)
                constexpr int MODE = decltype(mode_c)::value;
                constexpr bool SAFE = MODE >= 1;
                // one (thread, column) step: R kernel values feed the row sums and the column partials
                auto pair_step = [&](int cidx, T (&cacc)[NC]) {
                    T xb[D];
#pragma unroll
                    for (int tt = 0; tt < D; ++tt) xb[tt] = xc[tt * JC + cidx];
                    const T w = sc[cidx];
                    if constexpr (MODE == 2) {
                        const T nb = nbc[cidx];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            T u = nb;  // |b|^2 - 2 a.b  (xb holds -2 b)
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) u = fma(ar[r][tt], xb[tt], u);
                            const T kv = exp2neg16<true>(u, tab_s);  // K' / 2^(-|a|^2)
                            racc[r] = fma(kv, w, racc[r]);
                            cacc[SPREAD ? r : 0] = fma(kv, srow[r], cacc[SPREAD ? r : 0]);
                        }
                    } else if constexpr (MODE == 3) {
                        const T nb = nbc[cidx];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            T u = nb;  // |b|^2 - 2 a.b  (xb holds -2 b)
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) u = fma(ar[r][tt], xb[tt], u);
                            const T kv = kernel_value<T, FAM, true>(narow[r] + u, tab_s);  // fast_sqrt clamps q at 0+
                            racc[r] = fma(kv, w, racc[r]);
                            cacc[SPREAD ? r : 0] = fma(kv, srow[r], cacc[SPREAD ? r : 0]);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const T q = sqdist<T, D>(ar[r], xb);
                            const T kv = kernel_value<T, FAM, SAFE>(q, tab_s);
                            racc[r] = fma(kv, w, racc[r]);
                            cacc[SPREAD ? r : 0] = fma(kv, srow[r], cacc[SPREAD ? r : 0]);
                        }
                    }
                };
                for (int sub = 0; sub < clp; sub += 32) {
                    const int rem = cl - sub;
                    if (rem > kSymTailBroadcast) {
                        // ring: lane l visits column (l + t) mod 32 at step t, the column partial travels along
                        T cacc[NC];
#pragma unroll
                        for (int c = 0; c < NC; ++c) cacc[c] = T(0);
#pragma unroll 2
                        for (int t = 0; t < 32; ++t) {
                            pair_step(sub + ((lane + t) & 31), cacc);
#pragma unroll
                            for (int c = 0; c < NC; ++c) cacc[c] = shfl_next<T>(cacc[c], next_lane);
                        }
#pragma unroll
                        for (int c = 0; c < NC; ++c) colacc[(c * NWMAX + warp) * JC + sub + lane] = cacc[c];
                    } else {
                        // short tail of a block: all lanes take the same column, column sums by warp reduction
                        // (a 32-step ring for a handful of columns would waste most of its slots)
                        for (int jj = 0; jj < rem; ++jj) {
                            T cp[NC];
#pragma unroll
                            for (int c = 0; c < NC; ++c) cp[c] = T(0);
                            pair_step(sub + jj, cp);
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                T v = cp[c];
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                if (lane == 0) colacc[(c * NWMAX + warp) * JC + sub + jj] = v;
                            }
                        }
                    }
                }
            };
            if constexpr (sizeof(T) == 8) {
                if constexpr (CAN_EXPAND) {
                    if (expand) sweep(std::integral_constant<int, 2>{});
                    else if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                } else if constexpr (CAN_XDIST) {
                    if (xdist) sweep(std::integral_constant<int, 3>{});
                    else if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                } else {
                    if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                }
            } else {
                sweep(std::integral_constant<int, 0>{});
            }
            __syncthreads();
            // column sums: Mt[a_c][j in b], for the row blocks that own a symmetric (off != 0) tile
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (!mine[c] || (SPREAD ? off0 - c : off0) == 0) continue;
                const int ac = pl.a_of(c);
                for (int idx = tid; idx < cl; idx += nthreads) {
                    T v = T(0);
                    for (int w = 0; w < nw; ++w) v += colacc[(c * NWMAX + w) * JC + idx];
                    T* dst = col_panel + (int64_t)ac * p.ldm + j0 + c0 + idx;
                    if (pl.split) atomicAdd(dst, v);
                    else *dst = v;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (mine[SPREAD ? r : 0] && valid[r]) row_panel[(int64_t)b * p.ldm + irow[r]] = racc[r] * erow[r];
    }
}

template <typename T, int D, int FAM, bool ARD, int R, int JC, bool SPREAD>
__global__ void __launch_bounds__(256) ks_sym_bwd_kernel(const SymBwdArgs<T> p) {
    constexpr int NWMAX = 8;
    constexpr int NC = SPREAD ? R : 1;
    constexpr int NLS = ARD ? D : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool CAN_EXPAND = SymCanExpand<T, FAM, SPREAD>::value && !ARD;  // isotropic: only q is needed, not q_t
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc = tab + SymCanExpand<T, FAM, SPREAD>::kTabElems;  // [D][JC]
    T* sc = xc + D * JC;                  // [JC]
    T* nbc = sc + JC;                     // [JC]  |x_j|^2 of the staged columns (expanded form only)
    T* gc = nbc + JC;                     // [NC][JC]   G[j, a_c] for the columns
    T* colacc = gc + NC * JC;             // [NC][NWMAX][JC]
    __shared__ T red[NWMAX][NLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nw = nthreads >> 5;
    const bool safe = launch_is_safe<T, D, FAM>(p.bbox);  // uniform: no clamp needed in exp2neg
    const bool expand = CAN_EXPAND && launch_is_expandable<T, D, FAM>(p.bbox);  // uniform, see the forward kernel
    if (p.defer_expandable && launch_expand_level<T, D, FAM>(p.bbox) > 0) return;  // uniform: ks_symx.cuh does this launch
    // uniform: expanded distance for the smooth Matern families, isotropic lengthscale (only q is needed, not q_t)
    constexpr bool CAN_XDIST = sizeof(T) == 8 && (FAM == MATERN32 || FAM == MATERN52) && !ARD;
    const bool xdist = CAN_XDIST && safe && launch_is_xdist<T, D, FAM>(p.bbox);
    unsigned tab_s;
    if constexpr (CAN_EXPAND) {
        tab_s = expand ? exptab16_load(tab, tid, nthreads) : ExpTab<T>::load(tab, tid, nthreads);
    } else {
        tab_s = ExpTab<T>::load(tab, tid, nthreads);
    }
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymRows<R, SPREAD> pl = SymRows<R, SPREAD>::make(til, blockIdx.x, tid, nthreads);

    T ar[R][D], srow[R], erow[R], narow[R];
    int64_t irow[R];
    bool valid[R];
    __syncthreads();  // the exp2 table is used below (expanded form)
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool bok = pl.block_ok(r);
        const int a = bok ? pl.a_of(r) : 0;
        const int64_t len_a = bok ? lay.len(a) : 0;
        valid[r] = pl.o[r] < len_a;
        irow[r] = lay.start(a) + (valid[r] ? pl.o[r] : (len_a > 0 ? len_a - 1 : 0));
        if (irow[r] >= lay.n) irow[r] = lay.n - 1;
#pragma unroll
        for (int t = 0; t < D; ++t) ar[r][t] = p.xs[t * p.ld + irow[r]];
        srow[r] = valid[r] ? p.s[irow[r]] : T(0);
        erow[r] = T(1);
        narow[r] = T(0);
        if constexpr (CAN_EXPAND) {
            if (expand) {  // K' = erow * 2^(-(nb - 2 a.b)); every accumulator below carries the 1 / erow scale
#pragma unroll
                for (int t = 0; t < D; ++t) narow[r] = fma(ar[r][t], ar[r][t], narow[r]);
                erow[r] = exp2neg16(narow[r], tab_s);
            }
        }
        if constexpr (CAN_XDIST) {
            if (xdist) {
#pragma unroll
                for (int t = 0; t < D; ++t) narow[r] = fma(ar[r][t], ar[r][t], narow[r]);
            }
        }
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
    const int n_off0 = til.n_off0();
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(n_off0, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        bool mine[NC], any = false;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            int bb;
            mine[c] = pl.block_ok(c) && til.partner(pl.a_of(c), SPREAD ? off0 - c : off0, bb);
            any = any || mine[c];
        }
        if (!any) continue;
        const int b = (pl.a0 + off0) % lay.k;
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        T grow[R], l1[R][NLS];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            grow[r] = (mine[SPREAD ? r : 0] && valid[r]) ? p.Gt[(int64_t)b * p.ldg + irow[r]] : T(0);  // G[i, b]
            grow[r] *= erow[r];  // expanded form: the pair values are K' / erow (erow == 1 otherwise)
#pragma unroll
            for (int t = 0; t < NLS; ++t) l1[r][t] = T(0);
        }
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            for (int idx = tid; idx < clp; idx += nthreads) {
                const bool in = idx < cl;
                const int64_t j = j0 + c0 + (in ? idx : 0);
                T nb = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T v = in ? p.xs[t * p.ld + j] : T(0);
                    nb = fma(v, v, nb);
                    xc[t * JC + idx] = (expand || xdist) ? T(-2) * v : v;
                }
                nbc[idx] = nb;
                sc[idx] = in ? p.s[j] : T(0);
#pragma unroll
                for (int c = 0; c < NC; ++c)  // column weights G[j, a_c]; zero for idle row blocks and (diagonal
                    gc[c * JC + idx] = (in && mine[c]) ? p.Gt[(int64_t)pl.a_of(c) * p.ldg + j] : T(0);  // handled below)
            }
            __syncthreads();
            auto sweep = [&](auto mode_c) {
                // 0 general, 1 safe range, 2 safe + expanded RBF form, 3 safe + expanded distance (Matern-3/2, -5/2)
                constexpr int MODE = decltype(mode_c)::value;
                constexpr bool SAFE = MODE >= 1;
                auto pair_step = [&](int cidx, T (&cacc)[NC]) {
                    T xb[D];
#pragma unroll
                    for (int tt = 0; tt < D; ++tt) xb[tt] = xc[tt * JC + cidx];
                    const T scol = sc[cidx];
                    if constexpr (MODE == 2) {
                        const T nb = nbc[cidx];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const T gcol_row = gc[cidx];
                            const T gcol = (off0 == 0) ? T(0) : gcol_row;
                            T u = nb;  // |b|^2 - 2 a.b  (xb holds -2 b)
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) u = fma(ar[r][tt], xb[tt], u);
                            const T kv = exp2neg16(u, tab_s);  // K' / erow; RBF: H = K'
                            const T hq = kv * (narow[r] + u);
                            sacc[r] = fma(kv, gcol_row, sacc[r]);
                            cacc[0] = fma(kv, grow[r], cacc[0]);
                            l1[r][0] = fma(hq, scol, l1[r][0]);
                            l2[r][0] = fma(hq, gcol, l2[r][0]);
                        }
                        return;
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int c = SPREAD ? r : 0;
                        const T gcol_row = gc[c * JC + cidx];
                        // diagonal tile (off == 0): ordered pairs, the row part only
                        const T gcol = ((SPREAD ? off0 - c : off0) == 0) ? T(0) : gcol_row;
                        T q = T(0);
                        T qt[D];
                        if (ARD) {
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) {
                                const T df = ar[r][tt] - xb[tt];
                                qt[tt] = df * df;
                                q += qt[tt];
                            }
                        } else if constexpr (MODE == 3) {
                            T u = nbc[cidx];  // |b|^2 - 2 a.b  (xb holds -2 b)
#pragma unroll
                            for (int tt = 0; tt < D; ++tt) u = fma(ar[r][tt], xb[tt], u);
                            q = fmax(narow[r] + u, T(0));  // the weight H q must not go negative by rounding
                        } else {
                            q = sqdist<T, D>(ar[r], xb);
                        }
                        T kv, h;
                        kernel_value_grad<T, FAM, SAFE>(q, tab_s, kv, h);
                        sacc[r] = fma(kv, gcol_row, sacc[r]);
                        cacc[c] = fma(kv, grow[r], cacc[c]);
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
                };
                for (int sub = 0; sub < clp; sub += 32) {
                    const int rem = cl - sub;
                    if (rem > kSymTailBroadcast) {
                        T cacc[NC];
#pragma unroll
                        for (int c = 0; c < NC; ++c) cacc[c] = T(0);
#pragma unroll 2
                        for (int t = 0; t < 32; ++t) {
                            pair_step(sub + ((lane + t) & 31), cacc);
#pragma unroll
                            for (int c = 0; c < NC; ++c) cacc[c] = shfl_next<T>(cacc[c], next_lane);
                        }
#pragma unroll
                        for (int c = 0; c < NC; ++c) colacc[(c * NWMAX + warp) * JC + sub + lane] = cacc[c];
                    } else {
                        for (int jj = 0; jj < rem; ++jj) {
                            T cp[NC];
#pragma unroll
                            for (int c = 0; c < NC; ++c) cp[c] = T(0);
                            pair_step(sub + jj, cp);
#pragma unroll
                            for (int c = 0; c < NC; ++c) {
                                T v = cp[c];
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                                if (lane == 0) colacc[(c * NWMAX + warp) * JC + sub + jj] = v;
                            }
                        }
                    }
                }
            };
            if constexpr (sizeof(T) == 8) {
                if constexpr (CAN_EXPAND) {
                    if (expand) sweep(std::integral_constant<int, 2>{});
                    else if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                } else if constexpr (CAN_XDIST) {
                    if (xdist) sweep(std::integral_constant<int, 3>{});
                    else if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                } else {
                    if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                }
            } else {
                sweep(std::integral_constant<int, 0>{});
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                if (!mine[c] || (SPREAD ? off0 - c : off0) == 0) continue;
                for (int idx = tid; idx < cl; idx += nthreads) {
                    T v = T(0);
                    for (int w = 0; w < nw; ++w) v += colacc[(c * NWMAX + w) * JC + idx];
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
        if (valid[r]) atomicAdd(p.sbar + irow[r], sacc[r] * erow[r]);
#pragma unroll
        for (int t = 0; t < NLS; ++t) ltot[t] = fma(srow[r] * erow[r], l2[r][t], ltot[t]);
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
        for (int w = 0; w < nw; ++w) v += red[w][tid];
        p.ls_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NLS + tid] = v;
    }
}

}  // namespace cagp
