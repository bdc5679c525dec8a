// float64 contractions of the k-sized algebra on the fp64 tensor core (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4;
// tcgen05 has no f64 kind), replacing the three cuBLAS Dgemm calls of round 1.  sm_100a.
//
//   gram_syrk_kernel     B' = M'^T M'   (k x k from Mt (k, m)): ONLY the lower-triangular 128 x 128 tiles are computed
//                        (n k^2 flops instead of 2 n k^2), the long m dimension is split over CTAs, partial tiles go to
//                        a workspace and gram_reduce_kernel sums them in a fixed order and mirrors (deterministic).
//                        Replaces solvers/cholesky.py:57-63 + distributions.py:53-56 (cola.diag of the lazy chain).
//   gemm_wm_kernel<COT>  Gt = W Mt + outer terms: the cotangent of M' (reverse mode of the above); the prefill pass of
//                        round 1 (a full write + re-read of the 8 n k byte panel) is the epilogue now.
//   gemm_wm_kernel<MOM>  predictive moments: Z = Pinv Mt is never written - the epilogue contracts it with Mt
//                        (var_i = sum_c Mt[c][i] Z[c][i]) and with w (mean), per 128-row tile, into 2 x (k/128) x m partials.
//
// One CTA = 128 x 128 output tile, 512 threads = 4 x 4 warps, warp tile 32 x 32 = 4 x 4 m8n8 accumulators (64
// registers), k-chunks of 32 double-buffered with cp.async.  (First version: 256 threads with 64 x 32 warp
// tiles - two warps per sub-partition could not keep the DMMA pipe busy: one warp issues a DMMA every ~38 cycles,
// the pipe takes one every 16; ncu: tensor pipe 85 % active, top stall `wait`.  Four warps per sub-partition can.)  Shared-memory tiles are [row][k] with row
// stride 36 doubles or [k][col] with row stride 132: both == 4 (mod 16), so the 16 lanes of a half-warp that load one
// fragment (4 k-values x 4 rows / columns, 8 bytes each) hit 16 different bank pairs.  Per k-step of 4 a warp issues
// 8 LDS.64 for 16 DMMA (256 fp64-pipe cycles): the pipe, not shared memory (25 % busy), is the limit.
#pragma once
#include "ks_dense.cuh"

namespace cagp {

constexpr int kGemmTile = 128;
constexpr int kGemmKC = 32;     // k-chunk per pipeline stage (one CTA-wide barrier per chunk: 128 DMMA per warp between barriers)
constexpr int kGemmStages = 2;
constexpr int kGemmThreads = 512;
constexpr int kGemmWM = 4, kGemmWN = 4;  // warp grid; warp tile 32 x 32
constexpr int kGemmMT = 4, kGemmNT = 4;  // m8n8 accumulator tiles per warp
constexpr int kGemmRS = kGemmKC + 4;     // row stride of [row][k] tiles
constexpr int kGemmCS = kGemmTile + 4;   // row stride of [k][col] tiles
constexpr int kGemmRowKElems = kGemmTile * kGemmRS;   // 2560 doubles
constexpr int kGemmKColElems = kGemmKC * kGemmCS;     // 2112 doubles

// 16-byte asynchronous copy of two doubles; `nvalid` (0, 1 or 2) leading elements are read, the rest is zero-filled
__device__ __forceinline__ void cp_async_2d(double* smem_dst, const double* gsrc, int nvalid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(gsrc), "r"(nvalid * 8) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// T[r][kk] = src[(r0 + r) * ld + k0 + kk] for r < 128, kk < kGemmKC (zero outside rows / kend).  V16: 16-byte copies
// (needs src 16-byte aligned, ld and k0 even).
template <bool V16>
__device__ __forceinline__ void gemm_load_rowk(double* __restrict__ dst, const double* __restrict__ src, int64_t ld,
                                               int64_t r0, int64_t rows, int64_t k0, int64_t kend, int tid) {
    if (V16) {
#pragma unroll
        for (int e = 0; e < kGemmTile * kGemmKC / 2 / kGemmThreads; ++e) {
            const int idx = tid + kGemmThreads * e, r = idx / (kGemmKC / 2), kk = (idx % (kGemmKC / 2)) * 2;
            const bool rok = r0 + r < rows;
            const int nv = rok ? (int)max((int64_t)0, min((int64_t)2, kend - (k0 + kk))) : 0;
            cp_async_2d(dst + r * kGemmRS + kk, src + (nv ? (r0 + r) * ld + k0 + kk : 0), nv);
        }
    } else {
#pragma unroll
        for (int e = 0; e < kGemmTile * kGemmKC / kGemmThreads; ++e) {
            const int idx = tid + kGemmThreads * e, r = idx / kGemmKC, kk = idx % kGemmKC;
            const bool ok = r0 + r < rows && k0 + kk < kend;
            cp_async_elem<double>(dst + r * kGemmRS + kk, src + (ok ? (r0 + r) * ld + k0 + kk : 0), ok);
        }
    }
}

// T[kk][c] = src[(k0 + kk) * ld + c0 + c] for kk < kGemmKC, c < 128 (zero outside kend / cols)
template <bool V16>
__device__ __forceinline__ void gemm_load_kcol(double* __restrict__ dst, const double* __restrict__ src, int64_t ld,
                                               int64_t k0, int64_t kend, int64_t c0, int64_t cols, int tid) {
    if (V16) {
#pragma unroll
        for (int e = 0; e < kGemmTile * kGemmKC / 2 / kGemmThreads; ++e) {
            const int idx = tid + kGemmThreads * e, kk = idx >> 6, c = (idx & 63) * 2;
            const bool kok = k0 + kk < kend;
            const int nv = kok ? (int)max((int64_t)0, min((int64_t)2, cols - (c0 + c))) : 0;
            cp_async_2d(dst + kk * kGemmCS + c, src + (nv ? (k0 + kk) * ld + c0 + c : 0), nv);
        }
    } else {
#pragma unroll
        for (int e = 0; e < kGemmTile * kGemmKC / kGemmThreads; ++e) {
            const int idx = tid + kGemmThreads * e, kk = idx >> 7, c = idx & 127;
            const bool ok = k0 + kk < kend && c0 + c < cols;
            cp_async_elem<double>(dst + kk * kGemmCS + c, src + (ok ? (k0 + kk) * ld + c0 + c : 0), ok);
        }
    }
}

// acc (warp tile 32 x 32) += A-tile [128][16] (row-k layout) x B-tile (row-k layout [128][16] if !B_KCOL, else [16][128])
template <bool B_KCOL>
__device__ __forceinline__ void gemm_mma_stage(const double* __restrict__ As, const double* __restrict__ Bs,
                                               double (&acc)[kGemmMT][kGemmNT][2], int wm, int wn, int grp, int tig) {
#pragma unroll
    for (int ks = 0; ks < kGemmKC / 4; ++ks) {
        double af[kGemmMT], bf[kGemmNT];
        const double* ap = As + (wm * 32 + grp) * kGemmRS + ks * 4 + tig;
#pragma unroll
        for (int mt = 0; mt < kGemmMT; ++mt) af[mt] = ap[mt * 8 * kGemmRS];
        if (B_KCOL) {
            const double* bp = Bs + (ks * 4 + tig) * kGemmCS + wn * 32 + grp;
#pragma unroll
            for (int nt = 0; nt < kGemmNT; ++nt) bf[nt] = bp[nt * 8];
        } else {
            const double* bp = Bs + (wn * 32 + grp) * kGemmRS + ks * 4 + tig;
#pragma unroll
            for (int nt = 0; nt < kGemmNT; ++nt) bf[nt] = bp[nt * 8 * kGemmRS];
        }
#pragma unroll
        for (int mt = 0; mt < kGemmMT; ++mt)
#pragma unroll
            for (int nt = 0; nt < kGemmNT; ++nt) dmma_m8n8k4(acc[mt][nt], af[mt], bf[nt]);
    }
}

// ---------------------------------------------------------------------------------------------------------------
struct SyrkArgs {
    const double* Mt; int64_t ldm; int64_t m; int k;
    double* ws;       // [splits][k][k] partial tiles (only the lower-triangular tiles are written)
    int64_t chunk;    // columns of Mt (rows of M') per split, a multiple of kGemmKC
};

__host__ __device__ inline void tri_tile(int t, int& ta, int& tb) {  // t = ta (ta + 1) / 2 + tb, tb <= ta
    int a = 0;
    while ((a + 1) * (a + 2) / 2 <= t) ++a;
    ta = a;
    tb = t - a * (a + 1) / 2;
}

template <bool V16>
__global__ void __launch_bounds__(kGemmThreads, 1) gram_syrk_kernel(const SyrkArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);  // [stages][2][128][20]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, grp = lane >> 2, tig = lane & 3;
    int ta, tb;
    tri_tile(blockIdx.x, ta, tb);
    const bool diag = ta == tb;
    const int64_t a0 = (int64_t)ta * kGemmTile, b0 = (int64_t)tb * kGemmTile;
    const int64_t i_begin = (int64_t)blockIdx.y * p.chunk, i_end = min(p.m, i_begin + p.chunk);
    const int nk = (int)((i_end - i_begin + kGemmKC - 1) / kGemmKC);
    double acc[kGemmMT][kGemmNT][2];
#pragma unroll
    for (int mt = 0; mt < kGemmMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < kGemmNT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    auto load = [&](int stage, int kt) {
        double* As = sm + stage * 2 * kGemmRowKElems;
        const int64_t k0 = i_begin + (int64_t)kt * kGemmKC;
        gemm_load_rowk<V16>(As, p.Mt, p.ldm, a0, p.k, k0, i_end, tid);
        if (!diag) gemm_load_rowk<V16>(As + kGemmRowKElems, p.Mt, p.ldm, b0, p.k, k0, i_end, tid);
    };
#pragma unroll
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < nk) load(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait_group<kGemmStages - 2>();
        __syncthreads();  // stage kt has landed; everybody is done with the buffer that is refilled next
        if (kt + kGemmStages - 1 < nk) load((kt + kGemmStages - 1) % kGemmStages, kt + kGemmStages - 1);
        cp_async_commit();
        const double* As = sm + (kt % kGemmStages) * 2 * kGemmRowKElems;
        gemm_mma_stage<false>(As, diag ? As : As + kGemmRowKElems, acc, wm, wn, grp, tig);
    }
    double* __restrict__ out = p.ws + (int64_t)blockIdx.y * p.k * p.k;
#pragma unroll
    for (int mt = 0; mt < kGemmMT; ++mt) {
        const int64_t r = a0 + wm * 32 + mt * 8 + grp;
        if (r >= p.k) continue;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int64_t c = b0 + wn * 32 + nt * 8 + 2 * tig + e;
                if (c < p.k) out[r * p.k + c] = acc[mt][nt][e];
            }
    }
}

// B[r][c] = B[c][r] = sum over splits of the partial (r, c), r >= c; fixed summation order
__global__ void gram_reduce_kernel(const double* __restrict__ ws, int splits, int k, double* __restrict__ B) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    if (r >= k || c > r) return;
    double v = 0.0;
    for (int s = 0; s < splits; ++s) v += ws[((int64_t)s * k + r) * k + c];
    B[(int64_t)r * k + c] = v;
    B[(int64_t)c * k + r] = v;
}

// ---------------------------------------------------------------------------------------------------------------
enum GemmEpilogue : int { GEMM_COT = 0, GEMM_MOM = 1, GEMM_COT_PEER = 2 };
constexpr int kGemmMaxPeers = 16;

struct GemmWmArgs {
    const double* W; int k;                     // (k, k) row-major, A operand
    const double* Mt; int64_t ldm; int64_t m;   // (k, m): B operand
    int nbt;                                    // row tiles = ceil(k / 128)
    // COT: Gt[b][i] = sum_c W[b][c] Mt[c][i] + Abar[blk(i)][b] s_i + cbar[b] yt_i
    double* Gt; int64_t ldg;
    const double* Abar; const double* cbar; const double* s; const double* yt;
    int64_t row_offset; BlockLayout lay;
    // MOM: pv[bt][i] = sum_{c in tile bt} Mt[c][i] (W Mt)[c][i],  pm[bt][i] = sum_{c in tile bt} Mt[c][i] w[c]
    const double* w; double* pv; double* pm;
    // COT_PEER (one process per GPU): every rank h holds a FULL-width buffer peer[h] (k, peer_ld >= n_total), indexed by
    // GLOBAL row.  This rank's panel (its rows = columns [row_offset, row_offset + m) of the buffers) is stored into its
    // own buffer, and row b additionally into the buffer of the rank that owns action block b (block_begin[h] <= b <
    // block_begin[h + 1]) - remote st.global over NVLink, which replaces the all_to_all of the cotangent.
    double* peer[kGemmMaxPeers]; int64_t peer_ld; int block_begin[kGemmMaxPeers + 1]; int world, me;
};

template <int EPI, bool V16>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_wm_kernel(const GemmWmArgs p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sm = reinterpret_cast<double*>(smem_raw);  // [stages]{[128][20], [16][132]}
    constexpr int kStageElems = kGemmRowKElems + kGemmKColElems;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2, grp = lane >> 2, tig = lane & 3;
    const int bt = (int)(blockIdx.x % p.nbt);   // row tiles of one column tile run back to back (Mt tile stays in L2)
    const int64_t it = blockIdx.x / p.nbt;
    const int64_t b0 = (int64_t)bt * kGemmTile, i0 = it * kGemmTile;
    const int nk = (p.k + kGemmKC - 1) / kGemmKC;
    double acc[kGemmMT][kGemmNT][2];
#pragma unroll
    for (int mt = 0; mt < kGemmMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < kGemmNT; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
    auto load = [&](int stage, int kt) {
        double* As = sm + stage * kStageElems;
        const int64_t k0 = (int64_t)kt * kGemmKC;
        gemm_load_rowk<V16>(As, p.W, p.k, b0, p.k, k0, p.k, tid);
        gemm_load_kcol<V16>(As + kGemmRowKElems, p.Mt, p.ldm, k0, p.k, i0, p.m, tid);
    };
#pragma unroll
    for (int s = 0; s < kGemmStages - 1; ++s) {
        if (s < nk) load(s, s);
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait_group<kGemmStages - 2>();
        __syncthreads();
        if (kt + kGemmStages - 1 < nk) load((kt + kGemmStages - 1) % kGemmStages, kt + kGemmStages - 1);
        cp_async_commit();
        const double* As = sm + (kt % kGemmStages) * kStageElems;
        gemm_mma_stage<true>(As, As + kGemmRowKElems, acc, wm, wn, grp, tig);
    }
    if (EPI == GEMM_COT || EPI == GEMM_COT_PEER) {
        // a lane holds two consecutive columns (i, i + 1) of 4 x 4 rows: 16-byte stores, 64 contiguous bytes per row
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int64_t i = i0 + wn * 32 + nt * 8 + 2 * tig;
            if (i >= p.m) continue;
            const bool two = i + 1 < p.m;
            const double s0 = p.s[i], y0 = p.yt[i], s1 = two ? p.s[i + 1] : 0.0, y1 = two ? p.yt[i + 1] : 0.0;
            const double* __restrict__ ar0 = p.Abar + (int64_t)p.lay.block_of(p.row_offset + i) * p.k;
            const double* __restrict__ ar1 = p.Abar + (int64_t)p.lay.block_of(p.row_offset + i + (two ? 1 : 0)) * p.k;
#pragma unroll
            for (int mt = 0; mt < kGemmMT; ++mt) {
                const int64_t b = b0 + wm * 32 + mt * 8 + grp;
                if (b >= p.k) continue;
                const double cb = p.cbar[b];
                const double v0 = acc[mt][nt][0] + fma(ar0[b], s0, cb * y0);
                const double v1 = acc[mt][nt][1] + fma(ar1[b], s1, cb * y1);
                auto put = [&](double* dst) {
                    if (V16 && two) {
                        *reinterpret_cast<double2*>(dst) = make_double2(v0, v1);
                    } else {
                        dst[0] = v0;
                        if (two) dst[1] = v1;
                    }
                };
                if (EPI == GEMM_COT) {
                    put(p.Gt + b * p.ldg + i);
                } else {
                    const int64_t off = b * p.peer_ld + p.row_offset + i;
                    put(p.peer[p.me] + off);
                    int h = 0;
                    while (h + 1 < p.world && (int)b >= p.block_begin[h + 1]) ++h;
                    if (h != p.me) put(p.peer[h] + off);  // the owner of block b needs G'[j, b] for all rows j
                }
            }
        }
    } else {
        // per column i of the tile: sum over the tile's rows of Mt[c][i] * Z[c][i] and Mt[c][i] * w[c]
        __syncthreads();  // the pipeline buffers are free: reuse them for the cross-warp sums
        double* red = sm;  // [4 (wm)][2 (var, mean)][128 columns]
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int cl = wn * 32 + nt * 8 + 2 * tig + e;
                const int64_t i = i0 + cl;
                double sv = 0.0, smn = 0.0;
                if (i < p.m) {
#pragma unroll
                    for (int mt = 0; mt < kGemmMT; ++mt) {
                        const int64_t c = b0 + wm * 32 + mt * 8 + grp;
                        if (c < p.k) {
                            const double mv = p.Mt[c * p.ldm + i];
                            sv = fma(mv, acc[mt][nt][e], sv);
                            smn = fma(mv, p.w[c], smn);
                        }
                    }
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {  // over grp (lanes with the same tig)
                    sv += __shfl_xor_sync(0xffffffffu, sv, o);
                    smn += __shfl_xor_sync(0xffffffffu, smn, o);
                }
                if (grp == 0) {
                    red[(wm * 2 + 0) * kGemmTile + cl] = sv;
                    red[(wm * 2 + 1) * kGemmTile + cl] = smn;
                }
            }
        __syncthreads();
        if (tid < kGemmTile) {
            const int64_t i = i0 + tid;
            if (i < p.m) {
                double sv = 0.0, smn = 0.0;
#pragma unroll
                for (int w = 0; w < kGemmWM; ++w) {
                    sv += red[(w * 2 + 0) * kGemmTile + tid];
                    smn += red[(w * 2 + 1) * kGemmTile + tid];
                }
                p.pv[(int64_t)bt * p.m + i] = sv;
                p.pm[(int64_t)bt * p.m + i] = smn;
            }
        }
    }
}

// mean[i] = mu + var * sum_bt pm[bt][i];  v[i] = var - var^2 * sum_bt pv[bt][i];  scalars = {variance, mean_const}
__global__ void predict_finalize_kernel(const double* __restrict__ pv, const double* __restrict__ pm, int nbt, int64_t m,
                                        const double* __restrict__ scalars, double* __restrict__ mean,
                                        double* __restrict__ var) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double va = 0.0, ma = 0.0;
    for (int t = 0; t < nbt; ++t) {
        va += pv[(int64_t)t * m + i];
        ma += pm[(int64_t)t * m + i];
    }
    const double variance = scalars[0], mu = scalars[1];
    mean[i] = fma(variance, ma, mu);
    var[i] = variance - variance * variance * va;
}

}  // namespace cagp
