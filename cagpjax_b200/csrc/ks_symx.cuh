// Fast path of the symmetric (Gram) block-sparse kernels for the headline instance: float64, RBF, isotropic
// lengthscale, SINGLE CTA shape, expanded distance form (kernel_math.cuh: launch_is_expandable).  sm_100a.
//
// Same tile schedule, same ring walk and same outputs as ks_sym.cuh (which stays the general kernel: every other
// family / dtype / shape, and this instance whenever the bounding box does not prove |x|^2 < 300).  What differs
// is everything AROUND the fp64 instructions, which is what kept the general kernel at ~70 % of the fp64 pipe
// (profiles/r01_ncu_summary.md: ~10 non-fp64 instructions beside 11-16 fp64 ones per pair):
//
//   * R = 8 rows per thread on 128-thread CTAs (4 warps): the per-column work - shared-memory loads of the column,
//     the ring shuffle of the column partial, loop control - is shared by 8 pairs instead of 4;
//   * the staged column is PACKED as 16-byte pairs {-2 x_0, -2 x_1}, {-2 x_2, |x|^2}, {s, g} and read with LDS.128:
//     3 loads per column step instead of 5-6 LDS.64;
//   * every 32-column group is staged TWICE back to back (64 slots), so lane l reads slot l + t at ring step t:
//     the address is one per-lane register plus an immediate - no per-step index arithmetic;
//   * the per-row factor 2^(-|a|^2) is recomputed at the end of a tile instead of living in registers;
//   * the R kernel values of a column step are evaluated stage by stage (symx_exp2_vec) and the accumulations are
//     grouped by kind, so that instructions sharing a register operand sit back to back and the operand-reuse cache
//     serves them: a B200 sub-partition reads only ~two 32-bit register operands per cycle, an fp64 instruction with
//     three fresh 64-bit sources costs 3 cycles instead of 2-2.2 (profiles/r02_rf_probe.txt, r02_pair_kernels.md).
// Measured at C3 (n = 1M, k = 1024, d = 3): forward 488.7 -> 445.5 ms, backward 632 -> 585 ms.
//
// Launch protocol (sync-free): the launcher starts BOTH kernels; this one returns at once unless the bounding box
// proves the expanded form valid, the general one returns at once when it does (SymFwdArgs::defer_expandable).
#pragma once
#include "ks_sym.cuh"

namespace cagp {

constexpr int kSymxOptDadd = 1;     // range reduction by two dependent DADDs instead of I2F + FMA
constexpr int kSymxOptNoAtomic = 4; // EXPERIMENT ONLY (wrong results): skip the column atomics, to time them
constexpr int kSymxOptGroup = 8;    // evaluate the R kernel values of a column step first, then accumulate grouped by kind
// EXPERIMENT ONLY (wrong results), in-situ cost of the pieces of a pair step:
constexpr int kSymxAblColAdd = 16;  // column partial by DADD (2 fresh operands) instead of DFMA (3 fresh)
constexpr int kSymxAblNoTable = 32; // no exp2 table lookup / exponent adjust (LOP, 2 IMAD, LDS)
constexpr int kSymxAblNoShfl = 64;  // no ring shuffle
constexpr int kSymxAblNoPoly = 128; // no cubic
constexpr int kSymxAblNoRow = 256;  // no row accumulation
constexpr int kSymxOptUnroll1 = 512;   // ring loop not unrolled (default: 8 steps)
constexpr int kSymxOptUnroll2 = 1024;  // ring loop unrolled by 2
constexpr int kSymxOptUnroll4 = 2048;  // ring loop unrolled by 4
template <int OPT>
struct SymxUnroll {
    static constexpr int value = (OPT & kSymxOptUnroll1) ? 1 : (OPT & kSymxOptUnroll2) ? 2 : (OPT & kSymxOptUnroll4) ? 4 : 8;
};

// 2^(-u), |u| < 2000, from the 256-cell x 16-copy table (kernel_math.cuh: exptab16_load).
template <int OPT>
__device__ __forceinline__ double symx_exp2(double u, unsigned tab_lane) {
    constexpr double kMagic = 26388279066624.0;  // 1.5 * 2^44: ulp 2^-8
    const double t = u + kMagic;
    const unsigned lo = (unsigned)__double2loint(t);
    const double f = (OPT & kSymxOptDadd) ? u - (t - kMagic) : fma(__int2double_rn((int)lo), -0.00390625, u);
    if (OPT & kSymxAblNoTable) {
        constexpr double c0 = 0.999999999999982504724, c1 = -0.693147180559942884057, c2 = 0.2402265436493534809523,
                         c3 = -0.05550411375117055438833;
        return fma(f, fma(f, fma(f, c3, c2), c1), c0) * t;
    }
    unsigned idx, addr;
    asm("and.b32 %0, %1, 255;" : "=r"(idx) : "r"(lo));
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(idx), "r"(tab_lane));
    double tv;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(addr));
    constexpr double c0 = 0.999999999999982504724;
    constexpr double c1 = -0.693147180559942884057;
    constexpr double c2 = 0.2402265436493534809523;
    constexpr double c3 = -0.05550411375117055438833;
    const double p = (OPT & kSymxAblNoPoly) ? f : fma(f, fma(f, fma(f, c3, c2), c1), c0);
    int thi;
    asm("mad.lo.s32 %0, %1, -4096, %2;" : "=r"(thi) : "r"(lo), "r"(__double2hiint(tv)));
    return __hiloint2double(thi, __double2loint(tv)) * p;
}

// stage-ordered exp2 of R arguments: kernel_math.cuh (exp2neg16_vec); the ablation flags of OPT are forwarded
template <int OPT, int R, bool CLAMP>
__device__ __forceinline__ void symx_exp2_vec(const double (&u)[R], unsigned tab_lane, double (&kv)[R]) {
    if constexpr ((OPT & (kSymxAblNoTable | kSymxAblNoPoly | kSymxOptDadd)) != 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) kv[r] = symx_exp2<OPT>(CLAMP ? fmin(u[r], 1000.0) : u[r], tab_lane);
    } else {
        exp2neg16_vec<R, CLAMP>(u, tab_lane, kv);  // CLAMP: integer clamp of the exponent adjustment
    }
}

// Values of a staged column: v[0..D-1] = -2 x_t, v[D] = |x|^2, v[D+1] = s, v[D+2] = g (backward only).
template <int D, bool BWD>
struct SymxCols {
    static constexpr int NV = D + (BWD ? 3 : 2);
    static constexpr int NP = (NV + 1) / 2;  // 16-byte packs per column
};

// Shared-memory layout of a stage: NPF = NV / 2 arrays of 16-byte pairs [2 JC], then - when NV is odd (forward, D odd:
// the weight s) - one array of 8-byte scalars [2 JC].  A half-filled pair array would be read as LDS.64 with a 16-byte
// lane stride, a 2-way bank conflict (ncu: 2.65e8 conflicts per c3s forward launch before this split).
template <int D, bool BWD, int JC>
__device__ __forceinline__ void symx_load_col(const double2* __restrict__ slot, const double* __restrict__ sslot,
                                              double (&v)[2 * SymxCols<D, BWD>::NP]) {
    constexpr int NV = SymxCols<D, BWD>::NV, NPF = NV / 2;
#pragma unroll
    for (int q = 0; q < NPF; ++q) {
        const double2 t = slot[q * 2 * JC];
        v[2 * q] = t.x;
        v[2 * q + 1] = t.y;
    }
    if (NV & 1) {
        v[NV - 1] = *sslot;  // same slot index in the scalar array
        v[NV] = 0.0;
    }
}

// Stage columns [c0, c0 + cl) of block b: slot (g * 64 + c) and (g * 64 + 32 + c) both hold column g * 32 + c.
template <int D, bool BWD, int JC, int NT>
__device__ __forceinline__ void symx_stage(double2* __restrict__ pk, const double* __restrict__ xs, int64_t ld,
                                           const double* __restrict__ s, const double* __restrict__ grow_src,
                                           int64_t jbase, int cl, int clp, int tid) {
    constexpr int NP = SymxCols<D, BWD>::NP, NV = SymxCols<D, BWD>::NV, NPF = NV / 2;
    for (int idx = tid; idx < clp; idx += NT) {
        const bool in = idx < cl;
        const int64_t j = jbase + (in ? idx : 0);
        double v[2 * NP];
        double nb = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            const double x = in ? xs[t * ld + j] : 0.0;
            nb = fma(x, x, nb);
            v[t] = -2.0 * x;
        }
        v[D] = nb;
        v[D + 1] = in ? s[j] : 0.0;
        if (BWD) v[D + 2] = (in && grow_src != nullptr) ? grow_src[j] : 0.0;
        if (NV < 2 * NP) v[NV] = 0.0;
        const int slot = ((idx >> 5) << 6) + (idx & 31);
#pragma unroll
        for (int q = 0; q < NPF; ++q) {
            const double2 t = make_double2(v[2 * q], v[2 * q + 1]);
            pk[q * 2 * JC + slot] = t;
            pk[q * 2 * JC + slot + 32] = t;
        }
        if (NV & 1) {
            double* sc = reinterpret_cast<double*>(pk + NPF * 2 * JC);
            sc[slot] = v[NV - 1];
            sc[slot + 32] = v[NV - 1];
        }
    }
}

template <int D, bool BWD, int JC, int NT>
constexpr size_t symx_smem_bytes() {
    constexpr int NV = SymxCols<D, BWD>::NV;
    return (size_t)kExpTab16Elems * 8 + (size_t)(NV / 2) * 2 * JC * 16 + (size_t)(NV & 1) * 2 * JC * 8 + (size_t)(NT / 32) * JC * 8;
}

template <int D, int R, int NT, int JC, int MINB, int OPT>
__global__ void __launch_bounds__(NT, MINB) ks_symx_fwd_kernel(const SymFwdArgs<double> p) {
    constexpr int NW = NT / 32;
    constexpr int NP = SymxCols<D, false>::NP;
    const int level = launch_expand_level<double, D, RBF>(p.bbox);  // uniform
    if (level == 0) return;                                          // the general kernel does this launch
    const bool clamp = level == 2;  // |x|^2 up to 900: u is clamped at 1000 before the exp2 (kernel_math.cuh)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);
    double2* pk = reinterpret_cast<double2*>(tab + kExpTab16Elems);  // [NV / 2][2 JC] pairs, then [NV & 1][2 JC] scalars
    const double* sc = reinterpret_cast<const double*>(pk + (SymxCols<D, false>::NV / 2) * 2 * JC);
    double* colacc = reinterpret_cast<double*>(smem_raw + symx_smem_bytes<D, false, JC, NT>() - (size_t)NW * JC * 8);  // [NW][JC]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = exptab16_load(tab, tid, NT);
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymRows<R, false> pl = SymRows<R, false>::make(til, blockIdx.x, tid, NT);
    const int a = pl.a0;
    const int64_t len_a = lay.len(a), start_a = lay.start(a);

    double ar[R][D], srow[R];
    unsigned vmask = 0;
    __syncthreads();  // the table is used below
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool ok = pl.o[r] < len_a;
        vmask |= ok ? (1u << r) : 0u;
        const int64_t i = start_a + (ok ? pl.o[r] : len_a - 1);
        double na = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            ar[r][t] = p.xs[t * p.ld + i];
            na = fma(ar[r][t], ar[r][t], na);
        }
        srow[r] = ok ? p.s[i] * symx_exp2<OPT>(na, tab_s) : 0.0;  // column partials take s_i 2^(-|a|^2)
    }
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(til.n_off, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    double* __restrict__ row_panel = p.panel[p.me] - p.row_begin[p.me];
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        int b;
        if (!til.partner(a, off0, b)) continue;
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        int owner = 0;  // rank that owns the rows of column block b
        while (owner + 1 < p.world && j0 >= p.row_begin[owner + 1]) ++owner;
        double* __restrict__ col_panel = p.panel[owner] - p.row_begin[owner];
        double racc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) racc[r] = 0.0;
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            symx_stage<D, false, JC, NT>(pk, p.xs, p.ld, p.s, nullptr, j0 + c0, cl, clp, tid);
            __syncthreads();
            auto pair_step = [&](auto clamp_c, const double2* __restrict__ slot, double& cacc) {
                constexpr bool CLAMP = decltype(clamp_c)::value;
                double v[2 * NP];
                symx_load_col<D, false, JC>(slot, sc + (slot - pk), v);
                if constexpr ((OPT & kSymxOptGroup) != 0) {
                    double u[R], kv[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) u[r] = v[D];
#pragma unroll
                    for (int t = 0; t < D; ++t)
#pragma unroll
                        for (int r = 0; r < R; ++r) u[r] = fma(ar[r][t], v[t], u[r]);
                    symx_exp2_vec<OPT, R, CLAMP>(u, tab_s, kv);
                    if (!(OPT & kSymxAblNoRow)) {
#pragma unroll
                        for (int r = 0; r < R; ++r) racc[r] = fma(kv[r], v[D + 1], racc[r]);
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) cacc = (OPT & kSymxAblColAdd) ? cacc + kv[r] : fma(kv[r], srow[r], cacc);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        double u = v[D];  // |b|^2 - 2 a.b
#pragma unroll
                        for (int t = 0; t < D; ++t) u = fma(ar[r][t], v[t], u);
                        if (CLAMP) u = fmin(u, 1000.0);
                        const double kv = symx_exp2<OPT>(u, tab_s);  // K' / 2^(-|a|^2)
                        if (!(OPT & kSymxAblNoRow)) racc[r] = fma(kv, v[D + 1], racc[r]);
                        cacc = (OPT & kSymxAblColAdd) ? cacc + kv : fma(kv, srow[r], cacc);
                    }
                }
            };
            auto sweep = [&](auto clamp_c) {
                for (int sub = 0; sub < clp; sub += 32) {
                    const int rem = cl - sub;
                    const double2* __restrict__ gp = pk + 2 * sub;
                    if (rem > kSymTailBroadcast) {
                        const double2* __restrict__ lp = gp + lane;
                        double cacc = 0.0;
#pragma unroll SymxUnroll<OPT>::value
                        for (int t = 0; t < 32; ++t) {
                            pair_step(clamp_c, lp + t, cacc);
                            if (!(OPT & kSymxAblNoShfl)) cacc = __shfl_sync(0xffffffffu, cacc, next_lane);
                        }
                        colacc[warp * JC + sub + lane] = cacc;
                    } else {
                        for (int jj = 0; jj < rem; ++jj) {
                            double cp = 0.0;
                            pair_step(clamp_c, gp + jj, cp);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, o);
                            if (lane == 0) colacc[warp * JC + sub + jj] = cp;
                        }
                    }
                }
            };
            if (clamp) sweep(std::true_type{});
            else sweep(std::false_type{});
            __syncthreads();
            if (off0 != 0) {  // column sums: Mt[a][j in b]
                for (int idx = tid; idx < cl; idx += NT) {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) v += colacc[w * JC + idx];
                    double* dst = col_panel + (int64_t)a * p.ldm + j0 + c0 + idx;
                    if (pl.split) atomicAdd(dst, v);
                    else *dst = v;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (vmask & (1u << r)) {
                double na = 0.0;
#pragma unroll
                for (int t = 0; t < D; ++t) na = fma(ar[r][t], ar[r][t], na);
                row_panel[(int64_t)b * p.ldm + start_a + pl.o[r]] = racc[r] * symx_exp2<OPT>(na, tab_s);
            }
        }
    }
}

// Backward.  With e_i = 2^(-|a_i|^2), kv = K'_ij / e_i, u = |b_j|^2 - 2 a_i.b_j (so q_ij = |a_i|^2 + u):
//   s_bar_i += e_i sum_j kv g_j            g_j = Gt[a][j]      (row part)
//   s_bar_j += sum_i kv (e_i G[i, b])                          (column part, ring)
//   ls_bar  += sum_ij K'_ij q_ij (G[i, b] s_j + g_j s_i)       (g_j := 0 on the diagonal tile: ordered pairs there)
// (Accumulating kv u and restoring the |a_i|^2 part from Mt was measured SLOWER: 669 vs 600 ms, r02_sweep4.log.)
template <int D, int R, int NT, int JC, int MINB, int OPT>
__global__ void __launch_bounds__(NT, MINB) ks_symx_bwd_kernel(const SymBwdArgs<double> p) {
    constexpr int NW = NT / 32;
    constexpr int NP = SymxCols<D, true>::NP;
    const int level = launch_expand_level<double, D, RBF>(p.bbox);  // uniform, see the forward kernel
    if (level == 0) return;
    const bool clamp = level == 2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* tab = reinterpret_cast<double*>(smem_raw);
    double2* pk = reinterpret_cast<double2*>(tab + kExpTab16Elems);  // [NV / 2][2 JC] pairs, then [NV & 1][2 JC] scalars
    const double* sc = reinterpret_cast<const double*>(pk + (SymxCols<D, true>::NV / 2) * 2 * JC);
    double* colacc = reinterpret_cast<double*>(smem_raw + symx_smem_bytes<D, true, JC, NT>() - (size_t)NW * JC * 8);  // [NW][JC]
    __shared__ double red[NW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = exptab16_load(tab, tid, NT);
    const SymTiling& til = p.til;
    const BlockLayout& lay = til.lay;
    const SymRows<R, false> pl = SymRows<R, false>::make(til, blockIdx.x, tid, NT);
    const int a = pl.a0;
    const int64_t len_a = lay.len(a), start_a = lay.start(a);

    double ar[R][D], narow[R], sacc[R], l2[R];
    unsigned vmask = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const bool ok = pl.o[r] < len_a;
        vmask |= ok ? (1u << r) : 0u;
        const int64_t i = start_a + (ok ? pl.o[r] : len_a - 1);
        narow[r] = 0.0;
#pragma unroll
        for (int t = 0; t < D; ++t) {
            ar[r][t] = p.xs[t * p.ld + i];
            narow[r] = fma(ar[r][t], ar[r][t], narow[r]);
        }
        sacc[r] = 0.0;
        l2[r] = 0.0;
    }
    double ltot = 0.0;
    const int off_begin = blockIdx.y * til.offs_per_cta;
    const int off_end = min(til.n_off, off_begin + til.offs_per_cta);
    const int next_lane = (lane + 1) & 31;
    for (int off0 = off_begin; off0 < off_end; ++off0) {
        int b;
        if (!til.partner(a, off0, b)) continue;
        const int64_t j0 = lay.start(b), len_b = lay.len(b);
        double grow[R], l1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const bool ok = (vmask >> r) & 1u;
            // G[i, b] e_i: the pair values are K' / e_i
            grow[r] = ok ? p.Gt[(int64_t)b * p.ldg + start_a + pl.o[r]] * symx_exp2<OPT>(narow[r], tab_s) : 0.0;
            l1[r] = 0.0;
        }
        for (int64_t c0 = 0; c0 < len_b; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len_b - c0);
            const int clp = (cl + 31) & ~31;
            __syncthreads();
            symx_stage<D, true, JC, NT>(pk, p.xs, p.ld, p.s, p.Gt + (int64_t)a * p.ldg, j0 + c0, cl, clp, tid);
            __syncthreads();
            auto sweep = [&](auto diag_c, auto clamp_c) {
                constexpr bool DIAG = decltype(diag_c)::value;
                constexpr bool CLAMP = decltype(clamp_c)::value;
                auto pair_step = [&](const double2* __restrict__ slot, double& cacc) {
                    double v[2 * NP];
                    symx_load_col<D, true, JC>(slot, sc + (slot - pk), v);
                    if constexpr ((OPT & kSymxOptGroup) != 0) {
                        double u[R], kv[R], hq[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) u[r] = v[D];
#pragma unroll
                        for (int t = 0; t < D; ++t)
#pragma unroll
                            for (int r = 0; r < R; ++r) u[r] = fma(ar[r][t], v[t], u[r]);
                        // CLAMP: only kv is clamped; the weight q = |a|^2 + u keeps the true u (kv < 1e-300 there)
                        symx_exp2_vec<OPT, R, CLAMP>(u, tab_s, kv);
#pragma unroll
                        for (int r = 0; r < R; ++r) hq[r] = kv[r] * (narow[r] + u[r]);
#pragma unroll
                        for (int r = 0; r < R; ++r) sacc[r] = fma(kv[r], v[D + 2], sacc[r]);
#pragma unroll
                        for (int r = 0; r < R; ++r) l1[r] = fma(hq[r], v[D + 1], l1[r]);
                        if (!DIAG) {
#pragma unroll
                            for (int r = 0; r < R; ++r) l2[r] = fma(hq[r], v[D + 2], l2[r]);
                        }
#pragma unroll
                        for (int r = 0; r < R; ++r) cacc = fma(kv[r], grow[r], cacc);
                    } else {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            double u = v[D];
#pragma unroll
                            for (int t = 0; t < D; ++t) u = fma(ar[r][t], v[t], u);
                            const double kv = symx_exp2<OPT>(CLAMP ? fmin(u, 1000.0) : u, tab_s);
                            const double hq = kv * (narow[r] + u);
                            sacc[r] = fma(kv, v[D + 2], sacc[r]);
                            cacc = fma(kv, grow[r], cacc);
                            l1[r] = fma(hq, v[D + 1], l1[r]);
                            if (!DIAG) l2[r] = fma(hq, v[D + 2], l2[r]);
                        }
                    }
                };
                for (int sub = 0; sub < clp; sub += 32) {
                    const int rem = cl - sub;
                    const double2* __restrict__ gp = pk + 2 * sub;
                    if (rem > kSymTailBroadcast) {
                        const double2* __restrict__ lp = gp + lane;
                        double cacc = 0.0;
#pragma unroll SymxUnroll<OPT>::value
                        for (int t = 0; t < 32; ++t) {
                            pair_step(lp + t, cacc);
                            cacc = __shfl_sync(0xffffffffu, cacc, next_lane);
                        }
                        colacc[warp * JC + sub + lane] = cacc;
                    } else {
                        for (int jj = 0; jj < rem; ++jj) {
                            double cp = 0.0;
                            pair_step(gp + jj, cp);
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) cp += __shfl_xor_sync(0xffffffffu, cp, o);
                            if (lane == 0) colacc[warp * JC + sub + jj] = cp;
                        }
                    }
                }
            };
            if (clamp) {
                if (off0 == 0) sweep(std::true_type{}, std::true_type{});
                else sweep(std::false_type{}, std::true_type{});
            } else {
                if (off0 == 0) sweep(std::true_type{}, std::false_type{});
                else sweep(std::false_type{}, std::false_type{});
            }
            __syncthreads();
            if (off0 != 0 && !(OPT & kSymxOptNoAtomic)) {
                for (int idx = tid; idx < cl; idx += NT) {
                    double v = 0.0;
#pragma unroll
                    for (int w = 0; w < NW; ++w) v += colacc[w * JC + idx];
                    atomicAdd(p.sbar + j0 + c0 + idx, v);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            ltot = fma(grow[r], l1[r], ltot);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if ((vmask >> r) & 1u) {
            const int64_t i = start_a + pl.o[r];
            const double e = symx_exp2<OPT>(narow[r], tab_s);
            const double se = p.s[i] * e;
            atomicAdd(p.sbar + i, sacc[r] * e);
            ltot = fma(se, l2[r], ltot);
        }
    }
    double v = ltot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += red[w];
        p.ls_part[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
}

}  // namespace cagp
