// Per-pair kernel arithmetic for the CaGP hot path (sm_100a).
//
// The kernels work on PRE-SCALED coordinates (cagp_scale_inputs): with q = sum_t (a_t - b_t)^2,
//   RBF       K' = 2^(-q)                         scale c = sqrt(0.5 log2 e) / l
//   Matern12  K' = 2^(-sqrt q)                    scale c = log2 e / l
//   Matern32  K' = (1 + ln2 sqrt q) 2^(-sqrt q)   scale c = sqrt 3 log2 e / l
//   Matern52  K' = (1 + r + r^2/3) 2^(-sqrt q), r = ln2 sqrt q,  scale c = sqrt 5 log2 e / l
// which restates the GPJax formulas the reference evaluates through
// DenseKernelComputation.cross_covariance (operators/lazy_kernel.py:83-85) with unit variance.
//
// d K'/d lengthscale_t = H * q_t * (family constant / lengthscale_t), q_t = (a_t - b_t)^2, where
//   RBF H = K' (const 2 ln2), Matern12 H = K'/sqrt q (const ln2), Matern32 H = 2^(-sqrt q)
//   (const ln2^2), Matern52 H = (1 + r) 2^(-sqrt q) / 3 (const ln2^2).
//
// fp64 2^(-u): the fp64 pipe is the roofline of this path (no SFU for doubles), so instead of
// libm exp (17 fp64 issue slots) we use a 2048-entry table of 2^(-j/2048) staged in shared
// memory, magic-number range reduction (3 DADD) and a cubic (3 DFMA + 1 DMUL): 7 fp64 slots,
// max error ~1.5 ulp.  fp32 uses MUFU.EX2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cagp {

enum Family : int { RBF = 0, MATERN12 = 1, MATERN32 = 2, MATERN52 = 3 };

constexpr int kExpTabBits = 11;                // 2048 cells of width 2^-11
constexpr int kExpTabSize = 1 << kExpTabBits;

__device__ const double g_exp2_tab[kExpTabSize] = {
#include "exp2_table.inc"
};

template <typename T>
struct ExpTab;

// The table is staged once per CTA in shared memory.  The high word of entry j is stored with the
// bias + (j << 9) so that the exponent adjustment in exp2neg is a single IMAD.
// (Measured alternatives, B200, n = 400k, k = 410, d = 3: this 2048-entry table + cubic 1564 Gpairs/s
//  in the symmetric forward; a 256-entry table replicated 16x to be bank-conflict free + quartic
//  1508 Gpairs/s - the extra DFMA costs more than the ~5 saved shared-memory wavefronts; two
//  interleaved copies of the 2048-entry table (fewer conflicts, same cubic): no gain, 1651 vs 1669.)
template <>
struct ExpTab<double> {
    static constexpr int kSmemElems = kExpTabSize;
    __device__ static unsigned load(double* tab, int tid, int nthreads) {
        for (int i = tid; i < kExpTabSize; i += nthreads) {
            const double v = g_exp2_tab[i];
            tab[i] = __hiloint2double(__double2hiint(v) + (i << 9), __double2loint(v));
        }
        return (unsigned)__cvta_generic_to_shared(tab);
    }
};
// Second layout for the EXPANDED-form RBF kernels, whose fp64 work per pair is small enough that the shared-memory
// pipe (table lookups with random bank conflicts, ~5 wavefronts per warp) becomes the limiter: 256 cells of width
// 2^-8, each replicated 16 times so that lane l reads copy l mod 16.  A 64-bit warp load is served per half-warp,
// and the 16 lanes of a half-warp then hit 16 different bank pairs: always 2 wavefronts (exp2neg16).
constexpr int kExpTab16Bits = 8;
constexpr int kExpTab16Cells = 1 << kExpTab16Bits;
constexpr int kExpTab16Elems = kExpTab16Cells * 16;  // 32 KB of doubles
__device__ __forceinline__ unsigned exptab16_load(double* tab, int tid, int nthreads) {
    for (int i = tid; i < kExpTab16Elems; i += nthreads) {
        const int j = i >> 4;
        const double v = g_exp2_tab[j << (kExpTabBits - kExpTab16Bits)];  // 2^(-j/256)
        tab[i] = __hiloint2double(__double2hiint(v) + (j << 12), __double2loint(v));
    }
    return (unsigned)__cvta_generic_to_shared(tab) + ((unsigned)(tid & 15) << 3);  // per-lane copy
}

template <>
struct ExpTab<float> {
    static constexpr int kSmemElems = 0;
    __device__ static unsigned load(float*, int, int) { return 0u; }
};

// 2^(-u) for u >= 0 (u huge or +inf gives ~2^-1000 ~ 0).  `tab_s` is the shared-window address of
// the table (ExpTab<double>::load).
//
// magic = 1.5 * 2^41 has ulp 2^-11, so the low mantissa word of t = u + magic is j = round(u * 2048)
// (u < 2^21; beyond that the high word changes and the result is clamped to ~0).  u = j/2048 + f with
// |f| <= 2^-12 exactly, hence 2^(-u) = 2^(-(j >> 11)) * T[j & 2047] * 2^(-f): a table value, an
// exponent adjustment and a cubic in f (Taylor remainder (2^-12 ln2)^4 / 24 = 3.4e-17).
// fp64 pipe cost: 3 DADD + 3 DFMA + 1 DMUL (libm exp: 17 slots).
template <bool SAFE = false>
__device__ __forceinline__ double exp2neg(double u, unsigned tab_s) {
    constexpr double kMagic = 3298534883328.0;  // 1.5 * 2^41
    constexpr int kMagicHi = 0x42880000;        // its high word
    const double t = u + kMagic;
    unsigned lo = (unsigned)__double2loint(t);
    const int hi = __double2hiint(t);
    const double f = u - (t - kMagic);
    unsigned idx, addr;
    asm("and.b32 %0, %1, 2047;" : "=r"(idx) : "r"(lo));
    asm("mad.lo.u32 %0, %1, 8, %2;" : "=r"(addr) : "r"(idx), "r"(tab_s));
    double tv;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(addr));
    constexpr double c1 = -0.693147180559945309417232121458;
    constexpr double c2 = 0.240226506959100712333551263163;
    constexpr double c3 = -0.0555041086648215799531422637686;
    const double p = fma(f, fma(f, fma(f, c3, c2), c1), 1.0);
    // clamp the integer part at 1000 (also when u >= 2^21 spilled into the high word).  SAFE: the caller
    // has proved u < 1000 for every pair of the launch (bounding box of the scaled inputs), no clamp.
    if (!SAFE) {
        lo = (hi == kMagicHi) ? lo : 0xffffffffu;
        lo = min(lo, 1000u << kExpTabBits);
    }
    (void)hi;
    // table high words carry + (j << 9); subtracting lo << 9 = (n << 20) + (j << 9) leaves 2^-n * T[j]
    int thi;
    asm("mad.lo.s32 %0, %1, -512, %2;" : "=r"(thi) : "r"(lo), "r"(__double2hiint(tv)));
    return __hiloint2double(thi, __double2loint(tv)) * p;
}

// 2^(-u) for |u| < 2000 from the replicated 256-cell table (exptab16_load); no clamp: callers hold a launch-wide
// range proof.  magic = 1.5 * 2^44 has ulp 2^-8, so |f| <= 2^-9.  The cubic is the Chebyshev-node interpolant of
// 2^(-f) on [-2^-9, 2^-9] (near-minimax, max relative error 1.75e-14, equi-oscillating - against 1.4e-13 for the
// Taylor cubic); a quartic would reach 4e-17 for one more DFMA per evaluation, which costs 5 % of the kernel for
// accuracy the 1e-10 parity bar does not need.  fp64 pipe cost: 3 DADD + 3 DFMA + 1 DMUL.
template <bool I2F = false>
__device__ __forceinline__ double exp2neg16(double u, unsigned tab_lane) {
    constexpr double kMagic = 26388279066624.0;  // 1.5 * 2^44
    const double t = u + kMagic;
    const unsigned lo = (unsigned)__double2loint(t);
    // I2F: J = round(256 u) sits in the low mantissa word; its conversion runs on the conversion unit instead of the
    // fp64 pipe and u - J / 256 is exact in one FMA (bit-identical to the two dependent DADDs).  Measured at C3:
    // symmetric forward 499 -> 489 ms, backward unchanged (633 vs 636 ms), so only the forward enables it.
    const double f = I2F ? fma(__int2double_rn((int)lo), -0.00390625, u) : u - (t - kMagic);
    unsigned idx, addr;
    asm("and.b32 %0, %1, 255;" : "=r"(idx) : "r"(lo));
    asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr) : "r"(idx), "r"(tab_lane));
    double tv;
    asm("ld.shared.f64 %0, [%1];" : "=d"(tv) : "r"(addr));
    constexpr double c0 = 0.999999999999982504724;
    constexpr double c1 = -0.693147180559942884057;
    constexpr double c2 = 0.2402265436493534809523;
    constexpr double c3 = -0.05550411375117055438833;
    const double p = fma(f, fma(f, fma(f, c3, c2), c1), c0);
    // table high words carry + (j << 12); subtracting lo << 12 = (n << 20) + (j << 12) leaves 2^-n * T[j]
    int thi;
    asm("mad.lo.s32 %0, %1, -4096, %2;" : "=r"(thi) : "r"(lo), "r"(__double2hiint(tv)));
    return __hiloint2double(thi, __double2loint(tv)) * p;
}
template <bool I2F = false>
__device__ __forceinline__ float exp2neg16(float u, unsigned) { return u; }  // never used (fp64 only)

// exp2neg16<true> for R independent arguments, written stage by stage (all R range reductions, then all R table loads, ...).
// ptxas largely keeps this order, which (a) gives R-way instruction-level parallelism inside one warp and (b) puts
// instructions that share a register operand - the cubic's c2, the table base - back to back, where the operand
// reuse cache serves them: on B200 a sub-partition reads only ~two 32-bit register operands per cycle, so an fp64
// instruction with three fresh 64-bit sources costs 3 cycles instead of 2 (profiles/r02_rf_probe.txt).
//
// CLAMP (u up to 2^23, i.e. the |x|^2 < 900 level of launch_expand_level): the integer part n of u is limited to 1000
// in the exponent adjustment - one 32-bit IMNMX on the integer pipe instead of a DMNMX on u (the fp64 pipe is the
// bound; measured at C3, l = 0.3: forward 525 -> see profiles/r02_sweep_clamp2.log).  For n > 1000 the table index
// no longer matches the adjustment, so the result is 2^-1000 times a factor in (1/4, 4): still < 1e-300, like the
// true value 2^(-u).
template <int R, bool CLAMP = false>
__device__ __forceinline__ void exp2neg16_vec(const double (&u)[R], unsigned tab_lane, double (&kv)[R]) {
    constexpr double kMagic = 26388279066624.0;  // 1.5 * 2^44: ulp 2^-8
    constexpr double c0 = 0.999999999999982504724;
    constexpr double c1 = -0.693147180559942884057;
    constexpr double c2 = 0.2402265436493534809523;
    constexpr double c3 = -0.05550411375117055438833;
    double t[R], f[R], tv[R], p[R];
    unsigned lo[R], addr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) t[r] = u[r] + kMagic;
#pragma unroll
    for (int r = 0; r < R; ++r) lo[r] = (unsigned)__double2loint(t[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = fma(__int2double_rn((int)lo[r]), -0.00390625, u[r]);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        unsigned idx;
        asm("and.b32 %0, %1, 255;" : "=r"(idx) : "r"(lo[r]));
        asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(addr[r]) : "r"(idx), "r"(tab_lane));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) asm("ld.shared.f64 %0, [%1];" : "=d"(tv[r]) : "r"(addr[r]));
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = fma(f[r], c3, c2);
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = fma(f[r], p[r], c1);
#pragma unroll
    for (int r = 0; r < R; ++r) p[r] = fma(f[r], p[r], c0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int thi;
        const unsigned lc = CLAMP ? (unsigned)min((int)lo[r], 1000 << 8) : lo[r];
        asm("mad.lo.s32 %0, %1, -4096, %2;" : "=r"(thi) : "r"(lc), "r"(__double2hiint(tv[r])));
        tv[r] = __hiloint2double(thi, __double2loint(tv[r]));
    }
#pragma unroll
    for (int r = 0; r < R; ++r) kv[r] = tv[r] * p[r];
}

template <int R, bool CLAMP = false>
__device__ __forceinline__ void exp2neg16_vec(const float (&u)[R], unsigned, float (&kv)[R]) {  // never used (fp64 only)
#pragma unroll
    for (int r = 0; r < R; ++r) kv[r] = u[r];
}


template <bool SAFE = false>
__device__ __forceinline__ float exp2neg(float u, unsigned) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-u));
    return r;
}

// sqrt(x) for x >= 0 on the fp64 pipe without libm: MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22 relative) and two
// Goldschmidt steps on (g, h) = (x y, y / 2) -> (sqrt x, 1 / (2 sqrt x)): 1 DMNMX + 2 DMUL + 5 DFMA = 8 fp64 slots and
// no slow path, against ~17 slots + a branch for libm's correctly rounded sqrt.  Max error 2.1 ulp (checked for seeds
// up to 2^-20 off, tools/sqrt_check.py); the clamp at 1e-300 makes x = 0 return 1e-150 (K'(x, x) = 1 to 1e-150)
// and folds GPJax's max(sq, tiny) guard into the same instruction.
__device__ __forceinline__ double fast_sqrt(double x) {
    x = fmax(x, 1e-300);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    return fma(g, e, g);
}
__device__ __forceinline__ float fast_sqrt(float x) { return sqrtf(x); }

template <typename T>
struct Consts;
template <>
struct Consts<double> {
    static constexpr double ln2 = 0.693147180559945309417232121458;
    static constexpr double tiny = 1e-300;  // scaled analogue of GPJax's max(sq, 1e-36) guard
};
template <>
struct Consts<float> {
    static constexpr float ln2 = 0.693147180559945309417232121458f;
    static constexpr float tiny = 1e-36f;
};

// q = sum_t (a_t - b_t)^2.  For D >= 6 two independent accumulation chains halve the dependent-FMA depth
// (the Matern d = 8 kernels were latency-bound on the single chain).
template <typename T, int D>
__device__ __forceinline__ T sqdist(const T (&a)[D], const T (&b)[D]) {
    if (D >= 6) {
        T q0 = T(0), q1 = T(0);
#pragma unroll
        for (int t = 0; t < D; t += 2) {
            const T d0 = a[t] - b[t];
            q0 = fma(d0, d0, q0);
            if (t + 1 < D) {
                const T d1 = a[t + 1] - b[t + 1];
                q1 = fma(d1, d1, q1);
            }
        }
        return q0 + q1;
    }
    T q = T(0);
#pragma unroll
    for (int t = 0; t < D; ++t) {
        const T df = a[t] - b[t];
        q = fma(df, df, q);
    }
    return q;
}

// K' from q.
template <typename T, int FAM, bool SAFE = false>
__device__ __forceinline__ T kernel_value(T q, unsigned tab) {
    if (FAM == RBF) {
        return exp2neg<SAFE>(q, tab);
    } else {
        const T tp = fast_sqrt(q);
        const T e = exp2neg<SAFE>(tp, tab);
        if (FAM == MATERN12) return e;
        const T r = tp * Consts<T>::ln2;
        if (FAM == MATERN32) return fma(r, e, e);
        // MATERN52: (1 + r + r^2/3) e
        return fma(fma(r, T(1.0 / 3.0), T(1)), r, T(1)) * e;
    }
}

// K' and H (see header comment) from q.
template <typename T, int FAM, bool SAFE = false>
__device__ __forceinline__ void kernel_value_grad(T q, unsigned tab, T& kv, T& h) {
    if (FAM == RBF) {
        kv = exp2neg<SAFE>(q, tab);
        h = kv;
    } else {
        const T tp = fast_sqrt(q);
        const T e = exp2neg<SAFE>(tp, tab);
        const T r = tp * Consts<T>::ln2;
        if (FAM == MATERN12) {
            kv = e;
            h = (q > Consts<T>::tiny) ? e / tp : T(0);
        } else if (FAM == MATERN32) {
            kv = fma(r, e, e);
            h = e;
        } else {
            kv = fma(fma(r, T(1.0 / 3.0), T(1)), r, T(1)) * e;
            h = fma(r, e, e) * T(1.0 / 3.0);
        }
    }
}

// Launch-wide range proof: bbox = {min_0..min_{D-1}, max_0..max_{D-1}} of ALL scaled points of the launch (or
// null).  q <= sum_t (max_t - min_t)^2; the exp2 argument is q (RBF) or sqrt(q) (Matern) and must stay < 1000.
template <typename T, int D, int FAM>
__device__ __forceinline__ bool launch_is_safe(const T* __restrict__ bbox) {
    if (sizeof(T) == 4 || bbox == nullptr) return false;
    T qm = T(0);
    for (int t = 0; t < D; ++t) {
        const T w = bbox[D + t] - bbox[t];
        qm = fma(w, w, qm);
    }
    return FAM == RBF ? (qm < T(990)) : (qm < T(990) * T(990));  // false for NaN
}

// Launch-wide proof for the EXPANDED distance form of the RBF kernels (fp64): with na = |a|^2, nb = |b|^2,
//   K' = 2^(-|a - b|^2) = 2^(-na) * 2^(-(nb - 2 a.b)),
// so the per-pair work is a D-term FMA chain started at nb (D fp64 slots instead of 2 D) and the row factor
// 2^(-na) is applied once per row.  Valid when every point of the launch has |x|^2 < 300: the exponent
// nb - 2 a.b then lies in (-300, 900) (no clamp, no overflow) and its rounding error, a few ulp of 900, is
// below 1e-12 relative in K' (the coordinates are origin-shifted by cagp_scale_inputs).
template <typename T, int D, int FAM>
__device__ __forceinline__ bool launch_is_expandable(const T* __restrict__ bbox) {
    if (sizeof(T) == 4 || bbox == nullptr || FAM != RBF) return false;
    T s = T(0);
    for (int t = 0; t < D; ++t) {
        const T m = fmax(fabs(bbox[t]), fabs(bbox[D + t]));
        s = fma(m, m, s);
    }
    return s < T(300);  // false for NaN
}

// Level of the expanded RBF form for the fast-path kernels (ks_symx.cuh): 1 = every |x|^2 < 300 (as above), 2 = every
// |x|^2 < 900: u = |b|^2 - 2 a.b can then exceed the ~1000 the exponent adjustment of exp2neg16 can represent, so u is
// clamped at 1000 first (one DMNMX more per pair; the true K' = 2^(-|a|^2 - u) is < 2^-1000 there anyway), and
// 2^(-u) <= 2^900 keeps the row sums finite.  Rounding error of u: a few ulp of 2700, < 1e-12 relative in K'.
// 0 = not expandable (the general kernel runs).  With the box centred, level 2 covers lengthscales down to
// ~0.25 at C3's domain (level 1: ~0.43).
template <typename T, int D, int FAM>
__device__ __forceinline__ int launch_expand_level(const T* __restrict__ bbox) {
    if (sizeof(T) == 4 || bbox == nullptr || FAM != RBF) return 0;
    T s = T(0);
    for (int t = 0; t < D; ++t) {
        const T m = fmax(fabs(bbox[t]), fabs(bbox[D + t]));
        s = fma(m, m, s);
    }
    return s < T(300) ? 1 : (s < T(900) ? 2 : 0);  // 0 for NaN
}

// Launch-wide proof for the EXPANDED DISTANCE of the smooth Matern families (fp64, Matern-3/2 and -5/2):
//   q = |a - b|^2 = |a|^2 + (|b|^2 - 2 a.b),
// a D-term FMA chain started at |b|^2 plus one add, instead of D subtractions + D FMAs.  The cancellation costs an
// absolute error of a few ulp of |a|^2 + |b|^2 in q; K'(q) of these families is smooth at q = 0 (dK'/dq bounded by
// ~0.25 in scaled units), so with every |x|^2 < 2000 the kernel value is off by < 5e-13 absolute (K' <= 1), far
// inside the 1e-10 parity bar.  Matern-1/2 (K' = 2^(-sqrt q), infinite slope at 0) keeps the exact differences.
template <typename T, int D, int FAM>
__device__ __forceinline__ bool launch_is_xdist(const T* __restrict__ bbox) {
    if (sizeof(T) == 4 || bbox == nullptr || !(FAM == MATERN32 || FAM == MATERN52)) return false;
    T s = T(0);
    for (int t = 0; t < D; ++t) {
        const T m = fmax(fabs(bbox[t]), fabs(bbox[D + t]));
        s = fma(m, m, s);
    }
    return s < T(2000);  // false for NaN
}

// host+device: coordinate scale and gradient constant per family
__host__ __device__ inline double family_scale(int fam) {
    const double log2e = 1.44269504088896340735992468100189214;
    switch (fam) {
        case RBF: return 0.84932180028801904272150283410288961;  // sqrt(0.5 * log2 e)
        case MATERN12: return log2e;
        case MATERN32: return 1.73205080756887729352744634150587237 * log2e;
        default: return 2.23606797749978969640917366873127624 * log2e;
    }
}
__host__ __device__ inline double family_grad_const(int fam) {
    const double ln2 = 0.693147180559945309417232121458;
    switch (fam) {
        case RBF: return 2.0 * ln2;
        case MATERN12: return ln2;
        default: return ln2 * ln2;
    }
}

}  // namespace cagp
