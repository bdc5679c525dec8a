// Block-sparse projection kernels: M' = K'(X1, X2) S and its reverse mode.  sm_100a.
//
// Both passes share one skeleton: a CTA owns NT*R "row" points (R per thread, coordinates in
// registers) and streams "column" points through shared memory in chunks of JC; each smem entry
// is {x_0..x_{D-1}, weight} so the inner loop reads it as broadcast LDS.128s.  Pair work is pure
// fp64 (or fp32) pipe arithmetic; HBM traffic is negligible (X is L2 resident), so the bound is
// the FMA pipe, not memory.
//
//   forward : rows = x1 points (any), columns = x2 points grouped by action block b, weight s_j,
//             output Mt[b][i] = sum_{j in b} K'_ij s_j.
//   backward: rows = x2 points of ONE action block b (so they share the weight row Gt[b][:]),
//             columns = x1 points, weight Gt[b][i]; outputs s_bar_j = sum_i K'_ij Gt[b][i] and
//             the lengthscale partial sum_i Gt[b][i] H_ij q_ij(t).  Column sums of the forward
//             formulation become row sums here, so no cross-thread reduction is needed.
#pragma once
#include <type_traits>

#include "kernel_math.cuh"

namespace cagp {

template <int D>
struct EntrySize {
    static constexpr int value = (D + 1 + 1) & ~1;  // D coords + weight, padded to even
};

template <typename T, int E>
__device__ __forceinline__ void load_entry(const T* __restrict__ p, T (&v)[E]);

template <int E>
__device__ __forceinline__ void load_entry_d(const double* __restrict__ p, double (&v)[E]) {
    const double2* p2 = reinterpret_cast<const double2*>(p);
#pragma unroll
    for (int i = 0; i < E / 2; ++i) {
        const double2 t = p2[i];
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
    }
}
template <int E>
__device__ __forceinline__ void load_entry_f(const float* __restrict__ p, float (&v)[E]) {
    const float2* p2 = reinterpret_cast<const float2*>(p);
#pragma unroll
    for (int i = 0; i < E / 2; ++i) {
        const float2 t = p2[i];
        v[2 * i] = t.x;
        v[2 * i + 1] = t.y;
    }
}
template <int E>
__device__ __forceinline__ void load_entry_any(const double* p, double (&v)[E]) { load_entry_d<E>(p, v); }
template <int E>
__device__ __forceinline__ void load_entry_any(const float* p, float (&v)[E]) { load_entry_f<E>(p, v); }

struct BlockLayout {
    int64_t n;   // total columns
    int k;       // number of action blocks
    int64_t bs;  // n / k
    __host__ __device__ int64_t start(int b) const { return (int64_t)b * bs; }
    __host__ __device__ int64_t len(int b) const { return (b == k - 1) ? (n - (int64_t)(k - 1) * bs) : bs; }
    __host__ __device__ int block_of(int64_t j) const {
        if (bs == 0) return k - 1;
        const int64_t b = j / bs;
        return (int)(b < k - 1 ? b : k - 1);
    }
};

// The factored RBF form is compiled into the fp64 RBF instances only (32 KB exp2 table instead of 16 KB).
template <typename T, int FAM>
struct FwdCanExpand {
    static constexpr bool value = sizeof(T) == 8 && FAM == RBF;
    static constexpr int kTabElems = value ? kExpTab16Elems : ExpTab<T>::kSmemElems;
};

template <typename T>
struct FwdArgs {
    const T* xs1; int64_t ld1; int64_t m;
    const T* xs2; int64_t ld2;
    const T* s;
    BlockLayout lay;
    T* Mt; int64_t ldm;
    int blocks_per_cta;
    const T* bbox;  // bounding box of ALL scaled points of the call (2 * D values) or null
};

template <typename T, int D, int FAM, int R, int NT, int JC>
__global__ void __launch_bounds__(NT) ks_fwd_kernel(const FwdArgs<T> p) {
    constexpr int E = EntrySize<D>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr bool CAN_EXPAND = FwdCanExpand<T, FAM>::value;
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* cols = tab + FwdCanExpand<T, FAM>::kTabElems;
    const int tid = threadIdx.x;
    const bool safe = launch_is_safe<T, D, FAM>(p.bbox);  // uniform: no clamp needed in exp2neg
    // uniform: fully factored RBF form K' = 2^(-|a|^2) 2^(-|b|^2) 2^(2 a.b) (kernel_math.cuh: launch_is_expandable).
    // The column factor is folded into the staged weight, the row factor is applied to the finished row sum.
    const bool expand = CAN_EXPAND && launch_is_expandable<T, D, FAM>(p.bbox);
    unsigned tab_s;
    if constexpr (CAN_EXPAND) {
        tab_s = expand ? exptab16_load(tab, tid, NT) : ExpTab<T>::load(tab, tid, NT);
    } else {
        tab_s = ExpTab<T>::load(tab, tid, NT);
    }
    __syncthreads();  // the table is used by the row set-up and the column staging below

    const int64_t row0 = (int64_t)blockIdx.x * (NT * R);
    T a[R][D], erow[R];
    int64_t irow[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        irow[r] = row0 + r * NT + tid;
        const int64_t ic = irow[r] < p.m ? irow[r] : p.m - 1;
#pragma unroll
        for (int t = 0; t < D; ++t) a[r][t] = p.xs1[t * p.ld1 + ic];
        erow[r] = T(1);
        if constexpr (CAN_EXPAND) {
            if (expand) {
                T na = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) na = fma(a[r][t], a[r][t], na);
                erow[r] = exp2neg16(na, tab_s);
            }
        }
    }
    const int b_begin = blockIdx.y * p.blocks_per_cta;
    const int b_end = min(p.lay.k, b_begin + p.blocks_per_cta);
    for (int b = b_begin; b < b_end; ++b) {
        const int64_t j0 = p.lay.start(b);
        const int64_t len = p.lay.len(b);
        T acc[R];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = T(0);
        for (int64_t c0 = 0; c0 < len; c0 += JC) {
            const int cl = (int)min((int64_t)JC, len - c0);
            __syncthreads();
            for (int idx = tid; idx < cl; idx += NT) {
                const int64_t j = j0 + c0 + idx;
                T nb = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T v = p.xs2[t * p.ld2 + j];
                    nb = fma(v, v, nb);
                    cols[idx * E + t] = expand ? T(-2) * v : v;
                }
                T w = p.s[j];
                if constexpr (CAN_EXPAND) {
                    if (expand) w *= exp2neg16(nb, tab_s);
                }
                cols[idx * E + D] = w;
            }
            __syncthreads();
            auto sweep = [&](auto mode_c) {
                constexpr int MODE = decltype(mode_c)::value;  // 0 general, 1 safe range, 2 safe + factored form
                constexpr bool SAFE = MODE >= 1;
#pragma unroll 2
                for (int j = 0; j < cl; ++j) {
                    T e[E];
                    load_entry_any<E>(cols + j * E, e);
                    if constexpr (MODE == 2) {
                        // stage-ordered over the R rows (kernel_math.cuh: exp2neg16_vec): instructions that share a
                        // register operand - the column coordinate, the cubic's c2, the weight - sit back to back
                        // and are served by the operand-reuse cache (profiles/r02_pair_kernels.md)
                        T u[R], kv[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) u[r] = a[r][0] * e[0];  // -2 a.b
#pragma unroll
                        for (int t = 1; t < D; ++t)
#pragma unroll
                            for (int r = 0; r < R; ++r) u[r] = fma(a[r][t], e[t], u[r]);
                        exp2neg16_vec<R>(u, tab_s, kv);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = fma(kv[r], e[D], acc[r]);
                        continue;
                    }
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if constexpr (MODE == 2) {
                        } else {
                            T q = T(0);
#pragma unroll
                            for (int t = 0; t < D; ++t) {
                                const T df = a[r][t] - e[t];
                                q = fma(df, df, q);
                            }
                            const T kv = kernel_value<T, FAM, SAFE>(q, tab_s);
                            acc[r] = fma(kv, e[D], acc[r]);
                        }
                    }
                }
            };
            if constexpr (sizeof(T) == 8) {
                if constexpr (CAN_EXPAND) {
                    if (expand) sweep(std::integral_constant<int, 2>{});
                    else if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                } else {
                    if (safe) sweep(std::integral_constant<int, 1>{});
                    else sweep(std::integral_constant<int, 0>{});
                }
            } else {
                sweep(std::integral_constant<int, 0>{});
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (irow[r] < p.m) p.Mt[(int64_t)b * p.ldm + irow[r]] = acc[r] * erow[r];
    }
}

template <typename T>
struct BwdArgs {
    const T* xs1; int64_t ld1; int64_t m;   // streamed points (carry the weight Gt[b][i])
    const T* xs2; int64_t ld2;              // row points, grouped by action block
    const T* s;
    BlockLayout lay;
    const T* Gt; int64_t ldg;
    T* sbar_part;   // [n_ichunks][n]
    T* ls_part;     // [gridDim.y * gridDim.x][NLS]
    int tiles_main; // tiles per main block
    int64_t ichunk; // streamed points per CTA
    const T* bbox;  // bounding box of ALL scaled points of the call (2 * D values) or null
};

// ARD == false: one accumulator (isotropic lengthscale); ARD == true: D accumulators.
template <typename T, int D, int FAM, bool ARD, int R, int NT, int JC>
__global__ void __launch_bounds__(NT) ks_bwd_kernel(const BwdArgs<T> p) {
    constexpr int E = EntrySize<D>::value;
    constexpr int NLS = ARD ? D : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* cols = tab + ExpTab<T>::kSmemElems;
    __shared__ T red[NT / 32][NLS];
    const int tid = threadIdx.x;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, NT);

    // which action block / which slice of it
    int b;
    int64_t off;
    {
        const int64_t nmain = (int64_t)(p.lay.k - 1) * p.tiles_main;
        if ((int64_t)blockIdx.x < nmain) {
            b = blockIdx.x / p.tiles_main;
            off = (int64_t)(blockIdx.x % p.tiles_main) * (NT * R);
        } else {
            b = p.lay.k - 1;
            off = ((int64_t)blockIdx.x - nmain) * (NT * R);
        }
    }
    const int64_t j0 = p.lay.start(b);
    const int64_t blen = p.lay.len(b);
    T a[R][D];
    int64_t jrow[R];
    bool valid[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t o = off + r * NT + tid;
        valid[r] = o < blen;
        jrow[r] = j0 + (valid[r] ? o : (blen > 0 ? blen - 1 : 0));
        if (blen == 0) jrow[r] = 0;
#pragma unroll
        for (int t = 0; t < D; ++t) a[r][t] = p.xs2[t * p.ld2 + jrow[r]];
    }
    T sacc[R];
    T lacc[R][NLS];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        sacc[r] = T(0);
#pragma unroll
        for (int t = 0; t < NLS; ++t) lacc[r][t] = T(0);
    }
    const int64_t i_begin = (int64_t)blockIdx.y * p.ichunk;
    const int64_t i_end = min(p.m, i_begin + p.ichunk);
    const T* __restrict__ grow = p.Gt + (int64_t)b * p.ldg;
    for (int64_t c0 = i_begin; c0 < i_end; c0 += JC) {
        const int cl = (int)min((int64_t)JC, i_end - c0);
        __syncthreads();
        for (int idx = tid; idx < cl; idx += NT) {
            const int64_t i = c0 + idx;
#pragma unroll
            for (int t = 0; t < D; ++t) cols[idx * E + t] = p.xs1[t * p.ld1 + i];
            cols[idx * E + D] = grow[i];
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < cl; ++j) {
            T e[E];
            load_entry_any<E>(cols + j * E, e);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                T q = T(0);
                T qt[D];
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T df = a[r][t] - e[t];
                    if (ARD) {
                        qt[t] = df * df;
                        q += qt[t];
                    } else {
                        q = fma(df, df, q);
                    }
                }
                T kv, h;
                kernel_value_grad<T, FAM>(q, tab_s, kv, h);
                const T g = e[D];
                if (FAM == RBF) {
                    const T kg = kv * g;
                    sacc[r] += kg;
                    if (ARD) {
#pragma unroll
                        for (int t = 0; t < D; ++t) lacc[r][t] = fma(kg, qt[t], lacc[r][t]);
                    } else {
                        lacc[r][0] = fma(kg, q, lacc[r][0]);
                    }
                } else {
                    sacc[r] = fma(kv, g, sacc[r]);
                    const T hg = h * g;
                    if (ARD) {
#pragma unroll
                        for (int t = 0; t < D; ++t) lacc[r][t] = fma(hg, qt[t], lacc[r][t]);
                    } else {
                        lacc[r][0] = fma(hg, q, lacc[r][0]);
                    }
                }
            }
        }
    }
    // outputs: s_bar partial for this streamed chunk, lengthscale partial weighted by s_j
    T lsum[NLS];
#pragma unroll
    for (int t = 0; t < NLS; ++t) lsum[t] = T(0);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (valid[r]) {
            p.sbar_part[(int64_t)blockIdx.y * p.lay.n + jrow[r]] = sacc[r];
            const T sj = p.s[jrow[r]];
#pragma unroll
            for (int t = 0; t < NLS; ++t) lsum[t] = fma(sj, lacc[r][t], lsum[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) {
        T v = lsum[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][t] = v;
    }
    __syncthreads();
    if (tid < NLS) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < NT / 32; ++w) v += red[w][tid];
        p.ls_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NLS + tid] = v;
    }
}

}  // namespace cagp
