// Materialised cross-covariance C = K'(X1, X2) (m, n) and its reverse mode w.r.t. the X2 points and the
// lengthscale(s).  This is the action producer of PseudoInputPolicy (reference: policies/pseudoinput.py:68-74,
// S = kernel.cross_covariance(train_inputs, pseudo_inputs)): m = number of training points (large), n = number of
// pseudo-inputs (small), and the pseudo-inputs are trainable, hence the X2 cotangent.
//
// HBM-bound (one (m, n) matrix written / read once; the x1 coordinates are read once per 128 columns), so the
// layout is what matters: one thread per COLUMN j keeps x2_j in registers and walks the rows of its chunk, which
// makes every load / store of C and Cbar a fully coalesced row segment; the x1 coordinates of the current row
// tile are staged in shared memory and read as broadcasts.
#pragma once
#include "kernel_math.cuh"

namespace cagp {

constexpr int kCrossThreads = 128;  // columns per CTA
constexpr int kCrossRowTile = 64;   // x1 rows staged per step

template <typename T>
struct CrossArgs {
    const T* xs1; int64_t ld1, m;
    const T* xs2; int64_t ld2, n;
    T* C;               // fwd: output (m, ldc); bwd: cotangent Cbar (read)
    int64_t ldc;
    int64_t rows_per_cta;  // multiple of kCrossRowTile
    // backward only
    T* x2_part;         // [gridDim.y][D][ldb] partial cotangents of the SCALED x2 coordinates
    int64_t ldb;
    T* ls_part;         // [gridDim.y * gridDim.x][NLS]
    T gc;               // family_grad_const
};

template <typename T, int D>
__device__ __forceinline__ void cross_stage_rows(const CrossArgs<T>& a, T* rows, int64_t base, int tid) {
    for (int e = tid; e < D * kCrossRowTile; e += kCrossThreads) {
        const int t = e / kCrossRowTile, r = e % kCrossRowTile;
        const int64_t i = base + r;
        rows[r * D + t] = i < a.m ? a.xs1[(int64_t)t * a.ld1 + i] : T(0);
    }
}

template <typename T, int D, int FAM>
__global__ void __launch_bounds__(kCrossThreads) cross_cov_fwd_kernel(CrossArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tabp = reinterpret_cast<T*>(smem_raw);
    T* rows = tabp + ExpTab<T>::kSmemElems;
    const int tid = threadIdx.x;
    const unsigned tab = ExpTab<T>::load(tabp, tid, kCrossThreads);
    const int64_t j = (int64_t)blockIdx.x * kCrossThreads + tid;
    T b[D];
#pragma unroll
    for (int t = 0; t < D; ++t) b[t] = j < a.n ? a.xs2[(int64_t)t * a.ld2 + j] : T(0);
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_cta;
    const int64_t r1 = min(r0 + a.rows_per_cta, a.m);
    for (int64_t base = r0; base < r1; base += kCrossRowTile) {
        __syncthreads();
        cross_stage_rows<T, D>(a, rows, base, tid);
        __syncthreads();
        const int nr = (int)min((int64_t)kCrossRowTile, r1 - base);
        if (j < a.n) {
#pragma unroll 4
            for (int r = 0; r < nr; ++r) {
                T x[D];
#pragma unroll
                for (int t = 0; t < D; ++t) x[t] = rows[r * D + t];
                const T q = sqdist<T, D>(x, b);
                a.C[(base + r) * a.ldc + j] = kernel_value<T, FAM>(q, tab);
            }
        }
    }
}

// x2_part[y][t][j] = gc * sum_{i in chunk y} Cbar_ij H_ij (x1_it - x2_jt);  ls_part[cta][t] = sum Cbar_ij H_ij q_t
template <typename T, int D, int FAM, bool ARD>
__global__ void __launch_bounds__(kCrossThreads) cross_cov_bwd_kernel(CrossArgs<T> a) {
    constexpr int NLS = ARD ? D : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tabp = reinterpret_cast<T*>(smem_raw);
    T* rows = tabp + ExpTab<T>::kSmemElems;
    __shared__ T red[kCrossThreads / 32][NLS];
    const int tid = threadIdx.x;
    const unsigned tab = ExpTab<T>::load(tabp, tid, kCrossThreads);
    const int64_t j = (int64_t)blockIdx.x * kCrossThreads + tid;
    T b[D], bb[D], lsacc[NLS];
#pragma unroll
    for (int t = 0; t < D; ++t) {
        b[t] = j < a.n ? a.xs2[(int64_t)t * a.ld2 + j] : T(0);
        bb[t] = T(0);
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) lsacc[t] = T(0);
    const int64_t r0 = (int64_t)blockIdx.y * a.rows_per_cta;
    const int64_t r1 = min(r0 + a.rows_per_cta, a.m);
    for (int64_t base = r0; base < r1; base += kCrossRowTile) {
        __syncthreads();
        cross_stage_rows<T, D>(a, rows, base, tid);
        __syncthreads();
        const int nr = (int)min((int64_t)kCrossRowTile, r1 - base);
        if (j < a.n) {
#pragma unroll(D >= 16 ? 1 : 2)
            for (int r = 0; r < nr; ++r) {
                T q = T(0);
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T df = rows[r * D + t] - b[t];
                    q = fma(df, df, q);
                }
                T kv, h;
                kernel_value_grad<T, FAM>(q, tab, kv, h);
                const T g = a.C[(base + r) * a.ldc + j] * h;
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T df = rows[r * D + t] - b[t];  // re-read (broadcast) instead of keeping D differences live
                    const T gd = g * df;
                    bb[t] += gd;
                    if (ARD) lsacc[t] = fma(gd, df, lsacc[t]);
                }
                if (!ARD) lsacc[0] = fma(g, q, lsacc[0]);
            }
        }
    }
    if (j < a.n) {
#pragma unroll
        for (int t = 0; t < D; ++t) a.x2_part[((int64_t)blockIdx.y * D + t) * a.ldb + j] = a.gc * bb[t];
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) {
        T v = lsacc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][t] = v;
    }
    __syncthreads();
    if (tid < NLS) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < kCrossThreads / 32; ++w) v += red[w][tid];
        a.ls_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NLS + tid] = v;
    }
}

}  // namespace cagp
