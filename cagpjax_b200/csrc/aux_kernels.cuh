// HBM-bound helper kernels around the pair kernels: input scaling, row projections
// (A' = S^T M', c' = M'^T yt), cotangent assembly, reductions.  All deterministic.
#pragma once
#include "ks_blocksparse.cuh"

namespace cagp {

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// xs[t*ld + i] = (X[i*d + t] - origin[t]) * c / lengthscale_t for t < d, 0 for d <= t < gridDim.y (padding)
template <typename T>
__global__ void scale_inputs_kernel(const T* __restrict__ X, int64_t n, int d, const T* __restrict__ ls, int n_ls, T c,
                                    const T* __restrict__ origin, T* __restrict__ xs, int64_t ld) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (i >= n) return;
    T v = T(0);
    if (t < d) {
        const T l = ls[n_ls > 1 ? t : 0];
        const T o = origin ? origin[t] : T(0);
        v = (X[i * d + t] - o) * (c / l);
    }
    xs[(int64_t)t * ld + i] = v;
}

// bbox[t] = min_i xs[t][i], bbox[dp + t] = max_i xs[t][i]; one CTA per coordinate row (deterministic)
template <typename T>
__global__ void __launch_bounds__(1024) bounding_box_kernel(const T* __restrict__ xs, int64_t n, int64_t ld, int dp,
                                                            T* __restrict__ bbox) {
    __shared__ T smin[32], smax[32];
    const int t = blockIdx.x;
    const T* __restrict__ row = xs + (int64_t)t * ld;
    T lo = row[0], hi = row[0];
    bool bad = false;  // comparisons skip a NaN, so it is tracked separately
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const T v = row[i];
        bad |= v != v;
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    bad = __syncthreads_or(bad);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const T a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            lo = smin[w] < lo ? smin[w] : lo;
            hi = smax[w] > hi ? smax[w] : hi;
        }
        // a NaN coordinate poisons the box: every launch_* predicate of kernel_math.cuh compares with '<', which is
        // false for NaN, so the clamped general kernels run (and propagate the NaN as the reference does)
        const T qnan = T(nan(""));
        bbox[t] = bad ? qnan : lo;
        bbox[dp + t] = bad ? qnan : hi;
    }
}

// One CTA per action column b.  For every action block a intersecting the caller's rows
// [row_offset, row_offset + m): A[a][b] = sum_i s_i Mt[b][i]; other rows of column b are zeroed.
// c[b] = sum_i Mt[b][i] yt_i over all local rows.  s, yt are indexed by LOCAL row.
template <typename T>
__global__ void __launch_bounds__(256) project_rows_kernel(const T* __restrict__ Mt, int64_t ldm, int64_t m,
                                                           int64_t row_offset, BlockLayout lay,
                                                           const T* __restrict__ s, const T* __restrict__ yt,
                                                           T* __restrict__ A, T* __restrict__ c) {
    __shared__ T cred[8];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const T* __restrict__ row = Mt + (int64_t)b * ldm;
    const int a_first = lay.block_of(row_offset);
    const int a_last = lay.block_of(row_offset + m - 1);
    for (int a = threadIdx.x; a < lay.k; a += 256)
        if (a < a_first || a > a_last) A[(int64_t)a * lay.k + b] = T(0);
    T cacc = T(0);
    for (int a = a_first + warp; a <= a_last; a += 8) {
        int64_t lo = lay.start(a) - row_offset, hi = lo + lay.len(a);
        if (lo < 0) lo = 0;
        if (hi > m) hi = m;
        T aacc = T(0);
        for (int64_t i = lo + lane; i < hi; i += 32) {
            const T mv = row[i];
            aacc = fma(s[i], mv, aacc);
            cacc = fma(yt[i], mv, cacc);
        }
        aacc = warp_sum(aacc);
        if (lane == 0) A[(int64_t)a * lay.k + b] = aacc;
    }
    cacc = warp_sum(cacc);
    if (lane == 0) cred[warp] = cacc;
    __syncthreads();
    if (threadIdx.x == 0) {
        T v = T(0);
#pragma unroll
        for (int w = 0; w < 8; ++w) v += cred[w];
        c[b] = v;
    }
}

// B is column-major lower == row-major upper after syrk; fill the other triangle.
template <typename T>
__global__ void mirror_lower_kernel(T* __restrict__ B, int k) {
    const int r = blockIdx.y * blockDim.y + threadIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < k && c < k && r > c) B[(int64_t)r * k + c] = B[(int64_t)c * k + r];
}

// Gt[b][i] = Abar[blk(i)][b] * s_i + cbar[b] * yt_i   (the GEMM term is accumulated afterwards)
template <typename T>
__global__ void cotangent_prefill_kernel(int64_t m, int64_t row_offset, BlockLayout lay, const T* __restrict__ Abar,
                                         const T* __restrict__ cbar, const T* __restrict__ s,
                                         const T* __restrict__ yt, T* __restrict__ Gt, int64_t ldg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= m) return;
    const int a = lay.block_of(row_offset + i);
    Gt[(int64_t)b * ldg + i] = fma(Abar[(int64_t)a * lay.k + b], s[i], cbar[b] * yt[i]);
}

// s_dir[i] = sum_b Abar[blk(i)][b] Mt[b][i];  yt_bar[i] = sum_b cbar[b] Mt[b][i]
template <typename T>
__global__ void project_rows_bwd_kernel(const T* __restrict__ Mt, int64_t ldm, int64_t m, int64_t row_offset,
                                        BlockLayout lay, const T* __restrict__ Abar, const T* __restrict__ cbar,
                                        T* __restrict__ s_dir, T* __restrict__ yt_bar) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int a = lay.block_of(row_offset + i);
    const T* __restrict__ arow = Abar + (int64_t)a * lay.k;
    T sd = T(0), yb = T(0);
#pragma unroll 4
    for (int b = 0; b < lay.k; ++b) {
        const T mv = Mt[(int64_t)b * ldm + i];
        sd = fma(arow[b], mv, sd);
        yb = fma(cbar[b], mv, yb);
    }
    s_dir[i] = sd;
    yt_bar[i] = yb;
}

// mean[i] = mu + var * sum_b Mt[b][i] w[b];  v[i] = var - var^2 * sum_b Mt[b][i] Z[b][i]
// scalars = {variance, mean_const}
template <typename T>
__global__ void predict_moments_kernel(const T* __restrict__ Mt, int64_t ldm, const T* __restrict__ Z, int64_t ldz,
                                       int64_t m, int k, const T* __restrict__ w, const T* __restrict__ scalars,
                                       T* __restrict__ mean, T* __restrict__ var) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    T ma = T(0), va = T(0);
#pragma unroll 4
    for (int b = 0; b < k; ++b) {
        const T mv = Mt[(int64_t)b * ldm + i];
        ma = fma(mv, w[b], ma);
        va = fma(mv, Z[(int64_t)b * ldz + i], va);
    }
    const T variance = scalars[0], mu = scalars[1];
    mean[i] = fma(variance, ma, mu);
    var[i] = variance - variance * variance * va;
}

// s_bar[j] = sum_c sbar_part[c][j];  ls_bar[t] = (gc / l_t) * sum_p ls_part[p][t]   (last CTA)
template <typename T>
__global__ void __launch_bounds__(256) bwd_finalize_kernel(const T* __restrict__ sbar_part, int n_chunks, int64_t n,
                                                           T* __restrict__ s_bar, const T* __restrict__ ls_part,
                                                           int64_t nparts, int nls, int n_ls_out,
                                                           const T* __restrict__ ls, T gc, T* __restrict__ ls_bar) {
    if (blockIdx.x + 1 < gridDim.x) {
        const int64_t j = (int64_t)blockIdx.x * 256 + threadIdx.x;
        if (j < n) {
            T v = T(0);
            for (int c = 0; c < n_chunks; ++c) v += sbar_part[(int64_t)c * n + j];
            s_bar[j] = v;
        }
        return;
    }
    __shared__ T red[8];
    for (int t = 0; t < n_ls_out; ++t) {
        T v = T(0);
        for (int64_t p = threadIdx.x; p < nparts; p += 256) v += ls_part[p * nls + t];
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) {
            T tot = T(0);
#pragma unroll
            for (int w = 0; w < 8; ++w) tot += red[w];
            ls_bar[t] = tot * gc / ls[t];
        }
    }
}

}  // namespace cagp
