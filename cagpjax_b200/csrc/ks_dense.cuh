// Dense right-hand side kernels: Y = K'(X1, X2) V for a dense V (n, p) and the lengthscale part of
// its reverse mode.  This is LazyKernel._matmat for dense operands (operators/lazy_kernel.py:79-90):
// matvecs (p = 1) and dense action matrices (BASELINE config 5).  sm_100a.
//
// Unlike the block-sparse case the contraction over j is a real GEMM (2 m n p flops), so a kernel
// value must be shared by all p output columns.  A CTA owns 64 rows and up to 128 output columns and
// walks the n columns of K' in chunks of JC = 32:
//   phase A: the 256 threads evaluate the 64 x 32 kernel tile once (8 values each, row coordinates in
//            registers, column coordinates broadcast from shared memory) and park it in shared memory;
//   phase B: register-tiled FMA GEMM tile (8 rows x 4 columns per thread): the K' values of a warp's 8
//            rows are warp-broadcast LDS.128s, the V values are conflict-free LDS.128s.
// p > 128 takes ceil(p / 128) passes (grid.z) that re-evaluate the kernel tile (45 of ~300 fp64 slots
// per pair at d = 16).  fp64 tensor-core MMA (DMMA) runs on the same datapath as DFMA on B200
// (profiles/r01_issue_model.txt), so the FMA formulation loses no peak throughput.
#pragma once
#include "ks_blocksparse.cuh"

namespace cagp {

constexpr int kDenseRows = 64;    // rows per CTA
constexpr int kDenseCols = 128;   // output columns per CTA pass
constexpr int kDenseJC = 32;      // kernel columns per stage
constexpr int kDenseThreads = 256;

template <typename T>
struct DenseFwdArgs {
    const T* xs1; int64_t ld1; int64_t m;
    const T* xs2; int64_t ld2; int64_t n;
    const T* V; int64_t ldv; int p;
    T* Y; int64_t ldy;
    int64_t jchunk;  // kernel columns per CTA (grid.y splits n; > 1 split => atomicAdd into zeroed Y)
};

template <typename T>
__device__ __forceinline__ void load4(const T* p, T (&v)[4]);
template <>
__device__ __forceinline__ void load4<double>(const double* p, double (&v)[4]) {
    const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&v)[4]) {
    const float4 a = reinterpret_cast<const float4*>(p)[0];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}

template <typename T, int D, int FAM>
__global__ void __launch_bounds__(kDenseThreads, 2) ks_dense_fwd_kernel(const DenseFwdArgs<T> p) {
    constexpr int JC = kDenseJC, TM = kDenseRows, TP = kDenseCols;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc = tab + ExpTab<T>::kSmemElems;  // [D][JC]
    T* Ks = xc + D * JC;                  // [JC][TM]
    T* Vs = Ks + JC * TM;                 // [JC][TP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, kDenseThreads);

    const int64_t row0 = (int64_t)blockIdx.x * TM;
    const int rowA = tid & (TM - 1);  // phase-A row of this thread
    const int jsub = tid >> 6;        // 0..3
    T ar[D];
    {
        const int64_t i = min(row0 + rowA, p.m - 1);
#pragma unroll
        for (int t = 0; t < D; ++t) ar[t] = p.xs1[t * p.ld1 + i];
    }
    T acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = T(0);
    const int colbase = blockIdx.z * TP;
    const int64_t j_begin = (int64_t)blockIdx.y * p.jchunk;
    const int64_t j_end = min(p.n, j_begin + p.jchunk);
    for (int64_t c0 = j_begin; c0 < j_end; c0 += JC) {
        const int cl = (int)min((int64_t)JC, j_end - c0);
        __syncthreads();
        for (int idx = tid; idx < D * JC; idx += kDenseThreads) {
            const int t = idx / JC, jj = idx % JC;
            xc[idx] = jj < cl ? p.xs2[t * p.ld2 + c0 + jj] : T(0);
        }
        for (int idx = tid; idx < JC * TP; idx += kDenseThreads) {
            const int jj = idx / TP, cc = idx % TP;
            const int gc = colbase + cc;
            Vs[idx] = (jj < cl && gc < p.p) ? p.V[(c0 + jj) * p.ldv + gc] : T(0);
        }
        __syncthreads();
        // phase A: 64 x 32 kernel tile, 8 values per thread
#pragma unroll 2
        for (int e = 0; e < JC / 4; ++e) {
            const int jj = jsub + 4 * e;
            T q = T(0);
#pragma unroll
            for (int t = 0; t < D; ++t) {
                const T df = ar[t] - xc[t * JC + jj];
                q = fma(df, df, q);
            }
            Ks[jj * TM + rowA] = kernel_value<T, FAM>(q, tab_s);
        }
        __syncthreads();
        // phase B: acc[8][4] += K'[rows of this warp][jj] * V[jj][cols of this lane]
#pragma unroll 4
        for (int jj = 0; jj < JC; ++jj) {
            T a[8], b[4];
            load4<T>(Ks + jj * TM + warp * 8, *reinterpret_cast<T(*)[4]>(&a[0]));
            load4<T>(Ks + jj * TM + warp * 8 + 4, *reinterpret_cast<T(*)[4]>(&a[4]));
            load4<T>(Vs + jj * TP + lane * 4, b);
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
    }
    const bool atomic = gridDim.y > 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int64_t i = row0 + warp * 8 + r;
        if (i >= p.m) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int gc = colbase + lane * 4 + c;
            if (gc < p.p) {
                T* dst = p.Y + i * p.ldy + gc;
                if (atomic) atomicAdd(dst, acc[r][c]);
                else *dst = acc[r][c];
            }
        }
    }
}

// ---- lengthscale cotangent:  ls_bar_t = const_t * sum_ij (Ybar_i . V_j) H_ij q_ij(t) -----------------
template <typename T>
struct DenseLsArgs {
    const T* xs1; int64_t ld1; int64_t m;
    const T* xs2; int64_t ld2; int64_t n;
    const T* Ybt; int64_t ldyb;  // Ybar transposed: (p, m)
    const T* Vt; int64_t ldv;    // V transposed: (p, n)
    int p;
    T* ls_part;  // [gridDim.y * gridDim.x][NLS]
    int64_t jchunk;
};

template <typename T, int D, int FAM, bool ARD>
__global__ void __launch_bounds__(128, 3) ks_dense_bwd_ls_kernel(const DenseLsArgs<T> p) {
    constexpr int TM = 64, TN = 64, PC = 32, NLS = ARD ? D : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* Ys = tab + ExpTab<T>::kSmemElems;  // [PC][TM]
    T* Vs = Ys + PC * TM;                 // [PC][TN]
    __shared__ T red[4][NLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, 128);
    const int rg = tid >> 4, cg = tid & 15;  // 8 row groups x 16 column groups
    const int64_t i0 = (int64_t)blockIdx.x * TM;
    T lacc[NLS];
#pragma unroll
    for (int t = 0; t < NLS; ++t) lacc[t] = T(0);
    const int64_t j_begin = (int64_t)blockIdx.y * p.jchunk;
    const int64_t j_end = min(p.n, j_begin + p.jchunk);
    for (int64_t j0 = j_begin; j0 < j_end; j0 += TN) {
        T z[8][4];
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) z[r][c] = T(0);
        for (int c0 = 0; c0 < p.p; c0 += PC) {
            __syncthreads();
            for (int idx = tid; idx < PC * TM; idx += 128) {
                const int cc = idx / TM, ii = idx % TM;  // coalesced along the point axis (transposed operands)
                const int64_t gi = i0 + ii;
                Ys[idx] = (gi < p.m && c0 + cc < p.p) ? p.Ybt[(int64_t)(c0 + cc) * p.ldyb + gi] : T(0);
                const int64_t gj = j0 + ii;
                Vs[idx] = (gj < j_end && c0 + cc < p.p) ? p.Vt[(int64_t)(c0 + cc) * p.ldv + gj] : T(0);
            }
            __syncthreads();
#pragma unroll 4
            for (int cc = 0; cc < PC; ++cc) {
                T a[8], b[4];
                load4<T>(Ys + cc * TM + rg * 8, *reinterpret_cast<T(*)[4]>(&a[0]));
                load4<T>(Ys + cc * TM + rg * 8 + 4, *reinterpret_cast<T(*)[4]>(&a[4]));
                load4<T>(Vs + cc * TN + cg * 4, b);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) z[r][c] = fma(a[r], b[c], z[r][c]);
            }
        }
        // epilogue: weight the scores with H q
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int64_t gi = i0 + rg * 8 + r;
            if (gi >= p.m) continue;
            T xr[D];
#pragma unroll
            for (int t = 0; t < D; ++t) xr[t] = p.xs1[t * p.ld1 + gi];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int64_t gj = j0 + cg * 4 + c;
                if (gj >= j_end) continue;
                T q = T(0), qt[D];
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T df = xr[t] - p.xs2[t * p.ld2 + gj];
                    qt[t] = df * df;
                    q += qt[t];
                }
                T kv, h;
                kernel_value_grad<T, FAM>(q, tab_s, kv, h);
                const T zh = z[r][c] * h;
                if (ARD) {
#pragma unroll
                    for (int t = 0; t < D; ++t) lacc[t] = fma(zh, qt[t], lacc[t]);
                } else {
                    lacc[0] = fma(zh, q, lacc[0]);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NLS; ++t) {
        T v = lacc[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][t] = v;
    }
    __syncthreads();
    if (tid < NLS) p.ls_part[((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * NLS + tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
}

}  // namespace cagp
