// Dense right-hand side kernels: Y = K'(X1, X2) V for a dense V (n, p) and the lengthscale part of
// its reverse mode.  This is LazyKernel._matmat for dense operands (operators/lazy_kernel.py:79-90):
// matvecs (p = 1) and dense action matrices (BASELINE config 5).  sm_100a.
//
// Unlike the block-sparse case the contraction over j is a real GEMM (2 m n p flops), so a kernel
// value must be shared by all p output columns.  A CTA owns 64 rows and up to 128 output columns and
// walks the n columns of K' in chunks of JC = 32:
//   phase A: the 256 threads evaluate the 64 x 32 kernel tile once (8 values each, row coordinates in
//            registers, column coordinates broadcast from shared memory) and park it in shared memory;
//   phase B: the 64 x 128 output tile accumulates K'tile (64 x 32) * Vtile (32 x 128).  fp64: on the
//            tensor core, mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; tcgen05 has no f64 kind), each warp a
//            32 x 32 sub-tile = 16 accumulator fragments, 8 conflict-free LDS.64 per 16 MMAs.  fp32: an
//            FMA register tile (8 rows x 4 columns per thread).
// p > 128 takes ceil(p / 128) passes (grid.z) that re-evaluate the kernel tile (45 of ~300 fp64 slots
// per pair at d = 16).  On B200 DMMA and DFMA share one datapath (37.1 vs 33.9 TFLOP/s measured,
// profiles/r01_issue_model.txt): the MMA form wins by needing 8x fewer issue slots, not by a higher peak.
#pragma once
#include "ks_blocksparse.cuh"

namespace cagp {

constexpr int kDenseRows = 64;    // rows per CTA
constexpr int kDenseCols = 128;   // output columns per CTA pass
constexpr int kDenseJC = 32;      // kernel columns per stage
constexpr int kDenseThreads = 256;
// shared-memory row strides: == 4 (mod 16) doubles so that the 16 lanes of a half-warp loading an
// mma.m8n8k4 fragment (4 k-values x 4 rows/cols) hit 16 different bank pairs
constexpr int kDenseKsStride = kDenseRows + 4;   // 68
constexpr int kDenseVsStride = kDenseCols + 4;   // 132

// D (8x8) += A (8x4, row) * B (4x8, col) in fp64 on the tensor core (DMMA.8x8x4).  Fragments: lane l holds
// A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

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

// 8- or 4-byte asynchronous global -> shared copy; !valid writes zeros (src-size 0, the source is not read)
template <typename T>
__device__ __forceinline__ void cp_async_elem(T* smem_dst, const T* gsrc, bool valid) {
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? (int)sizeof(T) : 0;
    if constexpr (sizeof(T) == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Stage pipeline: the V tile and the column coordinates of stage s + 1 travel global -> shared with cp.async
// while stage s is evaluated (phase A) and contracted (phase B); two barriers per stage.
template <typename T, int D, int FAM>
__global__ void __launch_bounds__(kDenseThreads, 2) ks_dense_fwd_kernel(const DenseFwdArgs<T> p) {
    constexpr int JC = kDenseJC, TM = kDenseRows, TP = kDenseCols;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* xc0 = tab + ExpTab<T>::kSmemElems;  // [2][D][JC]
    constexpr int KS = kDenseKsStride, VS = kDenseVsStride;
    constexpr bool kMma = sizeof(T) == 8;  // fp64: tensor-core MMA for the GEMM phase; fp32: FMA register tile
    T* Ks = xc0 + 2 * D * JC;             // [JC][KS]
    T* Vs0 = Ks + JC * KS;                // [2][JC][VS]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, kDenseThreads);
    // MMA warp tile: 32 rows x 32 columns (2 x 4 warp grid): 4 x 4 m8n8 accumulator tiles per warp
    const int wm = warp & 1, wn = warp >> 1;
    const int grp = lane >> 2, tig = lane & 3;

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
    auto stage_load = [&](int buf, int64_t c0) {
        const int cl = (int)min((int64_t)JC, j_end - c0);
        T* xc = xc0 + buf * (D * JC);
        T* Vs = Vs0 + buf * (JC * VS);
        for (int idx = tid; idx < D * JC; idx += kDenseThreads) {
            const int t = idx / JC, jj = idx % JC;
            const bool ok = jj < cl;
            cp_async_elem<T>(xc + idx, p.xs2 + t * p.ld2 + (ok ? c0 + jj : j_begin), ok);
        }
        for (int idx = tid; idx < JC * TP; idx += kDenseThreads) {
            const int jj = idx / TP, cc = idx % TP;
            const int gc = colbase + cc;
            const bool ok = jj < cl && gc < p.p;
            cp_async_elem<T>(Vs + jj * VS + cc, p.V + (ok ? (c0 + jj) * p.ldv + gc : 0), ok);
        }
        cp_async_commit();
    };
    if (j_begin < j_end) stage_load(0, j_begin);
    int buf = 0;
    for (int64_t c0 = j_begin; c0 < j_end; c0 += JC, buf ^= 1) {
        const int cl = (int)min((int64_t)JC, j_end - c0);
        const T* xc = xc0 + buf * (D * JC);
        const T* Vs = Vs0 + buf * (JC * VS);
        cp_async_wait_all();
        __syncthreads();  // stage data visible; everybody is done with phase B of the previous stage
        if (c0 + JC < j_end) stage_load(buf ^ 1, c0 + JC);
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
            Ks[jj * KS + rowA] = (jj < cl) ? kernel_value<T, FAM>(q, tab_s) : T(0);
        }
        __syncthreads();
        if constexpr (kMma) {
            // phase B on the fp64 tensor core: acc[mt*4+nt] += K'[32 rows of the warp][4 k] * V[4 k][32 cols]
#pragma unroll 2
            for (int ks = 0; ks < JC / 4; ++ks) {
                double af[4], bf[4];
                const double* kp = reinterpret_cast<const double*>(Ks) + (ks * 4 + tig) * KS + wm * 32 + grp;
                const double* vp = reinterpret_cast<const double*>(Vs) + (ks * 4 + tig) * VS + wn * 32 + grp;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    af[i] = kp[i * 8];
                    bf[i] = vp[i * 8];
                }
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        dmma_m8n8k4(*reinterpret_cast<double(*)[2]>(&acc[mt * 2 + nt / 2][(nt & 1) * 2]), af[mt], bf[nt]);
            }
        } else {
            // phase B: acc[8][4] += K'[rows of this warp][jj] * V[jj][cols of this lane]
#pragma unroll 4
            for (int jj = 0; jj < JC; ++jj) {
                T a[8], b[4];
                load4<T>(Ks + jj * KS + warp * 8, *reinterpret_cast<T(*)[4]>(&a[0]));
                load4<T>(Ks + jj * KS + warp * 8 + 4, *reinterpret_cast<T(*)[4]>(&a[4]));
                load4<T>(Vs + jj * VS + lane * 4, b);
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
            }
        }
    }
    const bool atomic = gridDim.y > 1;
    if constexpr (kMma) {
        // accumulator tile (mt, nt) lives in acc[mt*2 + nt/2][(nt&1)*2 + {0,1}]: row 8 mt + grp, cols 8 nt + 2 tig + {0,1}
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) {
            const int64_t i = row0 + wm * 32 + mt * 8 + grp;
            if (i >= p.m) continue;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int gc = colbase + wn * 32 + nt * 8 + 2 * tig + e;
                    if (gc < p.p) {
                        T* dst = p.Y + i * p.ldy + gc;
                        const T v = acc[mt * 2 + nt / 2][(nt & 1) * 2 + e];
                        if (atomic) atomicAdd(dst, v);
                        else *dst = v;
                    }
                }
        }
    } else {
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
    constexpr int TM = 64, TN = 64, PC = 32, NLS = ARD ? D : 1, ST = kDenseKsStride;
    constexpr bool kMma = sizeof(T) == 8;  // fp64: score tile on the tensor core (DMMA); fp32: FMA register tile
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* tab = reinterpret_cast<T*>(smem_raw);
    T* Ys = tab + ExpTab<T>::kSmemElems;  // [PC][ST]
    T* Vs = Ys + PC * ST;                 // [PC][ST]
    __shared__ T red[4][NLS];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned tab_s = ExpTab<T>::load(tab, tid, 128);
    const int rg = tid >> 4, cg = tid & 15;  // FMA path: 8 row groups x 16 column groups
    const int wm = warp & 1, wn = warp >> 1, grp = lane >> 2, tig = lane & 3;  // MMA path: 2 x 2 warps of 32 x 32
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
                Ys[cc * ST + ii] = (gi < p.m && c0 + cc < p.p) ? p.Ybt[(int64_t)(c0 + cc) * p.ldyb + gi] : T(0);
                const int64_t gj = j0 + ii;
                Vs[cc * ST + ii] = (gj < j_end && c0 + cc < p.p) ? p.Vt[(int64_t)(c0 + cc) * p.ldv + gj] : T(0);
            }
            __syncthreads();
            if constexpr (kMma) {
#pragma unroll 2
                for (int ks = 0; ks < PC / 4; ++ks) {
                    double af[4], bf[4];
                    const double* yp = reinterpret_cast<const double*>(Ys) + (ks * 4 + tig) * ST + wm * 32 + grp;
                    const double* vp = reinterpret_cast<const double*>(Vs) + (ks * 4 + tig) * ST + wn * 32 + grp;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        af[i] = yp[i * 8];
                        bf[i] = vp[i * 8];
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt)
                            dmma_m8n8k4(*reinterpret_cast<double(*)[2]>(&z[mt * 2 + nt / 2][(nt & 1) * 2]), af[mt], bf[nt]);
                }
            } else {
#pragma unroll 4
                for (int cc = 0; cc < PC; ++cc) {
                    T a[8], b[4];
                    load4<T>(Ys + cc * ST + rg * 8, *reinterpret_cast<T(*)[4]>(&a[0]));
                    load4<T>(Ys + cc * ST + rg * 8 + 4, *reinterpret_cast<T(*)[4]>(&a[4]));
                    load4<T>(Vs + cc * ST + cg * 4, b);
#pragma unroll
                    for (int r = 0; r < 8; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) z[r][c] = fma(a[r], b[c], z[r][c]);
                }
            }
        }
        // epilogue: weight the scores with H q.  Pair of z[ri][ci]:
        //   MMA: tile (mt, nt) element e -> row 32 wm + 8 mt + grp, column 32 wn + 8 nt + 2 tig + e
        //   FMA: row 8 rg + ri, column 4 cg + ci
#pragma unroll
        for (int ri = 0; ri < (kMma ? 4 : 8); ++ri) {
            const int64_t gi = i0 + (kMma ? wm * 32 + ri * 8 + grp : rg * 8 + ri);
            if (gi >= p.m) continue;
            T xr[D];
#pragma unroll
            for (int t = 0; t < D; ++t) xr[t] = p.xs1[t * p.ld1 + gi];
#pragma unroll
            for (int ci = 0; ci < (kMma ? 8 : 4); ++ci) {
                const int64_t gj = j0 + (kMma ? wn * 32 + (ci >> 1) * 8 + 2 * tig + (ci & 1) : cg * 4 + ci);
                if (gj >= j_end) continue;
                const T zv = kMma ? z[ri * 2 + (ci >> 2)][ci & 3] : z[ri][ci];
                T q = T(0), qt[D];
#pragma unroll
                for (int t = 0; t < D; ++t) {
                    const T df = xr[t] - p.xs2[t * p.ld2 + gj];
                    qt[t] = df * df;
                    q += qt[t];
                }
                T kv, h;
                kernel_value_grad<T, FAM>(q, tab_s, kv, h);
                const T zh = zv * h;
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
