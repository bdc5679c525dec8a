// Issue-model probe: how much does non-fp64 work cost beside a saturated fp64 pipe on sm_100a?
// Each variant interleaves NA extra instructions of a given kind per DFMA.
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

enum Kind { NONE = 0, IADD = 1, FFMA = 2, LDS_BCAST = 3, LDS_LANE = 4, DADD = 5, LOP = 6, IMAD = 7, LDS_RAND = 8,
            I2F64 = 9, F2I64 = 10, FRND64 = 11, PRMT = 12, SHFL = 13, LDS128_LANE = 14, MOV = 15 };

template <int KIND, int NUM, int DEN>  // NUM extra instrs per DEN DFMAs
__global__ void k_mix(double* out, int iters, double a, double b) {
    __shared__ double tab[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) tab[i] = 1.0 + i * 1e-9;
    __syncthreads();
    const unsigned tab_s = (unsigned)__cvta_generic_to_shared(tab);
    constexpr int ILP = 8;
    double acc[ILP];
    unsigned xi[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
    float xf[4] = {1.f, 2.f, 3.f, 4.f};
    double xd[4] = {1., 2., 3., 4.};
    double sink = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            acc[i] = fma(acc[i], a, b);
            if ((i % DEN) == 0) {
#pragma unroll
                for (int e = 0; e < NUM; ++e) {
                    const int s = (i / DEN * NUM + e) & 3;
                    if (KIND == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(xi[s]) : "r"(it));
                    if (KIND == LOP) asm volatile("xor.b32 %0, %0, %1;" : "+r"(xi[s]) : "r"(it));
                    if (KIND == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(xi[s]) : "r"(it));
                    if (KIND == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(xf[s]) : "f"(1.0001f));
                    if (KIND == DADD) asm volatile("add.f64 %0, %0, %1;" : "+d"(xd[s]) : "d"(a));
                    if (KIND == I2F64) { double v; asm volatile("cvt.rn.f64.s32 %0, %1;" : "=d"(v) : "r"(xi[s])); xd[s] = v; }
                    if (KIND == F2I64) { unsigned q; asm volatile("cvt.rni.s32.f64 %0, %1;" : "=r"(q) : "d"(acc[i])); xi[s] ^= q; }
                    if (KIND == FRND64) { double q; asm volatile("cvt.rni.f64.f64 %0, %1;" : "=d"(q) : "d"(acc[i])); xi[s] ^= (unsigned)__double2loint(q); }
                    if (KIND == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(xi[s]) : "r"(it));
                    if (KIND == MOV) asm volatile("mov.b32 %0, %1;" : "=r"(xi[s]) : "r"(xi[(s + 1) & 3]));
                    if (KIND == SHFL) asm volatile("shfl.sync.idx.b32 %0, %0, %1, 0x1f, 0xffffffff;" : "+r"(xi[s]) : "r"((it + threadIdx.x) & 31));
                    if (KIND == LDS128_LANE) { double v, w; asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v), "=d"(w) : "r"(tab_s + (unsigned)((threadIdx.x + it + e) & 1023) * 16u)); sink += (e == 0 && i == 0 && it == 12345678) ? v + w : 0; }
                    if (KIND == LDS_BCAST) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tab_s + (unsigned)((it + e) & 2047) * 8u)); sink += (e == 0 && i == 0 && it == 12345678) ? v : 0; }
                    if (KIND == LDS_LANE) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tab_s + (unsigned)((threadIdx.x + it + e) & 2047) * 8u)); sink += (e == 0 && i == 0 && it == 12345678) ? v : 0; }
                    if (KIND == LDS_RAND) { double v; xi[s] = xi[s] * 1664525u + 1013904223u; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(tab_s + ((xi[s] >> 16) & 2047) * 8u)); sink += (e == 0 && i == 0 && it == 12345678) ? v : 0; }
                }
            }
        }
    }
    double s = sink;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    s += xi[0] + xi[1] + xi[2] + xi[3] + xf[0] + xf[1] + xf[2] + xf[3] + xd[0] + xd[1] + xd[2] + xd[3];
    if (s == 12345.678) out[0] = s;
}

template <int KIND, int NUM, int DEN>
void run(const char* name, double* dout, int sms) {
    const int threads = 256, ctas = sms * 8, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_mix<KIND, NUM, DEN><<<ctas, threads>>>(dout, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_mix<KIND, NUM, DEN><<<ctas, threads>>>(dout, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double n_dfma = (double)threads * ctas * iters * 8;
    // cycles per DFMA warp-instruction per SMSP at 1.965 GHz
    double warp_instr = n_dfma / 32.0 / (sms * 4.0);
    double cyc = best * 1e-3 * 1.965e9 / warp_instr;
    printf("%-28s extra %d per %d DFMA: %.2f TF dfma, %.3f SMSP-cycles per DFMA (@1.965GHz)\n", name, NUM, DEN, 2 * n_dfma / (best * 1e-3) / 1e12, cyc);
}

template <int NDFMA, int NDMMA>
__global__ void k_dmma_mix(double* out, int iters, double a, double b) {
    double acc[8]; double c[4][2];
    double ma = threadIdx.x * 1e-6, mb = 1.0 - threadIdx.x * 1e-7;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < NDFMA) acc[i] = fma(acc[i], a, b);
            if (i < NDMMA) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i & 3][0]), "+d"(c[i & 3][1]) : "d"(ma), "d"(mb));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    if (s == 12345.678) out[0] = s;
}
template <int NDFMA, int NDMMA>
void run_dm(double* dout, int sms) {
    const int threads = 256, ctas = sms * 8, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dmma_mix<NDFMA, NDMMA><<<ctas, threads>>>(dout, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_dmma_mix<NDFMA, NDMMA><<<ctas, threads>>>(dout, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double thr = (double)threads * ctas * iters;
    double tf_dfma = 2 * thr * NDFMA / (best * 1e-3) / 1e12;
    double tf_dmma = thr / 32 * NDMMA * 512.0 / (best * 1e-3) / 1e12;
    printf("dfma x%d + dmma x%d per iter: %.3f ms, dfma %.2f TF + dmma %.2f TF = %.2f TF\n", NDFMA, NDMMA, best, tf_dfma, tf_dmma, tf_dfma + tf_dmma);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double* dout; cudaMalloc(&dout, 1024);
    run<NONE, 0, 1>("dfma only", dout, sms);
    run<IADD, 1, 1>("iadd", dout, sms);
    run<IADD, 2, 1>("iadd", dout, sms);
    run<IADD, 3, 1>("iadd", dout, sms);
    run<IADD, 1, 2>("iadd", dout, sms);
    run<LOP, 1, 1>("lop", dout, sms);
    run<IMAD, 1, 1>("imad", dout, sms);
    run<IMAD, 2, 1>("imad", dout, sms);
    run<FFMA, 1, 1>("ffma", dout, sms);
    run<FFMA, 2, 1>("ffma", dout, sms);
    run<DADD, 1, 1>("dadd", dout, sms);
    run<DADD, 1, 2>("dadd", dout, sms);
    run<LDS_BCAST, 1, 1>("lds.64 broadcast", dout, sms);
    run<LDS_BCAST, 1, 2>("lds.64 broadcast", dout, sms);
    run<LDS_BCAST, 1, 4>("lds.64 broadcast", dout, sms);
    run<LDS_LANE, 1, 2>("lds.64 per-lane consecutive", dout, sms);
    run<LDS_LANE, 1, 4>("lds.64 per-lane consecutive", dout, sms);
    run<LDS_LANE, 1, 8>("lds.64 per-lane consecutive", dout, sms);
    run<LDS_RAND, 1, 4>("lds.64 random 2048", dout, sms);
    run<LDS_RAND, 1, 8>("lds.64 random 2048", dout, sms);
    run<I2F64, 1, 1>("i2f.f64.s32", dout, sms);
    run<I2F64, 1, 2>("i2f.f64.s32", dout, sms);
    run<I2F64, 1, 4>("i2f.f64.s32", dout, sms);
    run<I2F64, 1, 8>("i2f.f64.s32", dout, sms);
    run<F2I64, 1, 2>("f2i.s32.f64", dout, sms);
    run<F2I64, 1, 4>("f2i.s32.f64", dout, sms);
    run<F2I64, 1, 8>("f2i.s32.f64", dout, sms);
    run<FRND64, 1, 4>("frnd.f64", dout, sms);
    run<FRND64, 1, 8>("frnd.f64", dout, sms);
    run<PRMT, 1, 1>("prmt", dout, sms);
    run<PRMT, 1, 2>("prmt", dout, sms);
    run<MOV, 1, 1>("mov", dout, sms);
    run<SHFL, 1, 2>("shfl", dout, sms);
    run<SHFL, 1, 4>("shfl", dout, sms);
    run<SHFL, 1, 8>("shfl", dout, sms);
    run<LDS128_LANE, 1, 2>("lds.128 per-lane consecutive", dout, sms);
    run<LDS128_LANE, 1, 4>("lds.128 per-lane consecutive", dout, sms);
    run<LDS128_LANE, 1, 8>("lds.128 per-lane consecutive", dout, sms);
    run<IMAD, 1, 2>("imad", dout, sms);
    run<LOP, 1, 2>("lop", dout, sms);
    run_dm<8,0>(dout, sms); run_dm<0,4>(dout, sms); run_dm<8,1>(dout, sms); run_dm<8,2>(dout, sms); run_dm<8,4>(dout, sms); run_dm<4,4>(dout, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
