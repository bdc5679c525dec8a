// Micro-benchmark: issue rate of DADD, DMUL, DFMA and of mixes of them on sm_100a (the pair kernels execute all
// three).  Register-resident chains, 8 independent chains per thread, 148 * 4 CTAs of 256 threads.
// Prints warp-instructions per cycle per SM-quadrant-equivalent as G-instr/s per thread-lane (T lane-ops/s).
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void k(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) acc[i] = fma(acc[i], a, b);
            if (OP == 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(b));
            if (OP == 2) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(a));
            if (OP == 3) {  // 1 DADD + 1 DFMA alternating
                if (i & 1) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(b));
                else acc[i] = fma(acc[i], a, b);
            }
            if (OP == 4) {  // 1 DMUL + 1 DFMA alternating
                if (i & 1) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(acc[i]) : "d"(a));
                else acc[i] = fma(acc[i], a, b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

template <int OP>
double run(double* d, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 4, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<148 * 4, 256>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return 148.0 * 4 * 256 * 8.0 * iters / (ms * 1e-3) / 1e12;  // T lane-instructions / s
}

int main() {
    double* d;
    cudaMalloc(&d, 8);
    const int iters = 200000;
    printf("{\"dfma_Tops\": %.2f, \"dadd_Tops\": %.2f, \"dmul_Tops\": %.2f, \"dadd_dfma_mix_Tops\": %.2f, \"dmul_dfma_mix_Tops\": %.2f}\n",
           run<0>(d, iters), run<1>(d, iters), run<2>(d, iters), run<3>(d, iters), run<4>(d, iters));
    return 0;
}
