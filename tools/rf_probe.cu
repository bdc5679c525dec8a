// Register-file probe: cost of a DFMA / DADD / DMUL as a function of how many of its register operands are FRESH
// (not served by the operand-reuse cache).  8 independent chains per thread, 8 CTAs of 256 threads per SM.
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: acc = fma(acc, a, b)          (a, b shared by all chains: .reuse)            1 fresh pair
// MODE 1: acc_i = fma(x_i, a, acc_i)    (x_i distinct per chain)                        2 fresh pairs
// MODE 2: acc_i = fma(x_i, y_i, acc_i)  (x_i, y_i distinct)                             3 fresh pairs
// MODE 3: acc_i = fma(x_i, y_j, acc_i) with y_j rotating so that consecutive DFMAs share y (reuse on one slot)
// MODE 4: pair-kernel like mix per chain: u = fma(x_i, a, b); t = u + M; p = fma(u, p, c); acc_i = fma(p, y_i, acc_i)
template <int MODE>
__global__ void k_rf(double* out, const double* in, int iters) {
    constexpr int ILP = 8;
    double acc[ILP], x[ILP], y[ILP];
    const double a = in[threadIdx.x], b = in[256 + threadIdx.x];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { acc[i] = in[(2 + i) * 256 + threadIdx.x]; x[i] = in[(10 + i) * 256 + threadIdx.x]; y[i] = in[(18 + i) * 256 + threadIdx.x]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) acc[i] = fma(acc[i], a, b);
            if (MODE == 1) acc[i] = fma(x[i], a, acc[i]);
            if (MODE == 2) acc[i] = fma(x[i], y[i], acc[i]);
            if (MODE == 3) acc[i] = fma(x[i], y[(i >> 2)], acc[i]);
            if (MODE == 5) acc[i] = acc[i] + x[i];
            if (MODE == 6) acc[i] = acc[i] * x[i];
            if (MODE == 7) acc[i] = acc[i] + a;
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}
template <int MODE>
void run(const char* name, double* dout, double* din, int sms) {
    const int threads = 256, ctas = sms * 8, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_rf<MODE><<<ctas, threads>>>(dout, din, iters);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0); k_rf<MODE><<<ctas, threads>>>(dout, din, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double n = (double)threads * ctas * iters * 8;
    double cyc = best * 1e-3 * 1.965e9 / (n / 32.0 / (sms * 4.0));
    printf("%-44s %.3f SMSP-cycles per instruction\n", name, cyc);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount;
    double *dout, *din; cudaMalloc(&dout, 1024); cudaMalloc(&din, 32 * 256 * 8);
    static double h[32 * 256]; unsigned st = 12345u;
    for (int i = 0; i < 32 * 256; ++i) { st = st * 1664525u + 1013904223u; h[i] = 1.0 + 1e-9 * (st >> 8); }
    for (int i = 256; i < 512; ++i) h[i] *= 1e-9;
    cudaMemcpy(din, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("dfma acc=fma(acc,a,b)   (2 operands reused)", dout, din, sms);
    run<1>("dfma acc=fma(x_i,a,acc) (1 reused, 2 fresh)", dout, din, sms);
    run<2>("dfma acc=fma(x_i,y_i,acc) (3 fresh)", dout, din, sms);
    run<3>("dfma acc=fma(x_i,y_j,acc) (y shared by 4)", dout, din, sms);
    run<5>("dadd acc=acc+x_i (2 fresh)", dout, din, sms);
    run<7>("dadd acc=acc+a (1 fresh)", dout, din, sms);
    run<6>("dmul acc=acc*x_i (2 fresh)", dout, din, sms);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
