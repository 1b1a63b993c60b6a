// Throughput of the FP64-pipe instruction flavours the band kernel uses (16 warps per SM, 8 independent chains per thread):
// cycles per warp instruction per SM sub-partition.  2.0 = the 16-lane pipe's rate.
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_k[8] = {1.0000001, 1e-9, 0.99999, 1.5e-9, 1.00001, 2e-9, 0.999999, 3e-9};
enum { F_DFMA = 0, F_DMUL, F_DADD, F_DFMA_CONST, F_DFMA_IMM, F_MIX, F_RCP64H, F_DFMA_RCP, F_DSETP };

template <int F>
__global__ void __launch_bounds__(512) k(int iters, double* sink, double m_, double c_) {
    double a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-6 + i * 1e-3;
    const double m = m_, c = c_;
    int cnt = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (F == F_DFMA) a[i] = fma(a[i], m, c);
            if (F == F_DMUL) a[i] = a[i] * m;
            if (F == F_DADD) a[i] = a[i] + c;
            if (F == F_DFMA_CONST) a[i] = fma(a[i], c_k[0], c_k[1]);
            if (F == F_DFMA_IMM) a[i] = fma(a[i], 0.5, 1.0);
            if (F == F_MIX) { if (i & 1) a[i] = a[i] * m; else if (i & 2) a[i] = a[i] + c; else a[i] = fma(a[i], m, c); }
            if (F == F_RCP64H) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[i])); a[i] = y; }
            if (F == F_DFMA_RCP) { a[i] = fma(a[i], m, c); if (i == 0) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a[7])); a[7] = y; } }
            if (F == F_DSETP) { cnt += a[i] > m ? 1 : 0; a[i] = fma(a[i], m, c); }
        }
    }
    double r = cnt;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i];
    if (r == 123.456) sink[0] = r;
}
template <int F>
void run(const char* name, double* sink, double per_body) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<F><<<148, 512>>>(100, sink, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<F><<<148, 512>>>(iters, sink, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3 / iters / 4;       // per SMSP-body (4 warps per SMSP run concurrently)
    printf("%-12s %.2f cycles per SMSP-body of %g warp instr -> %.2f cycles per instruction\n", name, cyc * 4, per_body * 4, cyc / per_body);
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    run<F_DFMA>("dfma", sink, 8); run<F_DMUL>("dmul", sink, 8); run<F_DADD>("dadd", sink, 8);
    run<F_DFMA_CONST>("dfma c[]", sink, 8); run<F_DFMA_IMM>("dfma imm", sink, 8); run<F_MIX>("mix", sink, 8);
    run<F_RCP64H>("rcp64h", sink, 8); run<F_DFMA_RCP>("dfma+1rcp", sink, 9); run<F_DSETP>("dfma+dsetp", sink, 16);
    return 0;
}
