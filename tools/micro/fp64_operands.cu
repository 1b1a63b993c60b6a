// How fast is the FP64 pipe when every DFMA reads three DIFFERENT 64-bit register operands out of a large live set (as real
// kernels do), compared with the usual peak microbenchmark whose chains share the multiplier and the addend?
// 16 warps per SM; prints cycles per warp instruction per SM sub-partition (2.0 = nominal rate of the 16-lane pipe).
#include <cstdio>
#include <cuda_runtime.h>

template <int NREG, int MODE>
__global__ void __launch_bounds__(512) k(int iters, double* sink, const double* init) {
    double a[NREG];
#pragma unroll
    for (int i = 0; i < NREG; ++i) a[i] = init[i] + threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NREG; ++i) {
            if (MODE == 0) a[i] = fma(a[i], a[(i + 5) % NREG], a[(i + 11) % NREG]);       // 3 distinct registers, accumulate in place
            if (MODE == 1) a[i] = fma(a[(i + 3) % NREG], a[(i + 5) % NREG], a[(i + 11) % NREG]);   // 3 sources + a 4th destination
            if (MODE == 2) a[i] = a[i] * a[(i + 5) % NREG];                               // DMUL, 2 distinct sources
            if (MODE == 3) a[i] = a[(i + 3) % NREG] + a[(i + 5) % NREG];                  // DADD, 2 sources + other destination
            if (MODE == 4) a[i] = fma(a[i], 0.999999, a[(i + 11) % NREG]);                // one immediate
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < NREG; ++i) r += a[i];
    if (r == 123.456) sink[0] = r;
}
template <int NREG, int MODE>
void run(const char* name, double* sink, const double* init) {
    const int iters = 8000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<NREG, MODE><<<148, 512>>>(50, sink, init);
    cudaEventRecord(e0);
    k<NREG, MODE><<<148, 512>>>(iters, sink, init);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cyc = ms * 1e-3 * khz * 1e3 / iters;              // per loop body; 4 warps of an SMSP run it concurrently
    printf("%-28s live doubles %2d : %.2f cycles per warp instruction per SMSP\n", name, NREG, cyc / (4.0 * NREG));
}
int main() {
    double h[48];
    for (int i = 0; i < 48; ++i) h[i] = 0.5 + 1e-3 * i;
    double *sink, *init;
    cudaMalloc(&sink, 8); cudaMalloc(&init, sizeof h); cudaMemcpy(init, h, sizeof h, cudaMemcpyHostToDevice);
    run<16, 0>("dfma a=a*b+c (3 regs)", sink, init); run<32, 0>("dfma a=a*b+c (3 regs)", sink, init); run<48, 0>("dfma a=a*b+c (3 regs)", sink, init);
    run<16, 1>("dfma d=a*b+c (4 regs)", sink, init); run<32, 1>("dfma d=a*b+c (4 regs)", sink, init); run<48, 1>("dfma d=a*b+c (4 regs)", sink, init);
    run<32, 2>("dmul a=a*b", sink, init); run<32, 3>("dadd d=a+b", sink, init); run<32, 4>("dfma a=a*imm+c", sink, init);
    return 0;
}
