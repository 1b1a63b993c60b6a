// Does an FP64 instruction (2 cycles of the 16-lane FP64 pipe per warp) leave the issue slot of the second cycle free for
// other pipes?  8 independent DFMA chains per thread, interleaved with K independent ops of another pipe per 8 DFMAs.
// Prints cycles per loop body per SM sub-partition-warp: 16 = FP64-bound (8 DFMA x 2 cycles) if the others hide completely.
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_NONE = 0, OP_ALU, OP_IMAD, OP_FFMA, OP_MUFU, OP_LDS };

template <int OP, int K>
__global__ void __launch_bounds__(512) k(int iters, double* sink) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = threadIdx.x; sh[threadIdx.x + 512] = 1.0;
    __syncthreads();
    double a[8];
    unsigned x[8];
    float f[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; x[i] = threadIdx.x + i; f[i] = threadIdx.x * 0.5f + i; }
    const double m = 1.0000001, c = 1e-9;
    const unsigned y = blockIdx.x | 1u;
    double lacc = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
#pragma unroll
            for (int j = 0; j < K / 8; ++j) {
                if (OP == OP_ALU) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[(i + j) & 7]) : "r"(y));
                if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x[(i + j) & 7]) : "r"(y));
                if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[(i + j) & 7]) : "f"(1.0001f));
                if (OP == OP_MUFU) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[(i + j) & 7]));
                if (OP == OP_LDS) { double v; asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(&sh[(threadIdx.x + 32 * (i + j)) & 1023]))); lacc += 0.0 * 0 + v * 0; }
            }
        }
    }
    double r = lacc;
    unsigned xr = 0; float fr = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r += a[i]; xr ^= x[i]; fr += f[i]; }
    if (r == 123.456 || xr == 0xdeadbeef || fr == 1.2345f) sink[0] = r + xr + fr;
}

template <int OP, int K>
void run(const char* name, double* sink) {
    const int iters = 20000, threads = 512, blocks = 148;   // 16 warps per SM = 4 per sub-partition
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP, K><<<blocks, threads>>>(100, sink);
    cudaEventRecord(e0);
    k<OP, K><<<blocks, threads>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int dev; cudaGetDevice(&dev);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double cycles = ms * 1e-3 * khz * 1e3;
    printf("%-5s K=%2d per 8 DFMA: %.2f cycles per body per warp-slot (4 warps/SMSP -> %.2f per SMSP-body)\n", name, K, cycles / iters, cycles / iters / 4);
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    run<OP_NONE, 0>("none", sink);
    run<OP_ALU, 8>("alu", sink);  run<OP_ALU, 16>("alu", sink);  run<OP_ALU, 24>("alu", sink);
    run<OP_IMAD, 8>("imad", sink); run<OP_IMAD, 16>("imad", sink);
    run<OP_FFMA, 8>("ffma", sink); run<OP_FFMA, 16>("ffma", sink);
    run<OP_MUFU, 8>("mufu", sink);
    run<OP_LDS, 8>("lds", sink);
    return 0;
}
