// Does an SFU-pipe instruction (MUFU.RCP64H, MUFU.RSQ64H, MUFU.RCP, MUFU.EX2, F2F.F32.F64, F2F.F64.F32) take time away from the
// FP64 pipe when the SFU itself is far from saturated?  And how many honest (non-foldable) integer-pipe instructions fit
// beside a saturated FP64 pipe?  16 independent DFMA chains per thread; per 16 DFMAs, K other instructions (K = 1, 2, 4).
// Prints cycles per 16-DFMA body per SM sub-partition: 32.x = FP64-bound; every cycle above that is what the extra costs.
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_NONE = 0, OP_RCP64H, OP_RSQ64H, OP_RCP32, OP_EX2, OP_F2F_DS, OP_F2F_SD, OP_ALU, OP_IMAD, OP_FFMA, OP_LDS, OP_SHFL };

template <int OP, int K>
__global__ void __launch_bounds__(512) k(int iters, double* sink) {
    __shared__ double sh[1024];
    sh[threadIdx.x] = 1.0 + threadIdx.x; sh[threadIdx.x + 512] = 2.0;
    __syncthreads();
    double a[16];
    unsigned x[4];
    double d[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = 0x3f800000u + (threadIdx.x << 8) + i; d[i] = 1.5 + i + threadIdx.x * 1e-3; }
    const double m = 1.0000001, c = 1e-9;
    const unsigned y = (blockIdx.x & 0xf) | 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
            if (i * K / 16 != (i + 1) * K / 16) {       // K evenly spread extras per body
                const int q = (i * K / 16) & 3;
                // every extra reads a value that changes per iteration and its result feeds the next one: nothing can be folded or hoisted
                if (OP == OP_RCP64H) asm volatile("{ .reg .b32 lo, hi; rcp.approx.ftz.f64 %0, %0; mov.b64 {lo, hi}, %0; xor.b32 hi, hi, %1; mov.b64 %0, {lo, hi}; }" : "+d"(d[q]) : "r"(y << 8));
                if (OP == OP_RSQ64H) asm volatile("{ .reg .b32 lo, hi; rsqrt.approx.ftz.f64 %0, %0; mov.b64 {lo, hi}, %0; xor.b32 hi, hi, %1; mov.b64 %0, {lo, hi}; }" : "+d"(d[q]) : "r"(y << 8));
                if (OP == OP_RCP32) asm volatile("{ .reg .f32 t; mov.b32 t, %0; rcp.approx.ftz.f32 t, t; mov.b32 %0, t; xor.b32 %0, %0, %1; }" : "+r"(x[q]) : "r"(y));
                if (OP == OP_EX2) asm volatile("{ .reg .f32 t; mov.b32 t, %0; ex2.approx.ftz.f32 t, t; mov.b32 %0, t; xor.b32 %0, %0, %1; and.b32 %0, %0, 0x3fffffff; }" : "+r"(x[q]) : "r"(y));
                if (OP == OP_F2F_DS) asm volatile("{ .reg .f32 t; .reg .b32 u; cvt.rn.f32.f64 t, %1; mov.b32 u, t; xor.b32 %0, %0, u; }" : "+r"(x[q]) : "d"(a[i]));
                if (OP == OP_F2F_SD) asm volatile("{ .reg .f32 t; .reg .f64 v; .reg .b32 lo, hi; xor.b32 %0, %0, %1; mov.b32 t, %0; cvt.f64.f32 v, t; mov.b64 {lo, hi}, v; xor.b32 %0, %0, hi; and.b32 %0, %0, 0x3fffffff; }" : "+r"(x[q]) : "r"(y));
                if (OP == OP_ALU) asm volatile("{ .reg .b32 t; shr.u32 t, %0, 7; xor.b32 %0, %0, t; add.u32 %0, %0, %1; }" : "+r"(x[q]) : "r"(y));     // 3 ALU instructions
                if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, 1664525, %1;" : "+r"(x[q]) : "r"(y));
                if (OP == OP_FFMA) asm volatile("{ .reg .f32 t; mov.b32 t, %0; fma.rn.f32 t, t, 0f3F7FFF00, 0f3A000000; mov.b32 %0, t; }" : "+r"(x[q]));
                if (OP == OP_LDS) asm volatile("{ .reg .b32 lo, hi; .reg .f64 v; ld.shared.f64 v, [%1]; mov.b64 {lo, hi}, v; xor.b32 %0, %0, lo; }" : "+r"(x[q]) : "r"((unsigned)__cvta_generic_to_shared(&sh[(threadIdx.x + (x[q] & 0x1e0)) & 1023])));
                if (OP == OP_SHFL) asm volatile("{ .reg .b32 t; shfl.sync.bfly.b32 t, %0, 1, 0x1f, 0xffffffff; xor.b32 %0, %0, t; add.u32 %0, %0, %1; }" : "+r"(x[q]) : "r"(y));
            }
        }
    }
    double r = 0;
    unsigned xr = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { r += d[i]; xr ^= x[i]; }
    if (r == 123.456 || xr == 0xdeadbeef) sink[0] = r + xr;
}

template <int OP, int K>
void run(const char* name, double* sink) {
    const int iters = 10000, threads = 512, blocks = 148;   // 16 warps per SM = 4 per sub-partition
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
    const double cycles = ms * 1e-3 * khz * 1e3 / iters / 4;
    static double base = 0;
    if (OP == OP_NONE) base = cycles;
    printf("%-14s K=%2d per 16 DFMA: %7.2f cycles per body per SMSP  (+%.2f per extra)\n", name, K, cycles, K ? (cycles - base) / K : 0.0);
}
#define RUN3(OP, name) run<OP, 1>(name, sink); run<OP, 2>(name, sink); run<OP, 4>(name, sink);
int main() {
    double* sink; cudaMalloc(&sink, 8);
    run<OP_NONE, 0>("none", sink);
    RUN3(OP_RCP64H, "mufu.rcp64h") RUN3(OP_RSQ64H, "mufu.rsq64h") RUN3(OP_RCP32, "mufu.rcp") RUN3(OP_EX2, "mufu.ex2")
    RUN3(OP_F2F_DS, "f2f.f32.f64") RUN3(OP_F2F_SD, "f2f.f64.f32")
    run<OP_ALU, 4>("3 alu", sink); run<OP_ALU, 8>("3 alu", sink); run<OP_ALU, 16>("3 alu", sink);
    run<OP_IMAD, 4>("imad", sink); run<OP_IMAD, 8>("imad", sink); run<OP_IMAD, 16>("imad", sink);
    run<OP_FFMA, 4>("ffma", sink); run<OP_FFMA, 8>("ffma", sink); run<OP_FFMA, 16>("ffma", sink);
    run<OP_LDS, 2>("lds.64", sink); run<OP_LDS, 4>("lds.64", sink); run<OP_LDS, 8>("lds.64", sink);
    run<OP_SHFL, 2>("shfl", sink); run<OP_SHFL, 4>("shfl", sink); run<OP_SHFL, 8>("shfl", sink);
    return 0;
}
