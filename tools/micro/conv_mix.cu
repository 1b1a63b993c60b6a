// Which non-FP64 instructions are free next to a saturated FP64 pipe?  Round-2 question: the fp64 band kernel is bound by the
// FP64 pipe, so work is moved to the FP32 / integer / SFU pipes (fp32-seeded Newton steps, fp32 polynomial tails).  That only
// pays if the conversions between the two worlds do not themselves occupy the FP64 pipe.  8 independent DFMA chains per
// thread, K other instructions per 8 DFMAs; prints cycles per loop body per SM sub-partition (16.9 = FP64-bound).
#include <cstdio>
#include <cuda_runtime.h>

enum { OP_NONE = 0, OP_F2F_DS, OP_F2F_SD, OP_I2F_D, OP_F2I_D, OP_RCP32, OP_RCP64H, OP_RSQ32, OP_MIX, OP_INTCVT, OP_SHF };

template <int OP, int K>
__global__ void __launch_bounds__(512) k(int iters, double* sink) {
    double a[8], d[8];
    unsigned x[8];
    float f[8], g[8];
    double dd[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; d[i] = 1.0 + threadIdx.x * 1e-4 + i; x[i] = 0x3f800000u + (threadIdx.x << 8) + i; f[i] = 1.0f + threadIdx.x * 0.5f + i; g[i] = 2.f + i; dd[i] = 0.0; }
    const double m = 1.0000001, c = 1e-9;
    const unsigned y = blockIdx.x | 1u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(a[i]) : "d"(m), "d"(c));
#pragma unroll
            for (int j = 0; j < K / 8; ++j) {
                const int q = (i + j) & 7;
                if (OP == OP_F2F_DS) asm volatile("{ .reg .f32 t; .reg .b32 u; cvt.rn.f32.f64 t, %1; mov.b32 u, t; xor.b32 %0, %0, u; }" : "+r"(x[q]) : "d"(a[q]));
                if (OP == OP_F2F_SD) asm volatile("{ .reg .f32 t; .reg .f64 v; .reg .b32 lo, hi; xor.b32 %0, %0, %1; mov.b32 t, %0; cvt.f64.f32 v, t; mov.b64 {lo, hi}, v; xor.b32 %0, %0, hi; }" : "+r"(x[q]) : "r"(y & 0xffu));
                if (OP == OP_I2F_D) asm volatile("{ .reg .f64 v; .reg .b32 lo, hi; xor.b32 %0, %0, %1; cvt.rn.f64.s32 v, %0; mov.b64 {lo, hi}, v; xor.b32 %0, %0, hi; }" : "+r"(x[q]) : "r"(y & 0xffu));
                if (OP == OP_F2I_D) asm volatile("{ .reg .b32 u; cvt.rni.s32.f64 u, %1; xor.b32 %0, %0, u; }" : "+r"(x[q]) : "d"(a[q]));
                // MUFU results are perturbed by one ALU op (xor of low mantissa bits) so that ptxas cannot fold rcp(rcp(x))
                if (OP == OP_RCP32) asm volatile("{ .reg .f32 t; mov.b32 t, %0; rcp.approx.ftz.f32 t, t; mov.b32 %0, t; xor.b32 %0, %0, %1; }" : "+r"(x[q]) : "r"(y & 0xffu));
                if (OP == OP_RSQ32) asm volatile("{ .reg .f32 t; mov.b32 t, %0; rsqrt.approx.ftz.f32 t, t; mov.b32 %0, t; xor.b32 %0, %0, %1; }" : "+r"(x[q]) : "r"(y & 0xffu));
                if (OP == OP_RCP64H) asm volatile("{ .reg .f64 t; .reg .b32 lo, hi; rcp.approx.ftz.f64 t, %0; mov.b64 {lo, hi}, t; xor.b32 lo, lo, %1; mov.b64 %0, {lo, hi}; }" : "+d"(d[q]) : "r"(y & 0xffu));
                if (OP == OP_MIX) {      // 2 ALU + 1 FFMA per DFMA
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[q]) : "r"(y));
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(x[(q + 3) & 7]) : "r"(y));
                    asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[q]) : "f"(1.0001f));
                }
                if (OP == OP_INTCVT) {   // the integer-pipe double -> float -> double round trip used by the seeds (5 ALU ops)
                    unsigned hi, lo;
                    asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "d"(d[q]));
                    unsigned t, g;
                    asm volatile("sub.u32 %0, %1, 0x38000000;" : "=r"(t) : "r"(hi));
                    asm volatile("shf.l.wrap.b32 %0, %1, %2, 3;" : "=r"(g) : "r"(lo), "r"(t));
                    unsigned h2, l2;
                    asm volatile("shr.u32 %0, %1, 3;" : "=r"(h2) : "r"(g));
                    asm volatile("add.u32 %0, %0, 0x38000000;" : "+r"(h2));
                    asm volatile("shl.b32 %0, %1, 29;" : "=r"(l2) : "r"(g));
                    x[q] ^= h2 ^ l2;
                }
                if (OP == OP_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(x[q]) : "r"(y));
            }
        }
    }
    double r = 0;
    unsigned xr = 0; float fr = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { r += a[i] + d[i] + dd[i]; xr ^= x[i]; fr += f[i] + g[i]; }
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
    printf("%-22s K=%2d per 8 DFMA: %.2f cycles per body per SMSP\n", name, K, cycles / iters / 4);
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    run<OP_NONE, 0>("none", sink);
    run<OP_F2F_DS, 8>("cvt.f32.f64", sink); run<OP_F2F_DS, 16>("cvt.f32.f64", sink);
    run<OP_F2F_SD, 8>("cvt.f64.f32", sink); run<OP_F2F_SD, 16>("cvt.f64.f32", sink);
    run<OP_I2F_D, 8>("cvt.f64.s32", sink);
    run<OP_F2I_D, 8>("cvt.s32.f64", sink);
    run<OP_RCP32, 8>("mufu.rcp (f32)", sink); run<OP_RCP32, 16>("mufu.rcp (f32)", sink);
    run<OP_RSQ32, 8>("mufu.rsq (f32)", sink);
    run<OP_RCP64H, 8>("rcp.approx.f64", sink); run<OP_RCP64H, 16>("rcp.approx.f64", sink);
    run<OP_MIX, 8>("2 alu + 1 ffma", sink); run<OP_MIX, 16>("2 alu + 1 ffma", sink);
    run<OP_INTCVT, 8>("int d->f->d (5 alu)", sink); run<OP_INTCVT, 16>("int d->f->d (5 alu)", sink);
    run<OP_SHF, 8>("shf", sink); run<OP_SHF, 24>("shf", sink);
    return 0;
}
