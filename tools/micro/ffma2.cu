// Packed FP32 on sm_100: does FFMA2 (fma.rn.f32x2, two FMAs per lane and instruction) issue at the rate of FFMA?
// 32 warps per SM, 8 independent chains per thread.  Prints instructions and FMA lane-operations per clock per SM.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) k(int iters, float* sink, float m_, float c_) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(1.0f + threadIdx.x * 1e-6f + i, 0.5f + i);
    const float2 m = make_float2(m_, m_ * 0.999f), c = make_float2(c_, c_ * 1.001f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); }                                  // FFMA, one chain per register
                if (MODE == 1) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); } // two FFMA
                if (MODE == 2) { a[i] = __ffma2_rn(a[i], m, c); }                                    // FFMA2
                if (MODE == 3) { a[i] = __fmul2_rn(a[i], m); }                                       // FMUL2
                if (MODE == 4) { a[i] = __fadd2_rn(a[i], c); }                                       // FADD2
                if (MODE == 5) { a[i] = __ffma2_rn(a[i], a[(i + 3) & 7], a[(i + 5) & 7]); }          // FFMA2, three distinct register pairs
            }
    }
    float r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
    if (r == 123.456f) sink[0] = r;
}
template <int MODE>
void run(const char* name, float* sink, double lanes_ops_per_instr, double instr_per_body) {
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 2, 512>>>(100, sink, 1.0000001f, 1e-9f);
    cudaEventRecord(e0);
    k<MODE><<<148 * 2, 512>>>(iters, sink, 1.0000001f, 1e-9f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * khz * 1e3;
    const double warp_instr = 148.0 * 2 * 16 * iters * 32.0 * instr_per_body / 32.0;    // per body: 4 x 8 ops
    printf("%-26s %.2f warp instructions / clk / SM, %.0f fp32 lane-ops / clk / SM\n", name, warp_instr / 148 / cycles,
           warp_instr / 148 / cycles * 32 * lanes_ops_per_instr);
}
int main() {
    float* sink; cudaMalloc(&sink, 4);
    run<0>("ffma (8 chains)", sink, 1, 1); run<1>("ffma (16 chains)", sink, 1, 2); run<2>("ffma2", sink, 2, 1);
    run<3>("fmul2", sink, 2, 1); run<4>("fadd2", sink, 2, 1); run<5>("ffma2 3 reg pairs", sink, 2, 1);
    return 0;
}
