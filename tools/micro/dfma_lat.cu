// DFMA latency / throughput probe: warps per SM x independent chains per thread.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(int iters, double* sink) {
    double a[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) r += a[i];
    if (r == 123.456) sink[0] = r;
}
template <int ILP>
void run(int warps_per_sm, double* sink) {
    int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int threads = 32 * warps_per_sm > 1024 ? 1024 : 32 * warps_per_sm;
    int blocks = 148 * (32 * warps_per_sm / threads);
    k<ILP><<<blocks, threads>>>(100, sink);
    cudaEventRecord(e0);
    k<ILP><<<blocks, threads>>>(iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)blocks * threads / 32 * iters * 8.0 * ILP;   // warp-level DFMAs
    double per_sm_per_clk = inst / 148 / (ms * 1e-3 * 1.965e9);
    printf("warps/SM %2d ILP %d : %.3f warp-DFMA/clk/SM  (%.1f TFLOP/s)  cycles per dependent DFMA per warp %.1f\n", warps_per_sm, ILP,
           per_sm_per_clk, inst * 64 / (ms * 1e-3) / 1e12, (ms * 1e-3 * 1.965e9) / (iters * 8.0));
}
int main() {
    double* sink; cudaMalloc(&sink, 8);
    for (int w : {4, 8, 12, 16, 20, 24, 32}) { run<1>(w, sink); run<2>(w, sink); run<4>(w, sink); }
    return 0;
}
