// tcgen05 bring-up probe (sm_100a): D[128 x 128] = A[128 x 16] * B[128 x 16]^T with ONE CTA, tcgen05.mma kind::tf32, both
// operands in shared memory (K-major, SWIZZLE_NONE core-matrix layout), accumulator in TMEM, read back with tcgen05.ld.
// Checks the shared-memory descriptor fields (start address, leading / stride byte offsets), the instruction descriptor and
// the TMEM lane <-> row mapping against a host product, element by element.  Every wait is bounded (a wrong descriptor must
// never hang the GPU): on time-out the kernel reports through `status` and leaves.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/micro/bin/tc_probe tools/micro/tc_probe.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <vector>

#include <cuda_runtime.h>

constexpr int M = 128, N = 128, K = 16;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) of an R x 16 tf32 tile: core matrices of 8 rows x 16 bytes; the 4 atoms along K are R*16 bytes
// apart (leading byte offset), the 8-row groups 128 bytes apart (stride byte offset)
__host__ __device__ inline int tile_off(int R, int row, int k) { return (k / 4) * (R * 16) + (row / 8) * 128 + (row % 8) * 16 + (k % 4) * 4; }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
    return d;                                  // base offset 0, lbo mode 0, layout type 0 = SWIZZLE_NONE
}

__global__ void __launch_bounds__(128) k_probe(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status) {
    __shared__ __align__(128) unsigned char sa[M * K * 4];
    __shared__ __align__(128) unsigned char sb[N * K * 4];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int k = 0; k < K; ++k) {
        *reinterpret_cast<float*>(sa + tile_off(M, tid, k)) = A[tid * K + k];
        *reinterpret_cast<float*>(sb + tile_off(N, tid, k)) = B[tid * K + k];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy smem writes -> visible to the tensor core (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        for (int ks = 0; ks < K / 8; ++ks) {                           // UMMA_K = 8 for tf32: two 16-byte atoms per instruction
            const uint64_t ad = make_desc(smem_u32(sa) + ks * 2 * (M * 16), M * 16, 128);
            const uint64_t bd = make_desc(smem_u32(sb) + ks * 2 * (N * 16), N * 16, 128);
            const uint32_t acc = ks > 0;
            asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tbase), "l"(ad),
                         "l"(bd), "r"(idesc), "r"(acc)
                         : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // bounded wait for the commit
    uint32_t done = 0;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin)
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    if (!done) { if (tid == 0) status[0] = 1; }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (done) {
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t r[32];
            const uint32_t taddr = tbase + ((uint32_t)(32 * warp) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, "
                "%22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
                  "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
                  "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) D[tid * N + c0 + j] = __uint_as_float(r[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(128) : "memory");
}

int main() {
    std::vector<float> A(M * K), B(N * K), D(M * N, -1.0f), R(M * N);
    for (int i = 0; i < M * K; ++i) A[i] = (float)((i * 7 + 3) % 17 - 8) / 8.0f;       // multiples of 1/8: exact in tf32
    for (int i = 0; i < N * K; ++i) B[i] = (float)((i * 5 + 1) % 13 - 6) / 4.0f;
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += (double)A[i * K + k] * B[j * K + k];
            R[i * N + j] = (float)s;
        }
    float *dA, *dB, *dD;
    int* dS;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dS, 0, 4);
    k_probe<<<1, 128>>>(dA, dB, dD, dS);
    cudaError_t e = cudaDeviceSynchronize();
    int st = 0;
    cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    double worst = 0;
    for (int i = 0; i < M * N; ++i) {
        const double d = std::fabs((double)D[i] - R[i]);
        worst = std::fmax(worst, d);
        if (d > 1e-5) {
            if (bad < 8) printf("  mismatch at row %d col %d: got %g want %g\n", i / N, i % N, D[i], R[i]);
            ++bad;
        }
    }
    printf("tc_probe: cuda %s, status %d (1 = commit never arrived), %d of %d elements wrong, max |diff| %.3g\n", cudaGetErrorString(e), st, bad, M * N, worst);
    printf("D[0][0..3] = %g %g %g %g   want %g %g %g %g\n", D[0], D[1], D[2], D[3], R[0], R[1], R[2], R[3]);
    return (e == cudaSuccess && st == 0 && bad == 0) ? 0 : 1;
}
