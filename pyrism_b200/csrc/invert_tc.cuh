// LUT inversion on the 5th-generation tensor cores (tcgen05, sm_100a): the one contraction of this repository.
//
// Contract (include/prosail_b200.h, oracle/invert_oracle.py): for every observation the LUT row with the smallest
//   cost(q, m) = sum_d (s_d o_{q,d} - s_d l_{m,d})^2,   float32, ascending d, one fmaf per feature, ties to the lowest index.
// That contract is kept BIT FOR BIT.  The tensor cores only produce a candidate list:
//
//   cost = ||phi(o)||^2 + [ ||phi(l)||^2 - 2 phi(o) . phi(l) ],      phi(x)_d = s_d x_d - c_d   (c = column means of the LUT: centring
//                                                                    shrinks every product without changing any difference)
// and the bracket is ONE dot product of length 16:  A row = (-2 phi(o)_0 .. -2 phi(o)_14, 1),  B row = (phi(l)_0 .. phi(l)_14, ||phi(l)||^2)
// -- the 15 L8/ASTER band means fill the K = 16 of two tf32 MMAs exactly.  TF32 keeps 11 significant bits, so every operand is split
// x = hi + lo (both exactly representable in tf32, |x - hi - lo| <= 2^-22 |x|) and the product is accumulated as
// A_hi B_hi + A_hi B_lo + A_lo B_hi in fp32 inside the tensor core: 6 tcgen05.mma (M = 128 observations, N = 128 LUT rows, K = 8) per tile.
// Error model.  With m = ||phi(o)||, n = ||phi(l)|| the approximate bracket c^ differs from the real-arithmetic bracket of the same
// float32 operands by at most
//     eps(o, l) = KAPPA * (2 m n + n^2) + KN * n^2 + KC * (m + n) * sqrt(cost)            (2 m n + n^2 >= sum_k |A_k| |B_k|, Cauchy-Schwarz)
// KAPPA: operand representation (3 * 2^-22) plus the tensor core's fp32 accumulation over 6 MMAs; KN: float32 rounding and split of
// the norm slot (computed in double, rounded once); KC: the rounding of the centring subtraction itself.  Let l^ be the row with the
// smallest c^ and C1 its exact contract cost (one evaluation).  The contract's minimiser l* has cost <= C1, hence
// ||phi(l*)|| <= m + sqrt(C1) =: nmax, and   c^(l*) <= c^(l^) + eps(o, l^) + eps(o, nmax) + 2 GAMMA C1 =: c^(l^) + tau
// (GAMMA = float32 rounding of the 15-term contract sum).  The bound is PER OBSERVATION and uses only rows that can win -- a
// global bound over the whole LUT (round-2 first version) is dominated by bright outlier rows and flagged most observations of a
// real band-mean table.  Each thread keeps the TC_KEEP smallest c^ of its observation per LUT chunk; k_invert_tc_rerank evaluates
// the exact contract cost of the candidates within tau and applies the tie rule.  If a chunk's list is full of candidates within
// tau (the list may have dropped one) the observation is put on the fix-up list and k_invert_tc_fix redoes the listed observations
// with a plain exact scan -- correctness never depends on the bound being tight, only on it being a bound.  The bound itself is
// MONITORED in every call: the re-rank evaluates the real bracket of every kept candidate in double and records the largest
// |c^ - c| / (sum_k |A_k||B_k|) (status[2]); a candidate outside eps counts as a violation (status[3]) and sends its
// observation to the exact scan.  tests/test_gpu_inversion.py asserts the measured ratio stays below KAPPA / 4.
//
// Kernel shape (k_invert_tc): CTA = 128 observations = the 128 lanes of TMEM; 160 threads.  k_invert_tc_prep has written every
// 128-row LUT tile as the exact shared-memory image of the B operand (hi and lo parts, K-major core matrices, SWIZZLE_NONE; the
// descriptor fields were validated element by element with tools/micro/tc_probe.cu).  Warp 4, lane 0 streams those tiles through a
// 3-slot ring with cp.async.bulk + mbarrier and issues the six tcgen05.mma of each tile into one of two 128-column accumulators;
// warps 0-3 read their lane with tcgen05.ld and scan: an FMNMX tree per 8 columns and one compare against the current TC_KEEP-th
// best -- about one instruction per (observation, row) pair instead of the 30 of the FP32 kernel.  Two CTAs share an SM.  Every
// wait is bounded; a time-out is reported through `status` and sends the CTA's observations to the exact scan.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int TC_M = 128;                  // observations per CTA = TMEM lanes
constexpr int TC_N = 128;                  // LUT rows per tile = TMEM columns
constexpr int TC_K = 16;                   // 15 features + the norm slot
constexpr int TC_KEEP = 8;                 // candidates kept per (observation, LUT chunk)
constexpr int TC_SMEM_BYTES = 80 * 1024;   // A (16 KB) + three B slots (48 KB) + barriers + candidate lists (8 KB); two CTAs per SM
constexpr float TC_KAPPA = 8e-6f;          // |c^ - exact| <= KAPPA * sum |A_k| |B_k|   (default; InvertTcArgs::kappa)
constexpr float TC_KN = 4e-7f;             // norm slot: 2^-24 (double -> float) + 2^-22 (hi / lo split)
constexpr float TC_KC = 2.4e-7f;           // centring subtraction: 2 * 2^-24 (m + n) sqrt(cost), doubled
constexpr int TC_FIX_GROUP = 8;            // observations per CTA of the exact fix-up scan
constexpr float TC_GAMMA = 2e-6f;          // relative float32 rounding of the contract's 15-term fmaf sum (17 * 2^-24 < 1.1e-6)

struct InvertTcArgs {
    const float* lut;        // [n_lut][pitch_lut]
    long long n_lut, pitch_lut;
    const float* obs;        // [n_obs][pitch_obs]
    long long n_obs, pitch_obs;
    const float* scale;      // [16] sqrt(weights), zero beyond n_feat (device)
    float* center;           // [16] centres of the scaled features (device; written by k_invert_tc_center)
    int n_feat;
    long long n_tiles;       // ceil(n_lut / TC_N)
    long long tiles_per_chunk;
    int n_chunks;
    unsigned char* bpre;     // [n_tiles][2 * TC_N * TC_K * 4]: the B-operand tiles as they sit in shared memory (written by k_invert_tc_prep)
    float* cand_cost;        // [n_obs][n_chunks][TC_KEEP] approximate brackets, ascending
    int* cand_idx;           // [n_obs][n_chunks][TC_KEEP] LUT row (within this shard), -1 = empty
    int* status;             // [0] != 0: a bounded wait expired;  [1]: observations on the fix-up list;  [2]: float bits of the largest
                             // measured |c^ - c| / sum |A_k||B_k| over the kept candidates;  [3]: candidates outside the bound
    int* flags;              // [n_obs] the fix-up list: observations to redo with the exact scan (status[1] entries)
    unsigned long long* fix_best;   // [n_obs] (cost bits << 32 | row) of the exact scan, folded with atomicMin
    float kappa;
};

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (row, k) of an R x 16 tf32 operand tile, K-major, no swizzle: core matrices of 8 rows x 16 bytes; the four
// core matrices along K are R * 16 bytes apart (leading byte offset), consecutive 8-row groups 128 bytes (stride byte offset)
__device__ __forceinline__ int tc_tile_off(int R, int row, int k) { return (k >> 2) * (R * 16) + (row >> 3) * 128 + (row & 7) * 16 + (k & 3) * 4; }

__device__ __forceinline__ uint64_t tc_make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                    // descriptor version 1 (sm_100)
    return d;                                  // base offset 0, layout type 0 = SWIZZLE_NONE
}

// x = hi + lo, both with at most 11 significant bits (tf32), rounded to nearest: |x - hi - lo| <= 2^-22 |x|
__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    const float r = x - hi;                    // exact
    lo = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xFFFFE000u);
}

// store 16 values of one operand row (hi and lo parts) as four 16-byte core-matrix rows each
__device__ __forceinline__ void tc_store_row(unsigned char* __restrict__ th, unsigned char* __restrict__ tl, int R, int row, const float (&v)[TC_K]) {
#pragma unroll
    for (int a4 = 0; a4 < 4; ++a4) {
        float h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) tc_split(v[4 * a4 + j], h[j], l[j]);
        const int off = tc_tile_off(R, row, 4 * a4);
        *reinterpret_cast<float4*>(th + off) = make_float4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<float4*>(tl + off) = make_float4(l[0], l[1], l[2], l[3]);
    }
}

// ---- bounded mbarrier helpers: a wait gives up after ~2^20 polls or as soon as another role of the CTA reported a failure ----
__device__ __forceinline__ void tc_mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool tc_mbar_wait(uint64_t* bar, uint32_t parity, volatile int* fail) {
    for (int spin = 0; spin < (1 << 20); ++spin) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done)
                     : "r"(tc_smem_u32(bar)), "r"(parity)
                     : "memory");
        if (done) return true;
        if ((spin & 63) == 63 && *fail) return false;
    }
    *fail = 1;
    return false;
}
__device__ __forceinline__ void tc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tc_smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(tc_smem_u32(bar))
                 : "memory");
}

// centres of the scaled features: column means over (at most) the first 8192 rows.  ANY centre is correct; a good one only makes
// the bound eps small.  Also resets the per-call scalars.
__global__ void __launch_bounds__(256) k_invert_tc_center(const InvertTcArgs a) {
    __shared__ float acc[16][17];
    const int d = threadIdx.x & 15, lane16 = threadIdx.x >> 4;      // 16 row lanes x 16 features
    const long long n = a.n_lut < 8192 ? a.n_lut : 8192;
    float s = 0.0f;
    if (d < a.n_feat)
        for (long long r = lane16; r < n; r += 16) {
            const float x = a.lut[r * a.pitch_lut + d] * a.scale[d];
            if (fabsf(x) < 1e30f) s += x;                           // skips NaN / inf entries (their rows can never be candidates)
        }
    acc[lane16][d] = s;
    __syncthreads();
    if (threadIdx.x < 16) {
        float t = 0.0f;
        for (int i = 0; i < 16; ++i) t += acc[i][threadIdx.x];
        a.center[threadIdx.x] = threadIdx.x < a.n_feat ? t / (float)n : 0.0f;
    }
    if (threadIdx.x < 16) a.status[threadIdx.x] = 0;
}

// One pass over the LUT shard: every row becomes (phi(l), ||phi(l)||^2), split into tf32 hi / lo parts and written in the EXACT
// shared-memory image of a B-operand tile (128 rows: 8 KB of hi parts, 8 KB of lo parts, K-major core matrices).  The main
// kernel then needs no arithmetic at all to stage a tile: one 16 KB bulk asynchronous copy (cp.async.bulk + mbarrier).  The pass is
// amortised over all observation blocks (each re-reads every tile).  The norm is summed in double and rounded once (TC_KN).
__global__ void __launch_bounds__(TC_N) k_invert_tc_prep(const InvertTcArgs a) {
    const int tid = threadIdx.x;
    const long long tile = blockIdx.x;
    const long long r = tile * TC_N + tid;
    float v[TC_K];
    double n2 = 0.0;
    if (r < a.n_lut) {
        const float* lp = a.lut + r * a.pitch_lut;
#pragma unroll
        for (int d = 0; d < TC_K - 1; ++d) {
            const float x = d < a.n_feat ? __ldg(lp + d) * a.scale[d] - a.center[d] : 0.0f;
            v[d] = x;
            n2 = fma((double)x, (double)x, n2);
        }
    }
    if (r < a.n_lut && n2 < 1e29) {                                   // (false for NaN / inf features too)
        v[TC_K - 1] = (float)n2;
    } else {                                                          // rows past the shard and rows with a non-finite feature (the contract
#pragma unroll                                                        // gives those an infinite cost): an unreachable norm, never a candidate
        for (int d = 0; d < TC_K - 1; ++d) v[d] = 0.0f;
        v[TC_K - 1] = 1e30f;
    }
    unsigned char* base = a.bpre + tile * (2 * TC_N * TC_K * 4);
    tc_store_row(base, base + TC_N * TC_K * 4, TC_N, tid, v);
}

// Main kernel.  CTA = 128 observations (TMEM lanes).  Warps 0-3: epilogue (thread = observation).  Warp 4, lane 0: control -- streams
// the pre-formatted B tiles through a ring of TC_STAGES shared-memory slots with bulk asynchronous copies and issues the six
// tcgen05.mma of every tile into one of TWO 128-column TMEM accumulators, so that the tensor core works on tile t+1 while the
// epilogue warps scan tile t.  Barriers: full[s] / empty[s] per ring slot (copy landed / MMAs that read it retired), tfull[b] /
// tempty[b] per accumulator (MMAs retired / all 128 epilogue threads done reading).  Two CTAs per SM (2 x 256 TMEM columns).
constexpr int TC_STAGES = 3;
constexpr int TC_TILE_BYTES = 2 * TC_N * TC_K * 4;                  // hi + lo parts of one B tile
constexpr int TC_THREADS = TC_M + 32;

__global__ void __launch_bounds__(TC_THREADS, 2) k_invert_tc(const InvertTcArgs a) {
    extern __shared__ __align__(1024) unsigned char tc_smem[];
    unsigned char* const sAh = tc_smem;
    unsigned char* const sAl = tc_smem + 8192;
    unsigned char* const sB = tc_smem + 16384;                                   // TC_STAGES x 16 KB
    unsigned char* const misc = sB + TC_STAGES * TC_TILE_BYTES;
    uint64_t* const full = reinterpret_cast<uint64_t*>(misc);                    // [TC_STAGES]
    uint64_t* const empty = full + TC_STAGES;                                    // [TC_STAGES]
    uint64_t* const tfull = empty + TC_STAGES;                                   // [2]
    uint64_t* const tempty = tfull + 2;                                          // [2]
    uint32_t* const tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    volatile int* const s_fail = reinterpret_cast<volatile int*>(tmem_slot + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    // this thread's candidate list (ascending; TC_KEEP costs then TC_KEEP rows) lives in shared memory: insertions are rare after the
    // first tiles and must stay a real branch -- as register arrays the compiler turned the compare-and-swap steps into straight-line
    // selects that every lane of the warp executed whenever ANY lane had a hit
    float* const my_c = reinterpret_cast<float*>(misc + 256) + (tid & (TC_M - 1)) * (2 * TC_KEEP);
    int* const my_i = reinterpret_cast<int*>(my_c + TC_KEEP);

    if (tid == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { tc_mbar_init(&full[s], 1); tc_mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { tc_mbar_init(&tfull[b], 1); tc_mbar_init(&tempty[b], TC_M); }
        *s_fail = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmem_slot)), "r"(2 * TC_N) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- A operand: this thread's observation, (-2 phi(o), 1) ----
    const long long q = (long long)blockIdx.x * TC_M + tid;
    if (tid < TC_M) {
        float v[TC_K];
        const bool valid = q < a.n_obs;
        const long long row = valid ? q : 0;
#pragma unroll
        for (int d = 0; d < TC_K - 1; ++d)
            v[d] = (valid && d < a.n_feat) ? -2.0f * (a.obs[row * a.pitch_obs + d] * a.scale[d] - a.center[d]) : 0.0f;
        v[TC_K - 1] = valid ? 1.0f : 0.0f;
        tc_store_row(sAh, sAl, TC_M, tid, v);
#pragma unroll
        for (int j = 0; j < TC_KEEP; ++j) { my_c[j] = 1e29f; my_i[j] = -1; }      // 1e29: below the 1e30 that marks rows past the shard / non-finite rows
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");                  // generic-proxy stores of A -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = *tmem_slot;

    const long long t0 = (long long)blockIdx.y * a.tiles_per_chunk;               // first tile of this CTA's LUT chunk
    const long long t1g = t0 + a.tiles_per_chunk < a.n_tiles ? t0 + a.tiles_per_chunk : a.n_tiles;
    const int ntiles = (int)(t1g - t0);

    if (warp == 4) {
        if ((tid & 31) == 0) {
            // ===================== control thread: copies + MMAs =====================
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);   // f32 accum, tf32 x tf32, K-major
            auto load = [&](int t) {                                              // tile t -> ring slot t % TC_STAGES
                const int s = t % TC_STAGES;
                tc_mbar_expect_tx(&full[s], TC_TILE_BYTES);
                tc_bulk_g2s(sB + s * TC_TILE_BYTES, a.bpre + (t0 + t) * TC_TILE_BYTES, TC_TILE_BYTES, &full[s]);
            };
            for (int t = 0; t < TC_STAGES && t < ntiles; ++t) load(t);
            bool ok = true;
            for (int t = 0; t < ntiles && ok; ++t) {
                const int s = t % TC_STAGES, b = t & 1;
                ok = tc_mbar_wait(&full[s], (uint32_t)((t / TC_STAGES) & 1), s_fail);                    // the tile has landed
                if (ok && t >= 2) ok = tc_mbar_wait(&tempty[b], (uint32_t)(((t >> 1) - 1) & 1), s_fail);  // the epilogue has drained this accumulator
                if (!ok) break;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                uint32_t acc = 0;
                const uint32_t sbh = tc_smem_u32(sB + s * TC_TILE_BYTES), sbl = sbh + TC_N * TC_K * 4;
#pragma unroll
                for (int term = 0; term < 3; ++term) {                            // A_hi B_hi + A_hi B_lo + A_lo B_hi
                    const uint32_t sa = tc_smem_u32(term == 2 ? sAl : sAh), sb = term == 1 ? sbl : sbh;
#pragma unroll
                    for (int ks = 0; ks < TC_K / 8; ++ks) {                       // UMMA_K = 8 for tf32: two core matrices per instruction
                        const uint64_t ad = tc_make_desc(sa + ks * 2 * (TC_M * 16), TC_M * 16, 128);
                        const uint64_t bd = tc_make_desc(sb + ks * 2 * (TC_N * 16), TC_N * 16, 128);
                        asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(
                                         tbase + (uint32_t)(b * TC_N)),
                                     "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
                                     : "memory");
                        acc = 1;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&empty[s])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&tfull[b])) : "memory");
                // refill the slot of the PREVIOUS tile (its MMAs were issued one iteration ago and have, as a rule, retired)
                if (t >= 1 && t - 1 + TC_STAGES < ntiles) {
                    const int sp = (t - 1) % TC_STAGES;
                    ok = tc_mbar_wait(&empty[sp], (uint32_t)(((t - 1) / TC_STAGES) & 1), s_fail);
                    if (ok) load(t - 1 + TC_STAGES);
                }
            }
        }
    } else {
        // ===================== epilogue warps: thread = observation = TMEM lane =====================
        float thr = 1e29f;
        bool ok = true;
        auto insert = [&](float c, int id) {                                      // rows ascend while scanning: an equal value stays behind the earlier row
            int k = TC_KEEP - 1;
#pragma unroll 1
            for (; k > 0 && c < my_c[k - 1]; --k) { my_c[k] = my_c[k - 1]; my_i[k] = my_i[k - 1]; }
            my_c[k] = c;
            my_i[k] = id;
            thr = my_c[TC_KEEP - 1];
        };
        for (int t = 0; t < ntiles && ok; ++t) {
            const int b = t & 1;
            ok = tc_mbar_wait(&tfull[b], (uint32_t)((t >> 1) & 1), s_fail);
            if (!ok) break;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int base = (int)((t0 + t) * TC_N);                              // shard-relative row of column 0 (n_lut < 2^31)
            auto ld32 = [&](uint32_t (&rr)[32], int c0) {
                const uint32_t taddr = tbase + ((uint32_t)(32 * warp) << 16) + (uint32_t)(b * TC_N + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
                    "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(rr[0]), "=r"(rr[1]), "=r"(rr[2]), "=r"(rr[3]), "=r"(rr[4]), "=r"(rr[5]), "=r"(rr[6]), "=r"(rr[7]), "=r"(rr[8]), "=r"(rr[9]),
                      "=r"(rr[10]), "=r"(rr[11]), "=r"(rr[12]), "=r"(rr[13]), "=r"(rr[14]), "=r"(rr[15]), "=r"(rr[16]), "=r"(rr[17]), "=r"(rr[18]), "=r"(rr[19]),
                      "=r"(rr[20]), "=r"(rr[21]), "=r"(rr[22]), "=r"(rr[23]), "=r"(rr[24]), "=r"(rr[25]), "=r"(rr[26]), "=r"(rr[27]), "=r"(rr[28]), "=r"(rr[29]),
                      "=r"(rr[30]), "=r"(rr[31])
                    : "r"(taddr));
            };
            auto scan32 = [&](const uint32_t (&rr)[32], int c0) {
                // minima of the four groups of eight columns, then of the 32: a hit (rare after the first tiles) only rescans its group
                float g[4];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    g[gi] = __uint_as_float(rr[8 * gi]);
#pragma unroll
                    for (int j = 1; j < 8; ++j) g[gi] = fminf(g[gi], __uint_as_float(rr[8 * gi + j]));
                }
                if (fminf(fminf(g[0], g[1]), fminf(g[2], g[3])) < thr) {
#pragma unroll
                    for (int gi = 0; gi < 4; ++gi)
                        if (g[gi] < thr) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float cj = __uint_as_float(rr[8 * gi + j]);
                                if (cj < thr) insert(cj, base + c0 + 8 * gi + j);
                            }
                        }
                }
            };
#pragma unroll 1
            for (int c0 = 0; c0 < TC_N; c0 += 64) {                               // two 32-column loads in flight per wait
                uint32_t ra[32], rb[32];
                ld32(ra, c0);
                ld32(rb, c0 + 32);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                scan32(ra, c0);
                scan32(rb, c0 + 32);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            tc_mbar_arrive(&tempty[b]);                                           // this thread is done with accumulator b
        }
        if (q < a.n_obs) {
            const bool failed = *s_fail != 0;
            float* pc = a.cand_cost + ((size_t)q * a.n_chunks + blockIdx.y) * TC_KEEP;
            int* pi = a.cand_idx + ((size_t)q * a.n_chunks + blockIdx.y) * TC_KEEP;
#pragma unroll
            for (int k = 0; k < TC_KEEP; ++k) { pc[k] = my_c[k]; pi[k] = failed ? -1 : my_i[k]; }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 0 && *s_fail) atomicExch(a.status, 1);
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(2 * TC_N) : "memory");
}

// exact float32 contract cost of one (observation, row) pair: pre-scaled features, ascending d, one fmaf per feature (k_invert)
__device__ __forceinline__ float tc_exact_cost(const float* __restrict__ o /* scaled */, const float* __restrict__ lrow, const float* __restrict__ scale, int n_feat) {
    float acc = 0.0f;
    for (int d = 0; d < n_feat; ++d) {
        const float df = o[d] - __ldg(lrow + d) * scale[d];
        acc = fmaf(df, df, acc);
    }
    return acc;
}

// eps(o, l) of the header for a row of squared norm n2 (m = ||phi(o)||, sc = bound on sqrt(cost) of the rows that matter)
__device__ __forceinline__ float tc_eps(float kappa, float m, float n2, float sc) {
    const float n = sqrtf(n2) * 1.000001f;
    return 1.0001f * (kappa * (2.0f * m * n + n * n) + TC_KN * n * n + TC_KC * (m + n) * sc);
}

// candidates -> exact arg-min (one thread per observation)
__global__ void __launch_bounds__(128) k_invert_tc_rerank(const InvertTcArgs a, long long index_offset, long long* __restrict__ best_idx,
                                                          float* __restrict__ best_cost) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= a.n_obs) return;
    float sc[16], o[16], xo[16], ce[16];
    float on2 = 0.0f;
    for (int d = 0; d < 16; ++d) {
        sc[d] = d < a.n_feat ? a.scale[d] : 0.0f;
        ce[d] = d < a.n_feat ? a.center[d] : 0.0f;
        o[d] = d < a.n_feat ? a.obs[q * a.pitch_obs + d] * sc[d] : 0.0f;
        xo[d] = d < a.n_feat ? o[d] - ce[d] : 0.0f;                                // phi(o) exactly as k_invert_tc forms it
        on2 = fmaf(xo[d], xo[d], on2);
    }
    const float m = sqrtf(on2) * 1.000001f;
    const float* pc = a.cand_cost + (size_t)q * a.n_chunks * TC_KEEP;
    const int* pi = a.cand_idx + (size_t)q * a.n_chunks * TC_KEEP;
    const int ncand = a.n_chunks * TC_KEEP;
    float cmin = __int_as_float(0x7f800000);
    int imin = -1;
    for (int i = 0; i < ncand; ++i)
        if (pi[i] >= 0 && pc[i] < cmin) { cmin = pc[i]; imin = pi[i]; }
    bool redo = imin < 0;                 // nothing kept: the CTA timed out, or no row of the shard has a finite cost (the scan settles which)
    float best = __int_as_float(0x7f800000);
    int bidx = -1;
    if (!redo) {
        // monitor: real bracket (double) of every kept candidate against its c^; also yields ||phi(l^)||^2
        float n2hat = 0.0f, worst = 0.0f;
        int viol = 0;
        for (int i = 0; i < ncand; ++i) {
            const int id = pi[i];
            if (id < 0 || !(pc[i] < 1e29f)) continue;
            const float* lrow = a.lut + (long long)id * a.pitch_lut;
            double n2 = 0.0, dot = 0.0, sabs = 0.0;
            for (int d = 0; d < a.n_feat; ++d) {
                const float xl = __ldg(lrow + d) * sc[d] - ce[d];
                n2 = fma((double)xl, (double)xl, n2);
                dot = fma((double)xo[d], (double)xl, dot);
                sabs += fabs((double)xo[d] * (double)xl);
            }
            if (!(n2 < 1e29)) continue;
            const double err = fabs((double)pc[i] - (n2 - 2.0 * dot));
            const double S = 2.0 * sabs + n2;
            if (S > 0.0) worst = fmaxf(worst, (float)(err / S));
            if (err > (double)tc_eps(a.kappa, m, (float)n2, 0.0f)) ++viol;
            if (id == imin) n2hat = (float)n2;
        }
        atomicMax(reinterpret_cast<unsigned*>(a.status) + 2, __float_as_uint(worst));
        if (viol) { atomicAdd(a.status + 3, viol); redo = true; }
        const float C1 = tc_exact_cost(o, a.lut + (long long)imin * a.pitch_lut, sc, a.n_feat);
        if (!(C1 < 1e29f)) redo = true;
        const float sC = sqrtf(C1) * 1.0001f;
        const float nmax = m + sC;
        const float tau = tc_eps(a.kappa, m, n2hat, sC) + tc_eps(a.kappa, m, nmax * nmax, sC) + 2.02f * TC_GAMMA * C1;
        const float lim = cmin + tau;
        for (int c = 0; c < a.n_chunks && !redo; ++c) {
            const float* cc = pc + c * TC_KEEP;
            const int* ci = pi + c * TC_KEEP;
            if (ci[TC_KEEP - 1] >= 0 && cc[TC_KEEP - 1] <= lim) redo = true;      // the list is full of candidates: it may have dropped one
            for (int k = 0; k < TC_KEEP; ++k) {
                const int id = ci[k];
                if (id < 0 || !(cc[k] <= lim)) continue;
                const float ex = tc_exact_cost(o, a.lut + (long long)id * a.pitch_lut, sc, a.n_feat);
                if (ex < best || (ex == best && id < bidx)) { best = ex; bidx = id; }
            }
        }
    }
    if (redo) {
        const int slot = atomicAdd(a.status + 1, 1);
        a.flags[slot] = (int)q;
        a.fix_best[q] = ~0ull;
    }
    best_idx[q] = bidx < 0 ? -1 : (long long)bidx + index_offset;
    best_cost[q] = bidx < 0 ? __int_as_float(0x7f800000) : best;
}

// the fix-up list: plain exact scan.  grid.x slices the shard, grid.y walks the list in groups of TC_FIX_GROUP observations, so
// one pass over a slice serves the whole group; the per-slice results fold with a 64-bit atomicMin on (cost bits, row) -- costs are
// non-negative floats, whose bit patterns order like the values, and the low word breaks ties towards the lowest row.
__global__ void __launch_bounds__(256) k_invert_tc_fix(const InvertTcArgs a) {
    const int nflag = a.status[1];
    __shared__ float s_o[TC_FIX_GROUP][16];
    __shared__ float s_sc[16];
    __shared__ unsigned long long s_red[8][TC_FIX_GROUP];
    for (int g0 = blockIdx.y * TC_FIX_GROUP; g0 < nflag; g0 += gridDim.y * TC_FIX_GROUP) {
        __syncthreads();
        if (threadIdx.x < 16) s_sc[threadIdx.x] = threadIdx.x < a.n_feat ? a.scale[threadIdx.x] : 0.0f;
        if (threadIdx.x < TC_FIX_GROUP * 16) {
            const int j = threadIdx.x >> 4, d = threadIdx.x & 15;
            const int qi = g0 + j < nflag ? a.flags[g0 + j] : a.flags[g0];
            s_o[j][d] = d < a.n_feat ? a.obs[(long long)qi * a.pitch_obs + d] * a.scale[d] : 0.0f;
        }
        __syncthreads();
        unsigned long long bestp[TC_FIX_GROUP];
#pragma unroll
        for (int j = 0; j < TC_FIX_GROUP; ++j) bestp[j] = ~0ull;
        for (long long r = (long long)blockIdx.x * 256 + threadIdx.x; r < a.n_lut; r += (long long)gridDim.x * 256) {
            const float* lrow = a.lut + r * a.pitch_lut;
            float acc[TC_FIX_GROUP];
#pragma unroll
            for (int j = 0; j < TC_FIX_GROUP; ++j) acc[j] = 0.0f;
            for (int d = 0; d < a.n_feat; ++d) {
                const float ls = __ldg(lrow + d) * s_sc[d];
#pragma unroll
                for (int j = 0; j < TC_FIX_GROUP; ++j) {
                    const float df = s_o[j][d] - ls;
                    acc[j] = fmaf(df, df, acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < TC_FIX_GROUP; ++j)
                if (acc[j] < __int_as_float(0x7f800000)) {                           // finite (false for NaN)
                    const unsigned long long pk = ((unsigned long long)__float_as_uint(acc[j]) << 32) | (unsigned long long)(unsigned)r;
                    bestp[j] = pk < bestp[j] ? pk : bestp[j];
                }
        }
#pragma unroll
        for (int j = 0; j < TC_FIX_GROUP; ++j) {
            unsigned long long v = bestp[j];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const unsigned long long w = __shfl_xor_sync(0xffffffffu, v, off);
                v = w < v ? w : v;
            }
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][j] = v;
        }
        __syncthreads();
        if (threadIdx.x < TC_FIX_GROUP && g0 + threadIdx.x < nflag) {
            unsigned long long v = s_red[0][threadIdx.x];
            for (int w = 1; w < 8; ++w) v = s_red[w][threadIdx.x] < v ? s_red[w][threadIdx.x] : v;
            if (v != ~0ull) atomicMin(a.fix_best + a.flags[g0 + threadIdx.x], v);
        }
    }
}

__global__ void __launch_bounds__(256) k_invert_tc_fix_out(const InvertTcArgs a, long long index_offset, long long* __restrict__ best_idx,
                                                           float* __restrict__ best_cost) {
    const int nflag = a.status[1];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nflag; i += gridDim.x * blockDim.x) {
        const int q = a.flags[i];
        const unsigned long long v = a.fix_best[q];
        best_idx[q] = v == ~0ull ? -1 : (long long)(unsigned)(v & 0xffffffffull) + index_offset;
        best_cost[q] = v == ~0ull ? __int_as_float(0x7f800000) : __uint_as_float((unsigned)(v >> 32));
    }
}

}  // namespace pb
