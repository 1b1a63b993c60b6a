// LUT inversion: for every observed feature vector (e.g. the 15 L8/ASTER window means of a pixel) find the LUT entry with the
// smallest weighted squared distance.  This is the consumer of the LUT the band kernel produces (BASELINE.json north star;
// SURVEY.md 8(f)-4; it has no counterpart in the reference, which stops at the forward model).
//
// cost(q, e) = sum_d ( s_d o_{q,d} - s_d l_{e,d} )^2 ,  s_d = sqrt(w_d)  (features are pre-scaled once, so the inner loop is a
// subtraction and one FMA per feature), accumulated in float32 in ascending d with fma(diff, diff, acc): a fixed, documented
// order that a host check can reproduce.  Ties go to the lowest LUT index (strict '<' while scanning indices upwards).
//
// Mapping: a CTA owns 256 threads x 4 observations (features in registers) and one chunk of the LUT, which it
// streams through shared memory in tiles; every LUT entry is read as 128-bit shared-memory broadcasts and reused by the four
// observations of each thread (4 x DP FMAs per DP / 4 LDS.128).  Per-(observation, chunk) minima go to a small partial buffer;
// k_invert_reduce folds the chunks in index order.  Bound: the FP32 lanes (packed FADD2 / FFMA2 on pairs of observations), not HBM: the LUT is re-read
// once per 1024 observations and stays in the 126 MB L2 for LUT shards up to ~2e6 entries x 16 features.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pb {

constexpr int INV_THREADS = 256;
constexpr int INV_TILE_FLOATS = 4096;              // floats per shared-memory tile buffer (16 KB, two buffers)
// DP = padded feature count (8, 16 or 32); OPT = observations per thread (4, or 2 when 32 features fill the registers)
template <int DP> struct InvCfg {
    static constexpr int OPT = DP <= 16 ? 4 : 2;
    static constexpr int OBS_PER_CTA = INV_THREADS * OPT;
    static constexpr int TILE = INV_TILE_FLOATS / DP;        // LUT entries per tile
};

struct InvertArgs {
    const float* lut;        // [n_lut][pitch_lut]
    long long n_lut, pitch_lut;
    const float* obs;        // [n_obs][pitch_obs]
    long long n_obs, pitch_obs;
    const float* scale;      // [DP] sqrt(weights), zero beyond n_feat (device)
    int n_feat;
    long long chunk;         // LUT entries per chunk (multiple of the tile size)
    int n_chunks;
    float* part_cost;        // [n_obs][n_chunks]
    long long* part_idx;     // [n_obs][n_chunks]
};

template <int DP>
__global__ void __launch_bounds__(INV_THREADS) k_invert(const InvertArgs a) {
    constexpr int INV_OPT = InvCfg<DP>::OPT, INV_OBS_PER_CTA = InvCfg<DP>::OBS_PER_CTA, INV_TILE = InvCfg<DP>::TILE;
    __shared__ __align__(16) float tile[2][INV_TILE_FLOATS];
    __shared__ float s_scale[DP];
    const int tid = threadIdx.x;
    if (tid < DP) s_scale[tid] = a.scale[tid];
    __syncthreads();

    // this thread's observations, pre-scaled (invalid rows replicate row 0 and are not written back), held as PAIRS of
    // observations per feature: the distance update runs in packed FP32 (FADD2 with the LUT value as a broadcast operand,
    // FFMA2), i.e. half the issue slots of the scalar form for the same arithmetic, bit for bit
    static_assert(INV_OPT % 2 == 0, "observations are processed in pairs");
    float2 o2[INV_OPT / 2][DP];
    long long q[INV_OPT];
#pragma unroll
    for (int j = 0; j < INV_OPT; ++j) {
        q[j] = (long long)blockIdx.x * INV_OBS_PER_CTA + j * INV_THREADS + tid;
        const long long row = q[j] < a.n_obs ? q[j] : 0;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            const float v = d < a.n_feat ? a.obs[row * a.pitch_obs + d] * s_scale[d] : 0.0f;
            if (j & 1) o2[j / 2][d].y = v; else o2[j / 2][d].x = v;
        }
    }
    float best[INV_OPT];
    long long bidx[INV_OPT];
#pragma unroll
    for (int j = 0; j < INV_OPT; ++j) { best[j] = __int_as_float(0x7f800000); bidx[j] = -1; }

    const long long e0 = (long long)blockIdx.y * a.chunk;
    const long long e1 = min(e0 + a.chunk, a.n_lut);
    const int ntiles = (int)((e1 - e0 + INV_TILE - 1) / INV_TILE);

    // cooperative tile load: element i of the tile = (entry i / DP, feature i % DP), scaled; coalesced when pitch_lut == n_feat
    auto load_tile = [&](int t, int buf) {
        const long long base = e0 + (long long)t * INV_TILE;
        for (int i = tid; i < INV_TILE * DP; i += INV_THREADS) {
            const int e = i / DP, d = i - e * DP;
            const long long row = base + e;
            float v = 0.0f;
            if (d < a.n_feat && row < e1) v = __ldg(a.lut + row * a.pitch_lut + d) * s_scale[d];
            tile[buf][i] = v;
        }
    };
    if (ntiles > 0) load_tile(0, 0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_tile(t + 1, buf ^ 1);          // the other buffer was released by the barrier below
        const long long base = e0 + (long long)t * INV_TILE;
        const int n = (int)min((long long)INV_TILE, e1 - base);
        const float4* tp = reinterpret_cast<const float4*>(tile[buf]);
#pragma unroll 2
        for (int e = 0; e < n; ++e) {
            float2 acc[INV_OPT / 2];
#pragma unroll
            for (int j = 0; j < INV_OPT / 2; ++j) acc[j] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int d4 = 0; d4 < DP / 4; ++d4) {
                const float4 l = tp[e * (DP / 4) + d4];          // broadcast: all lanes read the same 16 bytes
#pragma unroll
                for (int j = 0; j < INV_OPT / 2; ++j) {
                    float2 df = __fadd2_rn(o2[j][4 * d4 + 0], make_float2(-l.x, -l.x)); acc[j] = __ffma2_rn(df, df, acc[j]);
                    df = __fadd2_rn(o2[j][4 * d4 + 1], make_float2(-l.y, -l.y)); acc[j] = __ffma2_rn(df, df, acc[j]);
                    df = __fadd2_rn(o2[j][4 * d4 + 2], make_float2(-l.z, -l.z)); acc[j] = __ffma2_rn(df, df, acc[j]);
                    df = __fadd2_rn(o2[j][4 * d4 + 3], make_float2(-l.w, -l.w)); acc[j] = __ffma2_rn(df, df, acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < INV_OPT / 2; ++j) {
                if (acc[j].x < best[2 * j]) { best[2 * j] = acc[j].x; bidx[2 * j] = base + e; }
                if (acc[j].y < best[2 * j + 1]) { best[2 * j + 1] = acc[j].y; bidx[2 * j + 1] = base + e; }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < INV_OPT; ++j)
        if (q[j] < a.n_obs) {
            a.part_cost[q[j] * a.n_chunks + blockIdx.y] = best[j];
            a.part_idx[q[j] * a.n_chunks + blockIdx.y] = bidx[j];
        }
}

// fold the chunk minima of one observation in chunk (= index) order; strict '<' keeps the lowest index among equal costs
__global__ void k_invert_reduce(long long n_obs, int n_chunks, const float* __restrict__ part_cost, const long long* __restrict__ part_idx,
                                long long index_offset, long long* __restrict__ best_idx, float* __restrict__ best_cost) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_obs) return;
    float best = __int_as_float(0x7f800000);
    long long idx = -1;
    for (int c = 0; c < n_chunks; ++c) {
        const float v = part_cost[q * n_chunks + c];
        if (v < best) { best = v; idx = part_idx[q * n_chunks + c]; }
    }
    best_idx[q] = idx < 0 ? -1 : idx + index_offset;
    best_cost[q] = best;
}

// ---------------------------------------------------------------------------------------------------------------
// k best entries per observation (k <= INV_KMAX): the usual LUT inversion takes the mean / median of the parameters of the k
// closest entries.  One observation per thread (features + a sorted list of k (cost, index) pairs in registers); the same LUT
// streaming as k_invert.  An entry enters the list only if it beats the current k-th cost, which after the first few tiles is
// rare, so the inner loop is the distance evaluation plus one compare.  Order: ascending cost, equal costs by ascending index.
// ---------------------------------------------------------------------------------------------------------------
constexpr int INV_KMAX = 16;

__device__ __forceinline__ bool inv_before(float c, long long i, float c2, long long i2) { return c < c2 || (c == c2 && i < i2); }

template <int DP>
__global__ void __launch_bounds__(INV_THREADS) k_invert_topk(const InvertArgs a, int kk) {
    constexpr int TILE = InvCfg<DP>::TILE;
    __shared__ __align__(16) float tile[2][INV_TILE_FLOATS];
    __shared__ float s_scale[DP];
    const int tid = threadIdx.x;
    if (tid < DP) s_scale[tid] = a.scale[tid];
    __syncthreads();
    const long long q = (long long)blockIdx.x * INV_THREADS + tid;
    const long long row = q < a.n_obs ? q : 0;
    float o[DP];
#pragma unroll
    for (int d = 0; d < DP; ++d) o[d] = d < a.n_feat ? a.obs[row * a.pitch_obs + d] * s_scale[d] : 0.0f;
    float bc[INV_KMAX];
    long long bi[INV_KMAX];
#pragma unroll
    for (int j = 0; j < INV_KMAX; ++j) { bc[j] = __int_as_float(0x7f800000); bi[j] = -1; }
    float worst = __int_as_float(0x7f800000);      // cost of the current k-th entry (bc[kk - 1], kept in its own register)

    const long long e0 = (long long)blockIdx.y * a.chunk;
    const long long e1 = min(e0 + a.chunk, a.n_lut);
    const int ntiles = (int)((e1 - e0 + TILE - 1) / TILE);
    auto load_tile = [&](int t, int buf) {
        const long long base = e0 + (long long)t * TILE;
        for (int i = tid; i < TILE * DP; i += INV_THREADS) {
            const int e = i / DP, d = i - e * DP;
            const long long r = base + e;
            float v = 0.0f;
            if (d < a.n_feat && r < e1) v = __ldg(a.lut + r * a.pitch_lut + d) * s_scale[d];
            tile[buf][i] = v;
        }
    };
    if (ntiles > 0) load_tile(0, 0);
    __syncthreads();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) load_tile(t + 1, buf ^ 1);
        const long long base = e0 + (long long)t * TILE;
        const int n = (int)min((long long)TILE, e1 - base);
        const float4* tp = reinterpret_cast<const float4*>(tile[buf]);
        for (int e = 0; e < n; ++e) {
            float acc = 0.0f;
#pragma unroll
            for (int d4 = 0; d4 < DP / 4; ++d4) {
                const float4 l = tp[e * (DP / 4) + d4];
                float df = o[4 * d4 + 0] - l.x; acc = fmaf(df, df, acc);
                df = o[4 * d4 + 1] - l.y; acc = fmaf(df, df, acc);
                df = o[4 * d4 + 2] - l.z; acc = fmaf(df, df, acc);
                df = o[4 * d4 + 3] - l.w; acc = fmaf(df, df, acc);
            }
            if (acc < worst) {
                // sorted insertion (indices ascend while scanning, so among equal costs the earlier entry stays in front)
                float c = acc;
                long long id = base + e;
#pragma unroll
                for (int j = 0; j < INV_KMAX; ++j) {
                    const bool sw = j < kk && c < bc[j];
                    const float tc = bc[j];
                    const long long ti = bi[j];
                    bc[j] = sw ? c : tc;
                    bi[j] = sw ? id : ti;
                    c = sw ? tc : c;
                    id = sw ? ti : id;
                    if (j == kk - 1) worst = bc[j];
                }
            }
        }
        __syncthreads();
    }
    if (q < a.n_obs) {
        float* pc = a.part_cost + (q * a.n_chunks + blockIdx.y) * kk;
        long long* pi = a.part_idx + (q * a.n_chunks + blockIdx.y) * kk;
#pragma unroll
        for (int j = 0; j < INV_KMAX; ++j)
            if (j < kk) { pc[j] = bc[j]; pi[j] = bi[j]; }
    }
}

// merge the chunk lists of one observation (each sorted): k rounds of "smallest head", ties to the lower index
__global__ void k_invert_topk_reduce(long long n_obs, int n_chunks, int kk, const float* __restrict__ part_cost,
                                     const long long* __restrict__ part_idx, long long index_offset, long long* __restrict__ best_idx,
                                     float* __restrict__ best_cost) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_obs) return;
    const float* pc = part_cost + q * n_chunks * kk;
    const long long* pi = part_idx + q * n_chunks * kk;
    float lastc = -1.0f;
    long long lasti = -1;
    for (int j = 0; j < kk; ++j) {
        // smallest (cost, index) strictly after the previous pick: lists are short, a full scan per pick is cheap enough
        float bc = __int_as_float(0x7f800000);
        long long bi = -1;
        for (int i = 0; i < n_chunks * kk; ++i) {
            const float c = pc[i];
            const long long id = pi[i];
            if (id < 0) continue;
            const bool after = c > lastc || (c == lastc && id > lasti);
            if (after && (bi < 0 || inv_before(c, id, bc, bi))) { bc = c; bi = id; }
        }
        best_idx[q * kk + j] = bi < 0 ? -1 : bi + index_offset;
        best_cost[q * kk + j] = bc;
        if (bi >= 0) { lastc = bc; lasti = bi; }
        else { lastc = __int_as_float(0x7f800000); lasti = 0x7fffffffffffffffLL; }
    }
}

}  // namespace pb
