// fp32 primitives for the fp32 mode of the PROSAIL kernels (sm_100a).  Same names as the fp64 set in
// math64.cuh, selected by overload: the band kernel is one template over the real type.
//
// In fp32 the transcendental seeds of the SFU (MUFU.RCP / RSQ / EX2 / LG2, ~1-2 ulp) are already at working
// precision, so no refinement steps are needed; what needs care instead is cancellation (handled in the kernel:
// rinf = sigb/(att+m), J1 through expm1, see prosail_kernels.cuh).  Target of the mode: <= 1e-5 absolute in
// reflectance / transmittance against the float64 reference (north star).
#pragma once
#include "math64.cuh"

namespace pb {

struct alignas(16) MathTablesF {
    float tau[TAU_NROWS * TAU_ROW_F];   // degree-5 piecewise polynomials of g(k) = (1 - tau(k))/k, rows [O, c0..c5, pad]
};

template <typename T> struct Tables;
template <> struct Tables<double> { typedef MathTables type; };
template <> struct Tables<float> { typedef MathTablesF type; };

__device__ __forceinline__ bool is_pos(float x) { return __float_as_int(x) > 0; }
__device__ __forceinline__ bool is_zero(float x) { return (__float_as_int(x) << 1) == 0; }
__device__ __forceinline__ int hi_word(float x) { return __float_as_int(x); }
__device__ __forceinline__ bool below_tiny(float x) { return __float_as_int(x) < 0x03AA2425; }   // x < 1e-36f, negatives too
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_raw(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_pos(float x) { return is_pos(x) ? sqrt_raw(x) : 0.0f; }
__device__ __forceinline__ float rcp_k(float x) { return rcp(x); }      // float twins of the band kernel's reciprocals (math64.cuh)
__device__ __forceinline__ float rcp_b(float x) { return rcp(x); }
__device__ __forceinline__ float exp2_fast(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float log2_fast(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// exp(x), x <= 0
__device__ __forceinline__ float exp_neg(float x, const MathTablesF&) { return exp2_fast(x * 1.4426950408889634f); }
// b^e, b > 0
__device__ __forceinline__ float pow_pos(float b, float e, const MathTablesF&) { return exp2_fast(fminf(log2_fast(b) * e, 120.0f)); }
// expm1(x)/x for moderate |x| (J1 of 4SAIL in fp32): series below 0.25, direct above
__device__ __forceinline__ float expm1_over_x(float x) {
    if (fabsf(x) < 0.25f) {
        float p = fmaf(x, 1.0f / 5040.0f, 1.0f / 720.0f);
        p = fmaf(x, p, 1.0f / 120.0f);
        p = fmaf(x, p, 1.0f / 24.0f);
        p = fmaf(x, p, 1.0f / 6.0f);
        p = fmaf(x, p, 0.5f);
        return fmaf(x, p, 1.0f);
    }
    return (exp2_fast(x * 1.4426950408889634f) - 1.0f) * rcp(x);
}

// g(k) = (1 - tau(k))/k table path in float: same interval scheme as the fp64 tau table (k in [2^-14, 32)), index from
// the float bits.  The kernel forms 1 - tau = k g(k) (relative error ~1e-7) and tau = 1 - k g(k).
__device__ __forceinline__ float tau_fast(float k, const float* __restrict__ tab, bool& rare) {
    const int bits = __float_as_int(k);
    const int raw = (bits >> (23 - TAU_SB)) - ((127 + TAU_EMIN) << TAU_SB);
    const int idx = min(max(raw, 0), TAU_NROWS - 1);
    rare = idx != raw;
    const float mm = __int_as_float((bits & 0x007FFFFF) | ((127 + TAU_SB + 1) << 23));     // mantissa re-scaled to [2^(SB+1), 2^(SB+2))
    const float4* row = reinterpret_cast<const float4*>(tab + idx * TAU_ROW_F);
    const float4 a = row[0], b = row[1];                            // O c0 c1 c2 | c3 c4 c5 pad
    const float u = mm - a.x;
    float acc = fmaf(u, b.z, b.y);
    acc = fmaf(u, acc, b.x);
    acc = fmaf(u, acc, a.w);
    acc = fmaf(u, acc, a.z);
    return fmaf(u, acc, a.y);
}

// ---------------------------------------------------------------------------------------------------------------
// Packed FP32 (sm_100: FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.f32x2): one instruction works on the two bands of a thread.
// Measured on B200 (tools/micro/ffma2.cu, profiles/r1d_micro_ffma2.txt): a packed instruction occupies the FP32 pipes for
// two cycles (same lane throughput as two scalar ones) but ONE issue slot, and the fp32 band kernel is issue-bound; a
// scalar register can be broadcast to both halves as an operand (R.F32) and immediates are splatted, so per-item scalars
// cost nothing extra; negations fold into operand modifiers.  Like DFMA, a three-register-pair FFMA2 takes 3 cycles.
// MUFU seeds (rcp, rsq, ex2, lg2), table look-ups, guards and stores stay scalar per half.
// ---------------------------------------------------------------------------------------------------------------
typedef float2 f2;
using ::fma;      // keep the scalar overloads visible next to the packed ones declared below
__device__ __forceinline__ f2 F2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 F2(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 neg(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ f2 add(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 sub(f2 a, f2 b) { return __fadd2_rn(a, neg(b)); }
__device__ __forceinline__ f2 mul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 add(f2 a, float s) { return __fadd2_rn(a, F2(s)); }
__device__ __forceinline__ f2 mul(f2 a, float s) { return __fmul2_rn(a, F2(s)); }
__device__ __forceinline__ f2 fma(float s, f2 b, f2 c) { return __ffma2_rn(F2(s), b, c); }
__device__ __forceinline__ f2 fma(f2 a, f2 b, float c) { return __ffma2_rn(a, b, F2(c)); }
__device__ __forceinline__ f2 rcp(f2 a) { return F2(rcp(a.x), rcp(a.y)); }
__device__ __forceinline__ f2 sqrt_pos(f2 a) { return F2(sqrt_pos(a.x), sqrt_pos(a.y)); }
__device__ __forceinline__ f2 exp2_fast(f2 a) { return F2(exp2_fast(a.x), exp2_fast(a.y)); }
__device__ __forceinline__ f2 log2_fast(f2 a) { return F2(log2_fast(a.x), log2_fast(a.y)); }
// expm1(x)/x for |x| < 0.25, both halves: packed degree-6 series (truncation 0.25^7/40320 < 2e-9)
__device__ __forceinline__ f2 expm1_over_x_small(f2 x) {
    f2 p = fma(x, F2(1.0f / 5040.0f), 1.0f / 720.0f);
    p = fma(x, p, 1.0f / 120.0f);
    p = fma(x, p, 1.0f / 24.0f);
    p = fma(x, p, 1.0f / 6.0f);
    p = fma(x, p, 0.5f);
    return fma(x, p, 1.0f);
}

}  // namespace pb
