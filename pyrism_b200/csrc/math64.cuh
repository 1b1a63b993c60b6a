// fp64 primitives for the PROSAIL kernels (sm_100a).
//
// The PROSPECT->4SAIL chain is bound by the FP64 pipe (64 DFMA lanes per SM), not by HBM, so every
// primitive here is written to minimise FP64-pipe instructions: the MUFU seeds (RCP64H / RSQ64H),
// exponent/mantissa bit manipulation and table addressing run on the XU / ALU / LSU pipes, which
// issue in the shadow of the 2-cycle DFMA slots.  Accuracy target of each primitive: <= 4e-16
// relative (checked on the GPU against numpy/mpmath by tests/test_gpu_math.py through
// prosail_test_math()), which keeps the end-to-end results within 1e-12 of the reference where the
// parity bar is 1e-9.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "math_tables.inc"

#define PB_MAGIC 6755399441055744.0  /* 1.5 * 2^52: round-to-nearest-integer by addition */

namespace pb {

// Polynomial / range-reduction constants live in constant memory so that DFMA takes them straight from the
// constant bank (operand c[3][..]); as literals they would each cost two UMOV issue slots per use.
// (Values whose low 32 bits are zero -- 0.5, 1.0, 64.0, the rounding magic -- are encodable as immediates.)
enum { M_L2E64 = 0, M_LN2H, M_LN2L, M_E5, M_E4, M_E3, M_L5, M_L4, M_L3, M_L2, M_L1, M_EBIAS, M_P5, M_P4, M_P3, M_P2, M_P1, M_COUNT };
__constant__ double c_m[M_COUNT] = {
    92.332482616893656877,       // 64 log2(e)
    -0x1.62e42fefa0000p-7,       // -(ln2/64) high part (low 17 bits zero)
    -0x1.cf79abc9e3b3ap-46,      // -(ln2/64) low part
    0x1.1111111111111p-7,        // 1/120
    0x1.5555555555555p-5,        // 1/24
    0x1.5555555555555p-3,        // 1/6
    0x1.2776c50ef9bfep-2,        // 1/(5 ln2)
    -0x1.71547652b82fep-2,       // -1/(4 ln2)
    0x1.ec709dc3a03fdp-2,        // 1/(3 ln2)
    -0x1.71547652b82fep-1,       // -1/(2 ln2)
    0x1.71547652b82fep+0,        // 1/ln2
    4503601774854144.0,          // 2^52 + 2^31
    0x1.5d87fe78a6731p-10,       // ln2^5/120
    0x1.3b2ab6fba4e77p-7,        // ln2^4/24
    0x1.c6b08d704a0c0p-5,        // ln2^3/6
    0x1.ebfbdff82c58fp-3,        // ln2^2/2
    0x1.62e42fefa39efp-1,        // ln2
};

// ---- shared-memory image of the lookup tables (copied once per CTA) ----
struct alignas(16) MathTables {
    double tau[TAU_NROWS * TAU_ROW];   // piecewise polynomials of tau(k), see tools/gen_math_tables.py
    double exp2t[64];                  // 2^(j/64)
    double logt[256];                  // pairs (1/c_i, log2 c_i), c_i = 1 + (i+.5)/128
};

// 1/x: MUFU.RCP64H seed (upper-32-bit approximation, ~2^-20) + one cubic step (3 DFMA) -> ~1 ulp.
__device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double p = fma(e, e, e);
    y = fma(y, p, y);
#ifdef PB_RCP_EXTRA_STEP
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
#endif
    return y;
}

// sqrt(x), x > 0: MUFU.RSQ64H seed + two coupled Goldschmidt steps (7 FP64 ops).  x <= 0 -> 0.
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    return x > 0.0 ? g : 0.0;
}

// exp(x) for x <= 0 (clamped at -700): n = rint(64 x log2 e); exp = 2^(n>>6) * T[n&63] * P5(r).  10 FP64 ops.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ exp2t) {
    x = fmax(x, -700.0);
    const double t = fma(x, c_m[M_L2E64], PB_MAGIC);
    const int n = __double2loint(t);
    const double nd = t - PB_MAGIC;
    double r = fma(nd, c_m[M_LN2H], x);
    r = fma(nd, c_m[M_LN2L], r);
    double p = fma(r, c_m[M_E5], c_m[M_E4]);
    p = fma(r, p, c_m[M_E3]);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    const double v = exp2t[n & 63] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// log2(b) for normal b > 0: b = 2^E m, m in [1,2); i = top 7 mantissa bits; r = m/c_i - 1, |r| <= 2^-8;
// log2 b = E + log2 c_i + log2(1+r) with a degree-5 series.  8 FP64 ops + one 16-byte table load.
__device__ __forceinline__ double log2_pos(double b, const double* __restrict__ logt) {
    const int hi = __double2hiint(b);
    const int i = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(b));
    const double2 tc = *reinterpret_cast<const double2*>(logt + 2 * i);
    const double r = fma(m, tc.x, -1.0);
    // E as a double without I2F: 2^52 + 2^31 + (E biased by 2^31), minus the constant.
    const int E = (hi >> 20) - 1023;
    const double Ed = __hiloint2double(0x43300000, E ^ 0x80000000) - c_m[M_EBIAS];
    double q = fma(r, c_m[M_L5], c_m[M_L4]);
    q = fma(r, q, c_m[M_L3]);
    q = fma(r, q, c_m[M_L2]);
    q = fma(r, q, c_m[M_L1]);
    return fma(r, q, Ed + tc.y);
}

// 2^y for |y| < 1000: n = rint(64 y); r = y - n/64 (exact); 2^y = 2^(n>>6) T[n&63] P5(r ln2).  9 FP64 ops.
__device__ __forceinline__ double exp2_any(double y, const double* __restrict__ exp2t) {
    const double t = fma(y, 64.0, PB_MAGIC);
    const int n = __double2loint(t);
    const double nd = t - PB_MAGIC;
    const double r = fma(nd, -0.015625, y);
    double p = fma(r, c_m[M_P5], c_m[M_P4]);
    p = fma(r, p, c_m[M_P3]);
    p = fma(r, p, c_m[M_P2]);
    p = fma(r, p, c_m[M_P1]);
    p = fma(r, p, 1.0);
    const double v = exp2t[n & 63] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> 6) << 20), __double2loint(v));
}

// tau(k) = (1-k) e^-k + k^2 E1(k)   (reference models.py:953-957; tau = 1 for k <= 0).
// Interval index, scale S (a power of two) and the row address come from integer ops on the bits of k;
// the FP64 pipe sees 1 FMA (local variable) + TAU_DEG FMAs (Horner).
__device__ __forceinline__ double tau_plate(double k, const double* __restrict__ tab) {
    const int hi = __double2hiint(k);
    const int E = (hi >> 20) - 1023;
    if (__builtin_expect(E >= 6 || E < TAU_EMIN, 0)) {
        if (!(k > 0.0)) return 1.0;                              // k <= 0 (or NaN): reference leaves tau = 1
        if (E < TAU_EMIN) return fma(-2.0, k, 1.0);              // k < 2^-30: |k^2 ln k| < 2e-17
        // k >= 64: tau < 1e-29; asymptotic series of k e^k E1(k) (relative accuracy ~1e-8, absolute < 1e-37)
        const double ik = 1.0 / k;
        double s = fma(ik, -5040.0, 720.0);
        s = fma(ik, s, -120.0);
        s = fma(ik, s, 24.0);
        s = fma(ik, s, -6.0);
        s = fma(ik, s, 2.0);
        return exp(-k) * s * ik;
    }
    const bool lowk = E <= 0;
    const int sb = lowk ? 2 : E + 1;
    const int j = (hi & 0x000FFFFF) >> (20 - sb);
    const int idx = lowk ? ((E - TAU_EMIN) << 2) + j : TAU_NA - 4 + (2 << E) + j;
    const int sexp = lowk ? 3 - E : 2;
    const double S = __hiloint2double((1023 + sexp) << 20, 0);
    const double2* row = reinterpret_cast<const double2*>(tab + idx * TAU_ROW);
    double2 c01 = row[0];                                          // O, c0
    const double u = fma(k, S, -c01.x);
    double acc;
    {
        static_assert(TAU_DEG % 2 == 0, "row layout assumes an even degree (TAU_ROW even)");
        double2 c = row[TAU_ROW / 2 - 1];                          // c[D-1], c[D]
        acc = fma(u, c.y, c.x);
#pragma unroll
        for (int q = TAU_ROW / 2 - 2; q >= 1; --q) {
            c = row[q];
            acc = fma(u, acc, c.y);
            acc = fma(u, acc, c.x);
        }
        acc = fma(u, acc, c01.y);
    }
    return acc;
}

// warp-level sum of a double (5 butterfly steps).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pb
