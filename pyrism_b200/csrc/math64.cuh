// fp64 primitives for the PROSAIL kernels (sm_100a).
//
// The PROSPECT->4SAIL chain is bound by the FP64 pipe (64 DFMA lanes per SM), not by HBM, so every
// primitive here is written to minimise FP64-pipe instructions: the MUFU seeds (RCP64H / RSQ64H),
// exponent/mantissa bit manipulation and table addressing run on the XU / ALU / LSU pipes, which
// issue in the shadow of the 2-cycle DFMA slots.  Accuracy: round 1 held every primitive to <= 4e-16 relative (results
// within 1.3e-13 of the reference where the parity bar is 1e-9); round 2 fixes an error budget of 1e-11 per OUTPUT and spends it
// site by site (second-order Newton steps where the error only scales a result, economised polynomials
// of one degree less for log2 / 2^r together with an exact rare path for the leaves whose 1/D would amplify them, float tails
// of the tau polynomial).  Each primitive is checked on the GPU against numpy / mpmath by tests/test_gpu_math.py through
// prosail_test_math(); the end-to-end deviation over 60 000 random parameter sets is 6.4e-12 (tests/test_gpu_sweeps.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "math_tables.inc"

#define PB_MAGIC 6755399441055744.0  /* 1.5 * 2^52: round-to-nearest-integer by addition */

// Round 2n: algebraic reductions of the FP64 instruction count of the fused double instances (each can be switched off for A/B
// runs: make variant NAME=x EXTRA=-DPB_...=0; profiles/r2n_variants.txt)
#ifndef PB_STOKES_Y
#define PB_STOKES_Y 1    /* Stokes stage divided through by a: Y = bm1 (al - r^2) + D (3 FP64 instructions less, no cancellation) */
#endif
#ifndef PB_RINF_P
#define PB_RINF_P 1      /* rinf = (P - m)/(P + m), P = att + sigb = 1 + bf (rho - tau) (2 less) */
#endif
#ifndef PB_HALVES
#define PB_HALVES 1      /* Ap, Am carried doubled; the powers of two are taken out of invden / inv_omr2 on the integer pipe (2 less) */
#endif
#ifndef PB_FSFOLD
#define PB_FSFOLD 1      /* lai sumint folded into Fs, Ft by the prologue (1 less) */
#endif
#ifndef PB_SOILFOLD
#define PB_SOILFOLD 1    /* soil coupling as [A1 A2/dn + (tsstoo - tss too)] rs with A1 = tss + tsd, A2 = too + tdo (4 less) */
#endif
#ifndef PB_PIPELINE
#define PB_PIPELINE 0    /* double LUT instances: tau(kall) of the next item evaluated inside the FP64-dense block of the current one.
                            Measured SLOWER (BRF 26.06 vs 25.18 ms, leaf-only 12.15 vs 11.92, band means 11.77 vs 11.14, four planes
                            31.4 vs 29.2: profiles/r2n_variants_pipeline.txt) -- ptxas does merge the two stages into one basic block,
                            but at the register cap the extra live values cost more than the filled issue slots bring: off */
#endif
#ifndef PB_F32_ALG
#define PB_F32_ALG 1     /* the packed fp32 item loop with the same re-derived stages (Y form, P/Q form, A1 A2 soil coupling): ~18 packed instructions less per item */
#endif
#ifndef PB_EXPS
#define PB_EXPS 1        /* exp(-m lai): the record carries -lai 2048 log2(e), the product m lai is never formed (1 less) */
#endif
#ifndef PB_POWS
#define PB_POWS 1        /* b^(N-1): (N-1) log2 b is never formed, the rounding magic works in units of 2^-11 (1 less) */
#endif

namespace pb {

// Polynomial / range-reduction constants live in constant memory so that DFMA takes them straight from the
// constant bank (operand c[3][..]); as literals they would each cost two UMOV issue slots per use.
// (Values whose low 32 bits are zero -- 0.5, 1.0, 256.0, the rounding magic -- are encodable as immediates.)
enum { M_L2EN = 0, M_NLN2S, M_E3, M_L4, M_L3, M_L2, M_L1, M_EBIAS, M_P3, M_P2, M_P1, M_S3, M_S2, M_S1, M_L2E, M_P1Q, M_COUNT };
static_assert(EXP2_BITS == 11, "the range-reduction constants below are for a 2048-entry 2^(j/2048) table");
__constant__ double c_m[M_COUNT] = {
    2954.6394437405970201,       // 2048 log2(e)
    -0x1.62e42fefa39efp-12,      // -ln2/2048
    0x1.5555555555555p-3,        // 1/6
    -0x1.71547652b82fep-2,       // -1/(4 ln2)
    0x1.ec709dc3a03fdp-2,        // 1/(3 ln2)
    -0x1.71547652b82fep-1,       // -1/(2 ln2)
    0x1.71547652b82fep+0,        // 1/ln2
    4503601774854144.0,          // 2^52 + 2^31
    0x1.c6b08d704a0c0p-5,        // ln2^3/6
    0x1.ebfbdff82c58fp-3,        // ln2^2/2
    0x1.62e42fefa39efp-1,        // ln2
    0x1.c6b08d704a0c0p-38,       // (ln2/2048)^3/6      exp_neg_scaled: the reduced argument is in units of 1/2048 octave
    0x1.ebfbdff82c58fp-25,       // (ln2/2048)^2/2
    0x1.62e42fefa39efp-12,       // ln2/2048
    -0x1.715481dd5be27p-1,       // -(1/2 + h^2/4)/ln2, h = 2^-10: quadratic coefficient of the economised cubic of log2(1 + r), |r| <= h
    0x1.62e43004f3e59p-1,        // ln2 (1 + g^2/8), g = ln2/4096: linear coefficient of the economised quadratic of 2^r, |r| <= 1/4096
};

// Round 2: the band kernel is bound by the FP64 pipe INCLUDING its operand reads (a DFMA with three distinct register
// operands takes three pipe cycles), and the Horner steps of the tau polynomial are exactly such
// instructions (coefficients come from a per-lane table row).  PB_TAU_F32TAIL = 1 (default) evaluates the terms of degree
// >= 4 on the FP32 pipe: with |u| <= 1 and |c4| <= 1.7e-6 over the whole table, float rounding (6e-8 relative) leaves
// <= 2e-13 absolute in tau, where the parity bar is 1e-9.  The two conversions (F2F.F32.F64 of u, F2F.F64.F32 of the tail)
// run on the SFU pipe (8 cycles per warp instruction, measured: profiles/r2a_conv_mix.txt), which this kernel leaves
// 80% idle.  FP64-pipe instructions per tau: 5 instead of 9.
#ifndef PB_TAU_F32TAIL
#define PB_TAU_F32TAIL 1
#endif
struct alignas(16) TauRow {            // 48 bytes: three 128-bit shared-memory loads
    double O, c0, c1, c2;              // u = k S - O;  tau = c0 + u (c1 + u (c2 + u T3(u)))
    float f3, f4, f5, f6;              // T3(u) = f3 + u (f4 + u (f5 + u f6)) in float (|c3| <= 5.1e-6 over the table: <= 3e-13 absolute)
};
static_assert(sizeof(TauRow) == 48 && TAU_DEG == 6, "TauRow layout: degree 6, 16 sub-intervals per binade (tools/gen_math_tables.py)");

// ---- shared-memory image of the lookup tables (copied once per CTA) ----
struct alignas(16) MathTables {
#if PB_TAU_F32TAIL
    TauRow tau[TAU_NROWS];             // piecewise polynomials of tau(k), see tools/gen_math_tables.py (degree 8, tail in float)
#else
    double tau[TAU_NROWS * TAU_ROW];   // piecewise polynomials of tau(k), see tools/gen_math_tables.py
#endif
    double exp2t[1 << EXP2_BITS];      // 2^(j/2048)
    double logt[2 << LOG_BITS];        // pairs (1/c_i, log2 c_i + the constant of the economised cubic), c_i = 1 + (i+.5)/512
};

// Sign/threshold tests on the high word run on the integer pipe instead of costing an FP64-pipe DSETP.
__device__ __forceinline__ bool is_pos(double x) { return __double2hiint(x) > 0; }          // x > 0 (normal)
__device__ __forceinline__ bool is_zero(double x) { return ((__double2hiint(x) << 1) | __double2loint(x)) == 0; }
__device__ __forceinline__ int hi_word(double x) { return __double2hiint(x); }
__device__ __forceinline__ bool below_tiny(double x) { return __double2hiint(x) < 0x38754484; }   // x < 1e-36 (high word of 1e-36), negatives too
// |x| > thr decided on the high words only (integer pipe).  The threshold of 4SAIL's J1 series branch is
// soft: both branches agree to ~1e-13 around it, so a 2^-20-relative fuzz of the switch point is harmless.
__device__ __forceinline__ bool abs_gt_hi(double x, int thr_hi) { return (__double2hiint(x) & 0x7fffffff) > thr_hi; }

// 1/x: MUFU.RCP64H seed (upper-32-bit approximation, ~2^-20) + one cubic step (3 DFMA) -> ~1 ulp.
__device__ __forceinline__ double rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double p = fma(e, e, e);
    return fma(y, p, y);
}

// sqrt(x), x > 0: MUFU.RSQ64H seed y (~2^-20) + one cubic step: with t = x y, e = 1 - t y:
// sqrt(x) = t (1 + e/2 + 3 e^2/8 + O(e^3)).  5 FP64 ops.  x <= 0 -> 0 (integer-pipe select).
// sqrt_raw: no guard -- the caller knows x > 0 (or repairs / discards the lane afterwards).
__device__ __forceinline__ double sqrt_raw(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(e, 0.375, 0.5) * e;
    return fma(t, p, t);
}
__device__ __forceinline__ double sqrt_pos(double x) {
    const double g = sqrt_raw(x);
    return is_pos(x) ? g : 0.0;
}

// ---- Newton steps of second order: the build variant -DPB_QUADRATIC=1 (NOT the default) ----
// The MUFU seeds are good to about 2^-20 (they read the upper 32 bits of the argument), so ONE quadratic step leaves
// e^2 ~ 9.2e-13 relative (measured: tests/test_gpu_math.py) where the cubic steps above reach 1 ulp: 2 FP64-pipe instructions
// per reciprocal instead of 3, 3 per square root instead of 5, and the merged reciprocals of sail_diffuse /
// sail_directional_fast undone (a reciprocal is then cheaper than the multiplications that merge it): 365 instead of 395
// FP64-pipe instructions per pair of bands and item.  Measured on B200 (profiles/r1g_variants.txt): the kernel is only
// 1.4-1.9% faster (the FP64 instruction count is not what binds it), while the deviation from
// the oracle grows from 1.3e-13 to 1.3e-10 (60 000-set sweep) and to 3e-6 in the degenerate all-zero-pigment leaf, where
// 1 - r - t ~ 1e-16 amplifies every relative error by 1e8.  Not worth three digits of parity margin: kept as an experiment.
#ifndef PB_QUADRATIC
#define PB_QUADRATIC 0
#endif
// x * 2^-1 on the integer pipe (x normal and not within one binade of the subnormals)
__device__ __forceinline__ double half_of(double x) { return __hiloint2double(__double2hiint(x) - 0x00100000, __double2loint(x)); }

// max(x, floor) for a positive floor whose high word is `floor_hi`, decided on the high words alone (x < 0, -0.0 included, has a
// negative high word): one VIMNMX instead of a compare and two selects.  The result is >= floor (1 - 2^-20).
__device__ __forceinline__ double floor_hi(double x, int floor_hi_word) { return __hiloint2double(max(__double2hiint(x), floor_hi_word), __double2loint(x)); }
__device__ __forceinline__ double quarter_of(double x) { return __hiloint2double(__double2hiint(x) - 0x00200000, __double2loint(x)); }

__device__ __forceinline__ double rcp_q(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, e, y);
}
// sqrt(x), x > 0 normal: t = x y, e = 1 - t y, sqrt = t (1 + e/2) + O(3 e^2 / 8); the half of t is an exponent decrement
__device__ __forceinline__ double sqrt_q_raw(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    // e/2 = 1/2 - t (y/2): halving the SEED is one integer instruction on its high word (its low word is the zero the MUFU
    // result already carries), halving t would need a second register pair
    const double eh = fma(-t, half_of(y), 0.5);
    return fma(t, eh, t);
}
// ---- round 2: per-site error budget ----
// The parity bar is 1e-9 absolute; round 1 held every primitive to 4e-16.  A second-order Newton step leaves e^2 ~ 9e-13
// RELATIVE.  That is harmless wherever the result only scales an output or enters as the small quantity itself
// ("benign" sites: 1/(1 - x^2), 1/X of the Stokes stage, D = sqrt(...), m, 1/(att + m), 1/den, 1/(1 - rinf^2), 1/dn,
// 1/(k -+ m), 1/(K -+ m)), and it is not where a value close to 1 is later differenced (b = (b t)/t feeds b^(2(N-1)) - 1):
// that one keeps the third-order step.  rcp_b / sqrt_b_raw are the benign-site spellings.
#ifndef PB_BENIGN_Q
#define PB_BENIGN_Q 1
#endif
#ifndef PB_UNMERGE
#define PB_UNMERGE 0      /* 1: one reciprocal per denominator instead of merged products (experiment; slower) */
#endif
#if PB_BENIGN_Q
__device__ __forceinline__ double rcp_b(double x) { return rcp_q(x); }
__device__ __forceinline__ double sqrt_b_raw(double x) { return sqrt_q_raw(x); }
#else
__device__ __forceinline__ double rcp_b(double x) { return rcp(x); }
__device__ __forceinline__ double sqrt_b_raw(double x) { return sqrt_raw(x); }
#endif
__device__ __forceinline__ double sqrt_b_pos(double x) {
    const double g = sqrt_b_raw(x);
    return is_pos(x) ? g : 0.0;
}
#if PB_QUADRATIC
__device__ __forceinline__ double rcp_k(double x) { return rcp_q(x); }
__device__ __forceinline__ double sqrt_k_raw(double x) { return sqrt_q_raw(x); }
#else
__device__ __forceinline__ double rcp_k(double x) { return rcp(x); }
__device__ __forceinline__ double sqrt_k_raw(double x) { return sqrt_raw(x); }
#endif
__device__ __forceinline__ double sqrt_k_pos(double x) {
    const double g = sqrt_k_raw(x);
    return is_pos(x) ? g : 0.0;
}

// exp(x) for x <= 0 (clamped at about -700; positive x is outside the contract): n = rint(2048 x log2 e);
// exp = 2^(n>>11) * T[n&2047] * P3(r), |r| <= ln2/4096 (the cubic truncates at r^4/24 < 4e-17).  7 FP64 ops + one table load.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ exp2t) {
    // clamp at about -700: for x <= 0 the high word grows (as an unsigned integer) with |x|, so ONE integer min does it
    x = __hiloint2double((int)min((unsigned)__double2hiint(x), 0xC085E000u), __double2loint(x));
    const double t = fma(x, c_m[M_L2EN], PB_MAGIC);
    const int n = __double2loint(t);
    const double nd = t - PB_MAGIC;
    // ONE fused step: nd * (-ln2/2048) + x is rounded once, so the only reduction error is that of the constant,
    // |x| 2^-53 relative to exp(x) -- at most 4e-17 of absolute error after the multiplication by exp(x) <= 1
    const double r = fma(nd, c_m[M_NLN2S], x);
    double p = fma(r, c_m[M_E3], 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    const double v = exp2t[n & ((1 << EXP2_BITS) - 1)] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> EXP2_BITS) << 20), __double2loint(v));
}

// exp(-m lai) without ever forming the product (round 2n): the sample record carries c = -lai 2048 log2(e), so that
//   t = fma(m, c, 1.5 2^52) rounds x 2048 log2(e) to the integer n,  r = fma(m, c, -n) is the remainder in units of 1/2048 octave
// (one rounding, |r| <= 1/2), and exp = 2^(n>>11) T[n&2047] P3(r) with the cubic's coefficients scaled by powers of ln2/2048.
// 6 FP64 instructions + one table load (exp_neg of the product: 8).  The only new error is the rounding of c (2^-53 relative,
// i.e. |x| 1.1e-16 relative in the result -- what the product m lai had as well).  Underflow guard: ONE integer max on n
// (exponent floor 2^-1010); the prologue keeps |c| below 2^30 (lai < 3.6e5).
__device__ __forceinline__ double exp_neg_scaled(double m, double c, const double* __restrict__ exp2t) {
    const double t = fma(m, c, PB_MAGIC);
    const int n = max(__double2loint(t), -(1010 << EXP2_BITS));
    const double nd = t - PB_MAGIC;
    const double r = fma(m, c, -nd);
    double p = fma(r, c_m[M_S3], c_m[M_S2]);
    p = fma(r, p, c_m[M_S1]);
    p = fma(r, p, 1.0);
    const double v = exp2t[n & ((1 << EXP2_BITS) - 1)] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> EXP2_BITS) << 20), __double2loint(v));
}
#define PB_EXPS_SCALE 2954.6394437405970201   /* 2048 log2(e): what the double sample record multiplies -lai with (PB_EXPS) */

// log2(b) for normal b > 0: b = 2^E m, m in [1,2); i = top LOG_BITS = 9 mantissa bits; r = m/c_i - 1, |r| <= h = 2^-10;
// log2 b = E + log2 c_i + log2(1+r) with the cubic  (r - (1/2 + h^2/4) r^2 + r^3/3 + h^4/32)/ln2  -- the Taylor quartic with r^4
// economised to h^2 r^2 - h^4/8 (Chebyshev), its constant folded into the table's log2 c_i: absolute error <= h^4/(32 ln2) =
// 4.1e-14 (round 1/2k: 256 entries and the degree-4 series, 8e-15; b^(2(N-1)) - 1 amplifies it by <= 100).
// 6 FP64 ops + one 16-byte table load.
__device__ __forceinline__ double log2_pos(double b, const double* __restrict__ logt) {
    const int hi = __double2hiint(b);
    const int i = (hi >> (20 - LOG_BITS)) & ((1 << LOG_BITS) - 1);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(b));
    const double2 tc = *reinterpret_cast<const double2*>(logt + 2 * i);
    const double r = fma(m, tc.x, -1.0);
    // E as a double without I2F: 2^52 + 2^31 + (E biased by 2^31), minus the constant.
    const int E = (hi >> 20) - 1023;
    const double Ed = __hiloint2double(0x43300000, E ^ 0x80000000) - c_m[M_EBIAS];
    double q = fma(r, c_m[M_L3], c_m[M_L2E]);
    q = fma(r, q, c_m[M_L1]);
    return fma(r, q, Ed + tc.y);
}

// 2^y (clamped at about 500): n = rint(2048 y); r = y - n/2048 (exact); 2^y = 2^(n>>11) T[n&2047] P3(r ln2).
// 7 FP64 ops + one table load.
__device__ __forceinline__ double exp2_any(double y, const double* __restrict__ exp2t) {
    // clamp at about 500: one signed integer min on the high word (negative y has a negative high word and passes)
    y = __hiloint2double(min(__double2hiint(y), 0x407F4000), __double2loint(y));
    const double t = fma(y, (double)(1 << EXP2_BITS), PB_MAGIC);
    const int n = __double2loint(t);
    const double nd = t - PB_MAGIC;
    const double r = fma(nd, -1.0 / (1 << EXP2_BITS), y);
    double p = fma(r, c_m[M_P3], c_m[M_P2]);
    p = fma(r, p, c_m[M_P1]);
    p = fma(r, p, 1.0);
    const double v = exp2t[n & ((1 << EXP2_BITS) - 1)] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> EXP2_BITS) << 20), __double2loint(v));
}

// b^e = 2^(e log2 b) without forming the exponent (round 2n): with the rounding magic in units of 2^-11 (1.5 2^41; its low word
// is zero, so it is an immediate operand), t = fma(L, e, magic) rounds e L to a multiple of 1/2048, r = fma(L, e, -(t - magic))
// is the exact remainder rounded once, and the rest is exp2_any with a quadratic.  12 FP64 instructions (pow_pos: 15).  The clamp of the exponent
// at about 500 becomes an integer min of L against 500/|e| (high word, from the sample record: lmax_hi).
__device__ __forceinline__ double pow_pos_scaled(double b, double e, int lmax_hi, const MathTables& mt) {
    double L = log2_pos(b, mt.logt);
    L = __hiloint2double(min(__double2hiint(L), lmax_hi), __double2loint(L));
    const double MG = PB_MAGIC / (double)(1 << EXP2_BITS);
    const double t = fma(L, e, MG);
    const int n = __double2loint(t);
    const double nd = t - MG;                      // n / 2048
    const double r = fma(L, e, -nd);
    // economised quadratic of 2^r (r^3 ln2^3/6 ~ (g^2/8) r ln2 on |r ln2| <= g = ln2/4096): 2e-13 relative, one FMA less than the cubic
    double p = fma(r, c_m[M_P2], c_m[M_P1Q]);
    p = fma(r, p, 1.0);
    const double v = mt.exp2t[n & ((1 << EXP2_BITS) - 1)] * p;
    return __hiloint2double(__double2hiint(v) + ((n >> EXP2_BITS) << 20), __double2loint(v));
}

// type-generic spellings used by the kernel template (the float twins are in math32.cuh)
__device__ __forceinline__ double pow_pos(double b, double e, const MathTables& mt) { return exp2_any(log2_pos(b, mt.logt) * e, mt.exp2t); }
__device__ __forceinline__ double exp_neg(double x, const MathTables& mt) { return exp_neg(x, mt.exp2t); }

// ---- tau(k) = (1-k) e^-k + k^2 E1(k)   (reference models.py:953-957; tau = 1 for k <= 0) ----
//
// Fast path: branch-free table evaluation for k in [2^TAU_EMIN, 2^(TAU_EMAX+1)) = [6.1e-5, 32).  The interval index and the
// re-scaled mantissa come from integer operations on the bits of k; the FP64 pipe sees 1 subtraction (local variable) +
// TAU_DEG Horner steps (the upper ones on the FP32 pipe, PB_TAU_F32TAIL).  Arguments outside the range are clamped to a valid row
// (the value is then meaningless) and flagged in `rare`; the caller repairs flagged lanes with tau_rare() after
// both evaluations of a thread have been issued.
#if PB_TAU_F32TAIL
// Interval scheme (tools/gen_math_tables.py): 2^TAU_SB equal sub-intervals of every binade 2^TAU_EMIN .. 2^(TAU_EMAX+1).  The row
// index is the top 11 + TAU_SB bits of k minus a constant (one shift, one add, one clamp), and the local variable is
// u = mm - O with mm = the mantissa of k re-scaled to [2^(TAU_SB+1), 2^(TAU_SB+2)) by ONE logic instruction on the high word
// and O the odd integer stored in the row: a DADD with two register operands (round 1: FMA with a power-of-two scale built
// from the exponent, three registers, ~20 integer instructions for a two-region index).
__device__ __forceinline__ double tau_fast(double k, const TauRow* __restrict__ tab, bool& rare) {
    const int hi = __double2hiint(k);
    const int raw = (hi >> (20 - TAU_SB)) - ((1023 + TAU_EMIN) << TAU_SB);     // negative for k < 2^TAU_EMIN, k <= 0 (sign bit) and NaN payloads
    const int idx = min(max(raw, 0), TAU_NROWS - 1);
    rare = idx != raw;
    const double mm = __hiloint2double((hi & 0x000FFFFF) | ((1023 + TAU_SB + 1) << 20), __double2loint(k));
    const uint4* row = reinterpret_cast<const uint4*>(tab + idx);
    const uint4 q0 = row[0], q2 = row[2], q1 = row[1];
    const double u = mm - __hiloint2double((int)q0.y, (int)q0.x);
    const float uf = (float)u;                                     // F2F.F32.F64 (SFU pipe)
    float t = fmaf(uf, __uint_as_float(q2.w), __uint_as_float(q2.z));          // f5 + u f6
    t = fmaf(uf, t, __uint_as_float(q2.y));                        // f4
    t = fmaf(uf, t, __uint_as_float(q2.x));                        // f3
    double acc = fma(u, (double)t, __hiloint2double((int)q1.w, (int)q1.z));    // c2 + u T3   (F2F.F64.F32)
    acc = fma(u, acc, __hiloint2double((int)q1.y, (int)q1.x));    // c1
    return fma(u, acc, __hiloint2double((int)q0.w, (int)q0.z));   // c0
}
#else
__device__ __forceinline__ double tau_fast(double k, const double* __restrict__ tab, bool& rare) {
    const int hi = __double2hiint(k);
    const int raw = (hi >> (20 - TAU_SB)) - ((1023 + TAU_EMIN) << TAU_SB);
    const int idx = min(max(raw, 0), TAU_NROWS - 1);
    rare = idx != raw;
    const double mm = __hiloint2double((hi & 0x000FFFFF) | ((1023 + TAU_SB + 1) << 20), __double2loint(k));
    const double2* row = reinterpret_cast<const double2*>(tab + idx * TAU_ROW);
    const double2 c01 = row[0];                                    // O, c0
    const double u = mm - c01.x;
    static_assert(TAU_DEG % 2 == 0, "row layout assumes an even degree (TAU_ROW even)");
#ifdef PB_TAU_EVENODD
    // even / odd split: two Horner chains in u^2 (depth 7 instead of 10, one extra multiply)
    const double u2 = u * u;
    double2 c = row[TAU_ROW / 2 - 1];                              // c[D-1], c[D]
    double po = c.x, pe = c.y;
#pragma unroll
    for (int q = TAU_ROW / 2 - 2; q >= 1; --q) {
        c = row[q];                                                // c[2q-1], c[2q]
        po = fma(u2, po, c.x);
        pe = fma(u2, pe, c.y);
    }
    pe = fma(u2, pe, c01.y);
    return fma(u, po, pe);
#else
    double2 c = row[TAU_ROW / 2 - 1];                              // c[D-1], c[D]
    double acc = fma(u, c.y, c.x);
#pragma unroll
    for (int q = TAU_ROW / 2 - 2; q >= 1; --q) {
        c = row[q];
        acc = fma(u, acc, c.y);
        acc = fma(u, acc, c.x);
    }
    return fma(u, acc, c01.y);
#endif
}

#endif

// Rare path (never taken for realistic leaves, k in [6e-4, 16]): k <= 0, k < 2^-14, k >= 32.  Plain IEEE arithmetic.
__device__ __noinline__ double tau_rare(double k) {
    if (!(k > 0.0)) return 1.0;                                    // k <= 0 (or NaN): the reference leaves tau = 1
    if (k < 1.0) {
        // E1(k) = -gamma - ln k + k - k^2/4 + k^3/18 - ...   (k < 6.1e-5: the first omitted term is below 1e-27)
        const double e1 = -0.57721566490153286061 - log(k) + k * (1.0 - k * (0.25 - k * (1.0 / 18.0)));
        return (1.0 - k) * exp(-k) + k * k * e1;
    }
    // k >= 32: E1(k) = e^-k / (k + 1 - 1/(k + 3 - 4/(k + 5 - 9/(k + 7 - ...))))   (12 levels: < 1e-16 relative)
    double d = k + 25.0;
    for (int i = 12; i >= 1; --i) d = k + (2 * i - 1) - (double)(i * i) / d;
    return exp(-k) * ((1.0 - k) + k * k / d);
}

// Table-free tau(k) for any k (plain IEEE arithmetic + libm; absolute error < 2e-15): E1 by its power series below 1 and by
// the continued fraction above, evaluated bottom-up with a fixed depth (stable: no error accumulation over the levels).
// Used where no shared-memory table is at hand and speed does not matter (the one-plate attribute planes of PROSPECT).
__device__ __noinline__ double tau_full(double k) {
    if (!(k > 0.0)) return 1.0;
    double e1;
    if (k < 1.0) {                       // E1(k) = -gamma - ln k + sum_{n>=1} (-1)^(n+1) k^n / (n n!)
        double term = k, sum = k;
        for (int n = 2; n <= 24; ++n) {
            term *= -k / n;
            sum += term / n;
        }
        e1 = -0.57721566490153286061 - log(k) + sum;
    } else {                             // E1(k) = e^-k / (k + 1 - 1/(k + 3 - 4/(k + 5 - ...))), depth ~ 90/k + 20 levels
        const int levels = (int)(90.0 / k) + 20;
        double d = k + (2.0 * levels + 1.0);
        for (int i = levels; i >= 1; --i) d = k + (2.0 * i - 1.0) - (double)i * (double)i / d;
        e1 = exp(-k) / d;
    }
    return (1.0 - k) * exp(-k) + k * k * e1;
}

// full-range tau (test kernel, and the reference point for the fast path)
#if PB_TAU_F32TAIL
typedef TauRow TauTab;
#else
typedef double TauTab;
#endif
__device__ __forceinline__ double tau_plate(double k, const TauTab* __restrict__ tab) {
    bool rare;
    const double t = tau_fast(k, tab, rare);
    return rare ? tau_rare(k) : t;
}

// warp-level sum of a double (5 butterfly steps).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace pb
