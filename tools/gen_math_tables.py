#!/usr/bin/env python
"""Generate pyrism_b200/csrc/math_tables.inc: the fp64 lookup tables used by the sm_100a kernels.

1. TAU table -- piecewise polynomials for the plate transmissivity of PROSPECT,
       tau(k) = (1-k) exp(-k) + k^2 E1(k)  ( = 2 E3(k) ),        reference models.py:953-957,
   on intervals that are addressed with integer operations on the exponent/mantissa bits of k
   (so that the FP64 pipe only sees one FMA for the local variable plus the Horner chain); with SB = TAU_SB:
     region A, k in [2^-14, 2):  2^SB equal sub-intervals per binade                       (15 * 2^SB intervals)
     region B, k in [2, 32):     sub-intervals of width 2^(1-SB) (2^(E+SB-1) per binade E=1..4)  (15 * 2^SB intervals)
   SB = 3 and degree 8 (240 rows of 10 doubles) replaced SB = 2 and degree 10: two three-register DFMAs less per evaluation.
   Round 2n: SB = 4 and degree 6 (304 rows): |c3| <= 5.1e-6 over the whole table, so the terms of degree >= 3 can run in float
   (6e-8 relative: 3e-13 absolute) -- 4 FP64 instructions per evaluation instead of 5, 3 FFMA instead of 4, and the row
   (O, c0, c1, c2 in double + c3..c6 in float) is 48 bytes: three 128-bit loads instead of four.  The degree-6 fit is good to
   4e-14 absolute (bar of the table 1e-13, budget of an output 1e-11).
   Row layout: [O, c0, c1, ..., cD] with u = k*S - O in [-1, 1], S a power of two, tau ~= sum c_i u^i.
   Outside (rare path of the kernel): k < 2^-14 -> series with log(k);  k >= 32 -> continued fraction of E1.
2. EXP2 table -- 2^(j/2^EXP2_BITS), j = 0..2^EXP2_BITS - 1 (2048 entries: a cubic in the reduced argument is then exact to 4e-17).
3. LOG table  -- for c_i = 1 + (i + 0.5)/2^LOG_BITS, i = 0..2^LOG_BITS - 1:  1/c_i (rounded) and -log2(1/c_i) paired
   (round 2n: 512 entries, |r| <= 2^-10, so that an economised cubic of log2(1 + r) is good to 4e-14: one FMA less).

Coefficients are fitted with mpmath (50 digits) at Chebyshev nodes, converted to the monomial basis in u
and rounded to double.  The script then checks the double-precision Horner evaluation against mpmath on
random points and prints the worst absolute / relative error; it refuses to write the file if the
error target is missed.
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 60
DEG = int(os.environ.get("TAU_DEG", "6"))
SB = int(os.environ.get("TAU_SB", "4"))
LOG_BITS = int(os.environ.get("LOG_BITS", "9"))
EXP2_BITS = int(os.environ.get("EXP2_BITS", "11"))
EMIN = -14
EMAX = 4
NROWS = (EMAX - EMIN + 1) * 2 ** SB
OUT = os.environ.get("MATH_TABLES_OUT") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pyrism_b200", "csrc", "math_tables.inc")


def tau(k):
    k = mp.mpf(k)
    return (1 - k) * mp.exp(-k) + k * k * mp.e1(k)


def intervals():
    """(lo, hi, S, O) per interval in table order: 2^SB equal sub-intervals of EVERY binade 2^EMIN .. 2^(EMAX+1), so that the
    row index is just the top 11 + SB bits of k minus a constant and the local variable is the mantissa (scaled to
    [2^(SB+1), 2^(SB+2))) minus the odd integer O = 2^(SB+1) + 2 j + 1."""
    rows = []
    for E in range(EMIN, EMAX + 1):
        for j in range(2 ** SB):
            lo = mp.mpf(2) ** E * (1 + mp.mpf(j) / 2 ** SB)
            w = mp.mpf(2) ** (E - SB)
            rows.append((lo, lo + w, mp.mpf(2) ** (SB + 1 - E), 2 ** (SB + 1) + 2 * j + 1))
    return rows


def cheb_fit_monomial(f, lo, hi, deg):
    """Monomial coefficients (in u in [-1,1]) of the degree-`deg` Chebyshev interpolant of f on [lo,hi]."""
    n = deg + 1
    c, h = (lo + hi) / 2, (hi - lo) / 2
    nodes = [mp.cos(mp.pi * (2 * i + 1) / (2 * n)) for i in range(n)]
    fv = [f(c + h * x) for x in nodes]
    # Chebyshev coefficients
    a = []
    for j in range(n):
        s = mp.fsum(fv[i] * mp.cos(mp.pi * j * (2 * i + 1) / (2 * n)) for i in range(n))
        a.append(s * (1 if j == 0 else 2) / n)
    # T_j -> monomials
    T = [[mp.mpf(1)], [mp.mpf(0), mp.mpf(1)]]
    for j in range(2, n):
        prev, prev2 = T[j - 1], T[j - 2]
        cur = [mp.mpf(0)] + [2 * v for v in prev]
        for i, v in enumerate(prev2):
            cur[i] -= v
        T.append(cur)
    mono = [mp.mpf(0)] * n
    for j in range(n):
        for i, v in enumerate(T[j]):
            mono[i] += a[j] * v
    return mono


def main():
    rows = intervals()
    assert len(rows) == NROWS
    table = np.zeros((len(rows), DEG + 2))
    worst_abs = worst_rel = 0.0
    rng = np.random.default_rng(1)
    for r, (lo, hi, S, O) in enumerate(rows):
        mono = cheb_fit_monomial(tau, lo, hi, DEG)
        table[r, 0] = float(O)
        table[r, 1:] = [float(v) for v in mono]
        # validation: double Horner vs mpmath
        ks = np.concatenate([np.float64(lo) + (np.float64(hi) - np.float64(lo)) * rng.random(24),
                             [np.float64(lo), np.nextafter(np.float64(hi), 0)]])
        u = ks * float(S) - float(O)
        acc = np.full_like(u, table[r, -1])
        for i in range(DEG, 0, -1):
            acc = acc * u + table[r, i]
        for kk, got in zip(ks, acc):
            ref = tau(mp.mpf(float(kk)))
            ea = abs(mp.mpf(float(got)) - ref)
            worst_abs = max(worst_abs, float(ea))
            worst_rel = max(worst_rel, float(ea / ref))
    print("tau table: degree %d, %d intervals, worst abs err %.3e, worst rel err %.3e" % (DEG, len(rows), worst_abs, worst_rel))
    # what the plate formulas see is the ABSOLUTE error of tau (tau <= 1 enters r, t, Ra, Ta linearly; a tiny tau only yields a
    # tiny transmittance): the bar is 1e-9 on reflectance / transmittance, the budget of an output 1e-11, the
    # budget of this table 1e-13 absolute
    if worst_abs > 1e-13 or worst_rel > 1e-6:
        sys.exit("tau table misses the 1e-13 absolute / 1e-6 relative target")

    # float32 table for the fp32 mode: same intervals, degree DEG_F, rows [O, c0..c5, pad], of
    #     g(k) = (1 - tau(k))/k ,   so that   1 - tau = k g(k)   carries a RELATIVE error of ~1e-7.
    # The float kernels build 1 - r - t = t12 (1 - tau)/(1 - r21 tau) from it (no cancellation near zero absorption).
    def gfun(k):
        k = mp.mpf(k)
        return (1 - tau(k)) / k
    DEG_F = 5
    table_f = np.zeros((len(rows), 8), dtype=np.float32)
    worst_f = 0.0
    for r, (lo, hi, S, O) in enumerate(rows):
        mono = cheb_fit_monomial(gfun, lo, hi, DEG_F)
        table_f[r, 0] = float(O)
        table_f[r, 1:DEG_F + 2] = [float(v) for v in mono]
        ks = (np.float64(lo) + (np.float64(hi) - np.float64(lo)) * rng.random(16)).astype(np.float32)
        ks = ks[(ks >= np.float32(float(lo))) & (ks < np.float32(float(hi)))]
        u = ks * np.float32(float(S)) - table_f[r, 0]
        acc = np.full_like(u, table_f[r, DEG_F + 1])
        for i in range(DEG_F, 0, -1):
            acc = acc * u + table_f[r, i]
        for kk, got in zip(ks, acc):
            ref = gfun(mp.mpf(float(kk)))
            worst_f = max(worst_f, float(abs(mp.mpf(float(got)) - ref) / ref))
    print("(1 - tau)/k float table: degree %d, worst rel err %.3e" % (DEG_F, worst_f))
    if worst_f > 4e-7:
        sys.exit("float table misses the 4e-7 relative target")

    exp2t = [float(mp.mpf(2) ** (mp.mpf(j) / 2 ** EXP2_BITS)) for j in range(2 ** EXP2_BITS)]
    logt = []
    for i in range(2 ** LOG_BITS):
        c = 1 + (mp.mpf(i) + mp.mpf(1) / 2) / 2 ** LOG_BITS
        inv = float(1 / c)                                  # rounded reciprocal, used as-is in the kernel
        # log2(c') for the c' = 1/inv actually used, plus the constant h^4/(32 ln 2) of the economised cubic (math64.cuh: log2_pos)
        logt.append((inv, float(-mp.log(mp.mpf(inv), 2) + mp.mpf(2) ** (-4 * (LOG_BITS + 1)) / (32 * mp.log(2)))))

    with open(OUT, "w") as fh:
        fh.write("// GENERATED by tools/gen_math_tables.py -- do not edit.\n")
        fh.write("// tau(k) = (1-k)exp(-k) + k^2 E1(k): degree-%d piecewise polynomials; worst abs %.2e, worst rel %.2e (vs mpmath).\n"
                 % (DEG, worst_abs, worst_rel))
        fh.write("#define TAU_DEG %d\n#define TAU_ROW %d\n#define TAU_NROWS %d\n#define TAU_EMIN (%d)\n#define TAU_EMAX (%d)\n"
                 "#define TAU_SB %d\n#define EXP2_BITS %d\n#define LOG_BITS %d\n" % (DEG, DEG + 2, len(rows), EMIN, EMAX, SB, EXP2_BITS, LOG_BITS))
        fh.write("static const double h_tau_table[TAU_NROWS * TAU_ROW] = {\n")
        for r in range(len(rows)):
            fh.write("  " + ", ".join(float(v).hex() for v in table[r]) + ",\n")
        fh.write("};\n")
        fh.write("#define TAU_ROW_F 8\n// float32 table of g(k) = (1 - tau(k))/k (fp32 mode): degree 5, worst rel %.2e\n" % worst_f)
        fh.write("static const float h_tau_table_f[TAU_NROWS * TAU_ROW_F] = {\n")
        for r in range(len(rows)):
            fh.write("  " + ", ".join("%sf" % float(v).hex() for v in table_f[r]) + ",\n")
        fh.write("};\n")
        fh.write("static const double h_exp2_table[1 << EXP2_BITS] = {\n  " + ",\n  ".join(v.hex() for v in exp2t) + "\n};\n")
        fh.write("static const double h_log_table[2 << LOG_BITS] = {\n  " + ",\n  ".join("%s, %s" % (a.hex(), b.hex()) for a, b in logt) + "\n};\n")
    print("wrote", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
