"""fp64 primitives of the kernels, evaluated on the GPU through prosail_test_math(), against numpy / mpmath."""
import mpmath as mp
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
rng = np.random.default_rng(2026)


def rel(a, b):
    return float(np.max(np.abs(a - b) / np.abs(b)))


def test_rcp_and_sqrt(engine):
    x = np.concatenate([rng.uniform(1e-3, 1e3, 50000), 10.0 ** rng.uniform(-200, 200, 50000)])
    assert rel(engine.test_math(0, x), 1.0 / x) < 4e-16
    assert rel(engine.test_math(0, -x), -1.0 / x) < 4e-16
    assert rel(engine.test_math(1, x), np.sqrt(x)) < 4e-16
    assert np.array_equal(engine.test_math(1, np.array([0.0, -1.0])), [0.0, 0.0])


def test_quadratic_newton_steps(engine):
    """rcp_q / sqrt_q_raw (the PB_QUADRATIC=1 build variant of the band kernel): ONE second-order step on the ~2^-20 MUFU
    seeds -> about 2^-40 relative (measured 9.2e-13)."""
    x = np.concatenate([rng.uniform(1e-3, 1e3, 50000), 10.0 ** rng.uniform(-200, 200, 50000), 1.0 + rng.uniform(0, 1, 50000)])
    assert rel(engine.test_math(7, x), 1.0 / x) < 1.0e-12
    assert rel(engine.test_math(7, -x), -1.0 / x) < 1.0e-12
    assert rel(engine.test_math(8, x), np.sqrt(x)) < 3.0e-12      # e = 1 - x y^2 ~ 2 delta: 3/8 e^2 = 1.5 delta^2


def test_exp_of_nonpositive(engine):
    x = -np.concatenate([rng.uniform(0, 60, 50000), 10.0 ** rng.uniform(-14, 2.8, 50000), [0.0, 700.0, 5000.0]])
    got = engine.test_math(2, x)
    xc = np.maximum(x, -700.0)
    want = np.exp(xc)
    # one-step argument reduction: the rounding of the constant ln2/2048 adds |x| 2^-53 relative, i.e. < 4.1e-17 ABSOLUTE
    assert np.max(np.abs(got - want) / want / (6e-16 + 1.2e-16 * np.abs(xc))) < 1.0
    assert np.max(np.abs(got - want)) < 2.5e-16


def test_log2_exp2(engine):
    x = np.concatenate([rng.uniform(1.0, 4.0, 50000), 10.0 ** rng.uniform(0, 300, 20000), 1.0 + 10.0 ** rng.uniform(-12, 0, 20000)])
    got = engine.test_math(3, x)
    want = np.array([float(mp.log(mp.mpf(float(v)), 2)) for v in x[::37]])
    # design accuracy (round 2n): 512-entry table, economised cubic on |r| <= 2^-10 -> absolute error <= 2^-40 / (32 ln 2) = 4.1e-14
    # (plus rounding); round 1: 256 entries and the degree-4 series, 8.2e-15
    assert np.max(np.abs(got[::37] - want) / np.maximum(np.abs(want), 1.0)) < 4.6e-14
    assert np.max(np.abs(got - np.log2(x)) / np.maximum(np.log2(x), 1.0)) < 4.6e-14
    y = np.concatenate([rng.uniform(-300, 500, 50000), rng.uniform(-1, 1, 20000)])
    assert rel(engine.test_math(4, y), np.exp2(y)) < 6e-16


def test_scaled_exp_and_pow(engine):
    """Round 2n: exp(-m lai) from the pre-scaled record value (exp_neg_scaled) and b^(N-1) without forming the exponent
    (pow_pos_scaled: economised quadratic of 2^r, 2e-13 relative, on top of the 4.1e-14 absolute of log2)."""
    x = -np.concatenate([rng.uniform(0, 60, 50000), 10.0 ** rng.uniform(-14, 2.8, 50000), [0.0, 700.0]])
    got = engine.test_math(9, x)
    want = np.exp(x)
    # the record value carries one rounding (2^-53 relative of |x|), like the product m lai did
    assert np.max(np.abs(got - want) / want / (6e-16 + 2.3e-16 * np.abs(x))) < 1.0
    assert np.max(np.abs(got - want)) < 2.5e-16
    assert np.all(engine.test_math(9, np.array([-5000.0, -1.0e5])) < 1e-300)          # exponent floor, no wrap-around
    b = np.concatenate([1.0 + 10.0 ** rng.uniform(-8, 0, 30000), 10.0 ** rng.uniform(0, 12, 30000)])
    got = engine.test_math(10, b)
    want = np.array([float(mp.mpf(float(v)) ** mp.mpf(0.75)) for v in b[::29]])
    # relative error <= 2.1e-13 (quadratic) + 0.75 ln2 * 4.1e-14 (log2) + rounding
    assert np.max(np.abs(got[::29] - want) / want) < 2.6e-13
    assert np.isfinite(engine.test_math(10, np.array([1e300]))).all()                  # clamped at about 2^500


def test_tau_against_mpmath(engine):
    """tau(k) = (1-k)e^-k + k^2 E1(k) (models.py:953-957) over the whole argument range, incl. the interval edges."""
    edges = np.concatenate([2.0 ** np.arange(-31, 7), np.arange(2, 65, 0.5), [np.nextafter(2.0, 0), np.nextafter(64.0, 0), np.nextafter(1.0, 0)]])
    k = np.concatenate([10.0 ** rng.uniform(-12, 1.8, 3000), rng.uniform(0, 64, 3000), edges, edges * (1 + 1e-15)])
    k = k[k > 0]
    mp.mp.dps = 40
    want = np.array([float((1 - mp.mpf(float(v))) * mp.exp(-mp.mpf(float(v))) + mp.mpf(float(v)) ** 2 * mp.e1(mp.mpf(float(v)))) for v in k])
    got = engine.test_math(5, k)
    # round 2 error budget (csrc/math64.cuh, TauRow): the terms of degree >= 3 of the piecewise polynomial (degree 6, 16 sub-intervals
    # per binade) are evaluated in float -- |c3| <= 5.1e-6 times a few float roundings (6e-8 each) = a few 1e-13 absolute, and the same RELATIVE to the row's
    # own magnitude (the coefficients of a row scale with tau), against a parity bar of 1e-9 on reflectance / transmittance
    assert np.max(np.abs(got - want)) < 5e-13
    # relative accuracy where tau is O(1) (k <= 2); beyond, the table (16 sub-intervals per binade up to k = 32) is built to an
    # ABSOLUTE target -- tau(k > 2) < 0.06 only yields a correspondingly small transmittance
    inside = k <= 2
    assert rel(got[inside], want[inside]) < 2e-11
    mid = (k > 2) & (k < 64)
    assert rel(got[mid], want[mid]) < 1e-6
    # outside the tabulated range
    assert np.array_equal(engine.test_math(5, np.array([0.0, -3.0])), [1.0, 1.0])          # k <= 0 -> tau = 1
    big = np.array([64.0, 80.0, 200.0, 700.0])
    wb = np.array([float((1 - mp.mpf(v)) * mp.exp(-mp.mpf(v)) + mp.mpf(v) ** 2 * mp.e1(mp.mpf(v))) for v in big])
    gb = engine.test_math(5, big)
    assert np.all(np.isfinite(gb)) and np.max(np.abs(gb - wb)) < 1e-30 and rel(gb[:3], wb[:3]) < 1e-6


def test_tau_against_scipy_expi(engine):
    """...and against the reference's own formulation with scipy.special.expi."""
    from scipy.special import expi
    k = 10.0 ** rng.uniform(-6, 1.7, 20000)
    want = (1 - k) * np.exp(-k) + k ** 2 * (-expi(-k))
    assert np.max(np.abs(engine.test_math(5, k) - want)) < 5e-13      # fp32 tail of the polynomial, see test_tau_against_mpmath


def test_table_free_tau_against_mpmath(engine):
    """tau_full (op 6): the table-free evaluation behind the one-plate attribute planes."""
    import mpmath as mp
    rng = np.random.default_rng(5)
    k = np.concatenate([10.0 ** rng.uniform(-10, 2.3, 1500), rng.uniform(1.5, 2.5, 200), [1e-300, 2.0, 1.9999999, 50.0, 200.0]])
    ref = np.array([float((1 - mp.mpf(v)) * mp.exp(-mp.mpf(v)) + mp.mpf(v) ** 2 * mp.e1(mp.mpf(v))) for v in k])
    got = engine.test_math(6, k)
    assert np.max(np.abs(got - ref)) < 2e-15                       # absolute: what r, t, Ra, Ta see
    m = k <= 20.0                                                   # (1-k)e^-k + k^2 E1(k) cancels ~k^2/2 digits (in the reference too)
    assert np.max(np.abs(got[m] - ref[m]) / ref[m]) < 3e-13
    assert np.array_equal(engine.test_math(6, np.array([0.0, -2.0])), [1.0, 1.0])
