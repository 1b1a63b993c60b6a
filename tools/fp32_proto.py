#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""Numerical prototype of the fp32 mode (CPU, numpy float32): which formulation of the PROSPECT -> 4SAIL chain
stays within 1e-5 absolute of the float64 oracle?  Development tool -- not part of the product path.

    python tools/fp32_proto.py [n_samples]

Per-sample scalars (volume scattering, hotspot) are taken from the float64 oracle, as the fp32 kernels take them
from the float64 prologue.  numpy has no fma, so every a*b+c here rounds twice: a pessimistic model of the kernel.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import prosail_oracle as O   # noqa: E402
from pyrism_b200 import constants as K   # noqa: E402
from bench import lhs                    # noqa: E402

f32 = np.float32
TAU_ERR = float(os.environ.get("TAU_ERR", "4e-7"))


def tau_f(k):
    from scipy.special import expi
    k64 = k.astype(np.float64)
    t = (1 - k64) * np.exp(-k64) + k64 * k64 * (-expi(-k64))
    return (t + TAU_ERR * (np.random.rand(*k.shape) - 0.5)).astype(k.dtype)     # table error model


def leaf(T, p, talf, t12, t21, variant):
    kc = K.lib.p5
    c = [T(p[n] / p['N']) for n in ('Cab', 'Cxc', 'Cbr', 'Cw', 'Cm')]
    tabs = [kc.Kab.astype(T), kc.Kxc.astype(T), kc.Kbr.astype(T), kc.Kw.astype(T), kc.Km.astype(T)]
    kall = c[0] * tabs[0]
    for ci, ti in zip(c[1:], tabs[1:]):
        kall = kall + ci * ti
    tau = tau_f(kall)
    if variant >= 3:
        # table of g(k) = (1 - tau(k))/k with ~1e-7 relative error: omt = k g(k), tau = 1 - omt
        from scipy.special import expi
        k64 = np.maximum(kall.astype(np.float64), 1e-300)
        omt64 = -np.expm1(-k64) + k64 * np.exp(-k64) - k64 * k64 * (-expi(-k64))
        omt = (omt64 * (1 + 1e-7 * (np.random.rand(*kall.shape) - 0.5))).astype(T)
        omt = np.where(kall > 0, omt, T(0))
        tau = (T(1) - omt).astype(T)
    talf, t12, t21 = talf.astype(T), t12.astype(T), t21.astype(T)
    ralf, r12, r21 = T(1) - talf, T(1) - t12, T(1) - t21
    one = T(1)
    x = r21 * tau
    inv = one / (one - x * x)
    tt = tau * t21 * inv
    Ta = talf * tt
    t = t12 * tt
    Ra = ralf + x * Ta
    r = r12 + x * t
    pq, q = one + r, one - r
    P12 = (pq + t) * (pq - t)
    nm1 = T(p['N'] - 1)
    if variant >= 3:
        # 1 - r - t = t12 (1 - tau)/(1 - r21 tau): exact, no cancellation
        qmt = t12 * omt * (one + x) * inv
        P34 = (q + t) * qmt
        D = np.sqrt(np.maximum(P12 * P34, 0))
        da = T(0.5) * (P34 + D)                   # a r - r
        al = r + da
        db = T(0.5) * (qmt * (pq - t) + D)        # b t - t
        b = one + db / t
        bNm1 = np.exp2(np.log2(b) * nm1).astype(T)
        bN2 = bNm1 * bNm1
        bm1 = bN2 - one
        r2 = r * r
        dal = da * (al + r)                       # al^2 - r^2
        X = dal + (al * bm1) * (al - r2)
        with np.errstate(all='ignore'):
            TaI = Ta / X
            tran = TaI * (bNm1 * dal)
            refl = Ra + TaI * bm1 * ((t * al) * r)
        zero = ~(qmt > 0)
        if zero.any():
            Tsub = t / (t + (one - t) * nm1)
            Rsub = one - Tsub
            den = one - Rsub * r
            tran = np.where(zero, Ta * Tsub / den, tran)
            refl = np.where(zero, Ra + Ta * Rsub * t / den, refl)
        return refl.astype(T), tran.astype(T)
    if variant >= 1:
        # 1 - r - t without the cancellation of r12 + x t:  1 - r12 - x t - t = t12 - t (1 + x)
        qmt = t12 - t * (one + x)
    else:
        qmt = q - t
    P34 = (q + t) * qmt
    D = np.sqrt(np.maximum(P12 * P34, 0))
    A = P12 - T(2) * r
    hD = T(0.5) * D
    al = T(0.5) * A + hD
    be = one + hD - T(0.5) * A
    b = be / t
    bNm1 = np.exp2(np.log2(b) * nm1).astype(T)
    bN2 = bNm1 * bNm1
    bm1 = bN2 - one
    al2, r2 = al * al, r * r
    X = al2 * bN2 - r2 - (al * r2) * bm1
    with np.errstate(all='ignore'):
        TaI = Ta / X
        tran = TaI * (bNm1 * (al2 - r2))
        refl = Ra + TaI * bm1 * ((t * al) * r)
    return refl.astype(T), tran.astype(T)


def canopy(T, rho, tau, rs, g, lai, variant):
    one = T(1)
    ddb, ddf = T(0.5 * (1 + g['bf'])), T(0.5 * (1 - g['bf']))
    k, Kk = T(g['vol_ks']), T(g['vol_ko'])
    sdb, sdf = T(0.5 * (g['vol_ks'] + g['bf'])), T(0.5 * (g['vol_ks'] - g['bf']))
    dob, dof = T(0.5 * (g['vol_ko'] + g['bf'])), T(0.5 * (g['vol_ko'] - g['bf']))
    Fs, Ft = T(g['Fs']), T(g['Ft'])
    tss, too, tsstoo = T(g['tss']), T(g['too']), T(g['tsstoo'])
    laisum = T(lai * g['sumint'])
    z = T((1 - g['tss'] * g['too']) / (g['vol_ks'] + g['vol_ko']))
    lai = T(lai)
    sigb = ddb * rho + ddf * tau
    sigf = ddf * rho + ddb * tau
    att = one - sigf
    if variant >= 1:
        m = np.sqrt((att - sigb) * (att + sigb))
        rinf = sigb / (att + m)
    else:
        m = np.sqrt(att * att - sigb * sigb)
        rinf = (att - m) / sigb
    e1 = np.exp(-m * lai).astype(T)
    e2 = e1 * e1
    rinf2 = rinf * rinf
    re = rinf * e1
    invden = one / (one - rinf2 * e2)
    omr2 = one - rinf2
    tdd = omr2 * e1 * invden
    rdd = rinf * (one - e2) * invden
    rsrdd = rs * rdd
    rs_invdn = rs / (one - rsrdd)
    sb = sdb * rho + sdf * tau
    sf = sdf * rho + sdb * tau
    vb = dob * rho + dof * tau
    vf = dof * rho + dob * tau
    w = Fs * rho + Ft * tau
    km, kp, Km, Kp = k - m, k + m, Kk - m, Kk + m
    if variant >= 2:
        # J1 = e^{-kt} (e^{(k-m) t} - 1)/(k - m) = tss * lai * expm1(d)/d, d = (k - m) lai
        def em1x(d):
            d64 = d.astype(np.float64)
            with np.errstate(all='ignore'):
                v = np.where(np.abs(d64) < 1e-12, 1.0, np.expm1(d64) / d64)
            return v.astype(T)
        J1ks = tss * lai * em1x(km * lai)
        J1ko = too * lai * em1x(Km * lai)
    else:
        J1ks = (e1 - tss) / km
        J1ko = (e1 - too) / Km
    inv_kp, inv_Kp = one / kp, one / Kp
    J2ks = (one - e1 * tss) * inv_kp
    J2ko = (one - e1 * too) * inv_Kp
    u1, u2 = sf + sb * rinf, sf * rinf + sb
    v1, v2 = vf + vb * rinf, vf * rinf + vb
    Pss, Qss, Pv, Qv = u1 * J1ks, u2 * J2ks, v1 * J1ko, v2 * J2ko
    tsd = (Pss - re * Qss) * invden
    rsd = (Qss - re * Pss) * invden
    tdo = (Pv - re * Qv) * invden
    rdo = (Qv - re * Pv) * invden
    g1 = (z - J1ks * too) * inv_Kp
    g2 = (z - J1ko * tss) * inv_kp
    T1, T2 = (v2 * g1) * u1, (v1 * g2) * u2
    T3 = (rdo * Qss + tdo * Pss) * rinf
    rsod = (T1 + T2 - T3) / omr2
    rso = w * laisum + rsod
    rsodt = ((tss + tsd) * tdo + (tsd + tss * rsrdd) * too) * rs_invdn
    rsot = rso + tsstoo * rs + rsodt
    rddt = rdd + tdd * tdd * rs_invdn
    rsdt = rsd + (tsd + tss) * tdd * rs_invdn
    rdot = rdo + tdd * (tdo + too) * rs_invdn
    return dict(rsot=rsot, rddt=rddt, rsdt=rsdt, rdot=rdot)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    P = lhs(n, 777)
    talf, t12, t21 = K.interface_constants('5', 40)
    np.random.seed(1)
    for variant in (2, 3):
        worst = dict(ks=0.0, kt=0.0, rsot=0.0, rddt=0.0, rsdt=0.0, rdot=0.0)
        worst_mixed = dict(worst)
        for i in range(n):
            p = {k: P[k][i] for k in P}
            lf, so, ca = O.prosail(p['N'], p['Cab'], p['Cxc'], p['Cbr'], p['Cw'], p['Cm'], p['lai'], p['hotspot'], 35.0, 30.0, 50.0,
                                   p['soil_refl'], p['soil_moist'], lidf_type='campbell', a=p['lidf_a'])
            refl, tran = leaf(f32, p, talf, t12, t21, variant)
            worst['ks'] = max(worst['ks'], np.abs(refl - lf['ks']).max())
            worst['kt'] = max(worst['kt'], np.abs(tran - lf['kt']).max())
            rs = so['ref'].astype(f32)
            c = canopy(f32, refl, tran, rs, ca, p['lai'], variant)
            for key in ('rsot', 'rddt', 'rsdt', 'rdot'):
                worst[key] = max(worst[key], np.abs(c[key] - ca[key]).max())
            # canopy alone in float32 on exact leaf spectra (isolates the 4SAIL conditioning)
            c = canopy(f32, lf['ks'].astype(f32), lf['kt'].astype(f32), rs, ca, p['lai'], variant)
            for key in ('rsot', 'rddt', 'rsdt', 'rdot'):
                worst_mixed[key] = max(worst_mixed[key], np.abs(c[key] - ca[key]).max())
        print("variant %d  fused : " % variant + "  ".join("%s %.2e" % kv for kv in worst.items()))
        print("variant %d  4SAIL : " % variant + "  ".join("%s %.2e" % kv for kv in worst_mixed.items() if kv[0][0] == 'r'))


def extremes():
    talf, t12, t21 = K.interface_constants('5', 40)
    for kw in (dict(N=3.0, Cab=300., Cxc=80., Cbr=5., Cw=0.4, Cm=0.2), dict(N=1.0, Cab=0.01, Cxc=0.0, Cbr=0.0, Cw=1e-6, Cm=1e-6),
               dict(N=2.5, Cab=0., Cxc=0., Cbr=0., Cw=0., Cm=0.), dict(N=2.5, Cab=0., Cxc=0., Cbr=0., Cw=1e-6, Cm=0.),
               dict(N=2.5, Cab=0., Cxc=0., Cbr=0., Cw=1e-5, Cm=0.), dict(N=2.5, Cab=0., Cxc=0., Cbr=0., Cw=1e-4, Cm=0.),
               dict(N=1.2, Cab=1., Cxc=0., Cbr=0., Cw=1e-4, Cm=1e-4)):
        o = O.prospect(**{k: np.float64(v) for k, v in kw.items()})
        for variant in (2, 3):
            refl, tran = leaf(f32, kw, talf, t12, t21, variant)
            print(kw, "variant", variant, "ks %.2e kt %.2e" % (np.nanmax(np.abs(refl - o['ks'])), np.nanmax(np.abs(tran - o['kt']))),
                  "nan", int(np.isnan(refl).sum()))


if __name__ == "__main__":
    extremes()
    main()


def err_vs_omt():
    """leaf error of variant 3 binned by 1 - tau (which sets how close the plate is to zero absorption)"""
    talf, t12, t21 = K.interface_constants('5', 40)
    rows = []
    for Cw in 10.0 ** np.linspace(-7, -2, 26):
        for N in (1.0, 1.5, 2.5, 3.0):
            kw = dict(N=N, Cab=0., Cxc=0., Cbr=0., Cw=Cw, Cm=0.)
            o = O.prospect(**{k: np.float64(v) for k, v in kw.items()})
            refl, tran = leaf(f32, kw, talf, t12, t21, 3)
            kall = Cw * K.lib.p5.Kw.astype(np.float64) / N
            err = np.maximum(np.abs(refl - o['ks']), np.abs(tran - o['kt']))
            rows.append((kall, err))
    kall = np.concatenate([r[0] for r in rows]); err = np.concatenate([r[1] for r in rows])
    for lo in range(-12, 1):
        m = (kall >= 10.0 ** lo) & (kall < 10.0 ** (lo + 1))
        if m.any():
            print("k in [1e%d, 1e%d): n %6d  max err %.2e" % (lo, lo + 1, m.sum(), err[m].max()))
