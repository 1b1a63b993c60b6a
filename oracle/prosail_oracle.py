# -*- coding: utf-8 -*-
"""CPU oracle for the PROSPECT-5/D -> LSM -> 4SAIL -> band-mean path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``pyrism_b200/`` imports this module;
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may.  The product path has no CPU
fallback and does not route through this file.

What it is: a numpy/scipy restatement of the reference algorithm
(ibaris/pyrism 0.0.5, ``pyrism/models/models.py``), one parameter set and one
geometry per call, vectorised over the 2101 wavelengths only -- the same
granularity as the reference.  Every function cites the reference lines it
follows.  The arithmetic order is kept so that results are bit-comparable
with the reference when it is fed ``np.float64`` scalars.

Third-party arithmetic on the path: ``scipy.special.expi`` (reference call
site models.py:955; the reference leaves scipy unpinned, this image has scipy
1.18.x) and numpy's ufuncs (numpy 2.3.x; NEP-50 promotion rules are relied on
for the float32 intermediates in :func:`calctav`).

Parity pin: ``tests/golden/make_golden.py`` runs the *live* reference (from
/root/reference, with the 2-line ``scipy.misc.factorial`` shim) and this
oracle on the same inputs and stores the reference outputs as fixtures;
``tests/test_oracle.py`` checks this module against those fixtures, against
the reference's own golden files (prospect5_spectrum.txt,
prospect_d_test.mat, REFL_CAN.txt) and against the known-answer constants of
the reference's tests/test_volscat.py.
"""
from __future__ import division

import os

import numpy as np
from scipy.special import expi

NBANDS = 2101
WAVELENGTHS = np.arange(400, 2501)

# Inclusive integer-nm windows.  L8: models.py:836-841; ASTER: models.py:801-809.
L8_WINDOWS = (("B2", 452, 512), ("B3", 533, 590), ("B4", 636, 673),
              ("B5", 851, 879), ("B6", 1566, 1651), ("B7", 2107, 2294))
ASTER_WINDOWS = (("B1", 520, 600), ("B2", 630, 690), ("B3", 760, 860), ("B4", 1600, 1700),
                 ("B5", 2145, 2185), ("B6", 2185, 2225), ("B7", 2235, 2285),
                 ("B8", 2295, 2365), ("B9", 2360, 2430))

_TABLES = None


def load_tables(path=None):
    """float32 coefficient tables, as the reference loads them (library.py:43-64)."""
    global _TABLES
    if _TABLES is None or path is not None:
        if path is None:
            path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pyrism_b200", "data", "tables.npz")
        with np.load(path) as z:
            _TABLES = {k: z[k] for k in z.files}
    return _TABLES


# ----------------------------------------------------------------------------------------------
# PROSPECT
# ----------------------------------------------------------------------------------------------
def calctav(alpha, KN):
    """Average interface transmissivity tav(alpha, n).  Reference: models.py:961-998.

    ``KN`` must be the float32 table: n2, npx, nm, a, k and every expression
    built only from them stay float32 under numpy>=2 promotion, exactly as in
    the reference.
    """
    assert KN.dtype == np.float32
    n2 = KN * KN                                   # :972  float32
    npx = n2 + 1                                   # :973  float32
    nm = n2 - 1                                    # :974  float32
    a = (KN + 1) * (KN + 1) / 2.                   # :975  float32
    k = -(n2 - 1) * (n2 - 1) / 4.                  # :976  float32
    sa = np.sin(np.deg2rad(alpha))                 # :977  float64 scalar

    if alpha != 90:                                # :979-982
        b1 = np.sqrt((sa * sa - npx / 2) * (sa * sa - npx / 2) + k)
    else:
        b1 = 0.
    b2 = sa * sa - npx / 2                         # :983  float64
    b = b1 - b2                                    # :984
    b3 = b ** 3
    a3 = a ** 3                                    # float32
    ts = (k ** 2 / (6 * b3) + k / b - b / 2) - (k ** 2. / (6 * a3) + k / a - a / 2)   # :987

    tp1 = -2 * n2 * (b - a) / (npx ** 2)                                               # :989
    tp2 = -2 * n2 * npx * np.log(b / a) / (nm ** 2)                                    # :990
    tp3 = n2 * (1 / b - 1 / a) / 2                                                     # :991
    tp4 = 16 * n2 ** 2 * (n2 ** 2 + 1) * np.log((2 * npx * b - nm ** 2) / (2 * npx * a - nm ** 2)) / (
        npx ** 3 * nm ** 2)                                                            # :992-993
    tp5 = 16 * n2 ** 3 * (1. / (2 * npx * b - nm ** 2) - 1 / (2 * npx * a - nm ** 2)) / (npx ** 3)   # :994
    tp = tp1 + tp2 + tp3 + tp4 + tp5
    return (ts + tp) / (2 * sa ** 2)                                                   # :996


def band_window_means(spec, windows):
    """Unweighted inclusive boxcar means (models.py:811-823, 843-851, 1101-1186).

    ``spec`` is [2101] or [2101, C]; like the reference, rows are selected with a
    boolean mask and each column is averaged on its own with ``.mean()``.
    """
    out = {}
    for name, lo, hi in windows:
        rows = spec[(WAVELENGTHS >= lo) & (WAVELENGTHS <= hi)]
        if rows.ndim == 1:
            out[name] = rows.mean()
        else:
            out[name] = np.array([rows[:, c].mean() for c in range(rows.shape[1])], dtype=rows.dtype)
    return out


def prospect(N, Cab, Cxc, Cbr, Cw, Cm, Can=0, alpha=40, version='5'):
    """Leaf plate model.  Reference: models.py:900-1072 (+ band means 1074-1195).

    Returns a dict with ks, kt, ka, ke, om (float64[2101]), the one-layer terms
    r, t, Ra, Ta, denom, the float32 table ``int`` [2101, 6] and the float32
    L8 / ASTER means of (ks, kt, ka, ke, om).
    """
    if version != '5' and version != 'D':                                   # :915-917
        raise ValueError("version must be '5' or 'D'")
    if version == 'D' and Can == 0:                                         # :925-926
        raise AssertionError("PROSPECT-D needs Can != 0")
    # Oracle convention: the float64 path of the reference, i.e. np.float64 scalars.
    N, Cab, Cxc, Cbr, Cw, Cm, Can = (np.float64(v) for v in (N, Cab, Cxc, Cbr, Cw, Cm, Can))
    T = load_tables()
    p = 'p5_' if version == '5' else 'pd_'
    KN, Kab, Kxc, Kbr, Kw, Km = (T[p + n] for n in ('KN', 'Kab', 'Kxc', 'Kbr', 'Kw', 'Km'))
    Kan = np.zeros_like(Km) if version == '5' else T['pd_Kan']             # :935, :944

    # :950-951 (left-to-right sum, then / N)
    kall = (Cab * Kab + Cxc * Kxc + Can * Kan + Cbr * Kbr + Cw * Kw + Cm * Km) / N
    j = kall > 0                                                            # :953
    t1 = (1 - kall) * np.exp(-kall)                                         # :954
    with np.errstate(invalid='ignore', divide='ignore'):
        t2 = kall ** 2 * (-expi(-kall))                                     # :955
    tau = np.ones_like(t1)
    tau[j] = t1[j] + t2[j]                                                  # :957

    # one layer, models.py:1011-1025
    talf = calctav(alpha, KN)
    ralf = 1.0 - talf
    t12 = calctav(90, KN)
    r12 = 1. - t12
    t21 = t12 / (KN * KN)
    r21 = 1 - t21
    denom = 1. - r21 * r21 * tau * tau
    Ta = talf * tau * t21 / denom
    Ra = ralf + r21 * tau * Ta
    t = t12 * tau * t21 / denom
    r = r12 + r21 * tau * t

    # N layers (Stokes), models.py:1043-1068
    D = np.sqrt((1 + r + t) * (1 + r - t) * (1. - r + t) * (1. - r - t))
    rq = r * r
    tq = t * t
    a = (1 + rq - tq + D) / (2 * r)
    b = (1 - rq + tq + D) / (2 * t)
    bNm1 = np.power(b, N - 1)
    bN2 = bNm1 * bNm1
    a2 = a * a
    den = a2 * bN2 - 1
    Rsub = a * (bN2 - 1) / den
    Tsub = bNm1 * (a2 - 1) / den
    jz = r + t >= 1.                                                        # :1057-1059
    Tsub[jz] = t[jz] / (t[jz] + (1 - t[jz]) * (N - 1))
    Rsub[jz] = 1 - Tsub[jz]
    den2 = 1 - Rsub * r
    kt = Ta * Tsub / den2
    ks = Ra + Ta * Rsub * t / den2
    ka = 1 - ks - kt
    ke = ks + ka
    om = ks / ke

    tab = np.asarray([WAVELENGTHS, ks, kt, ka, ke, om], dtype=np.float32).transpose()   # :1070-1072
    five = tab[:, 1:6]
    return dict(ks=ks, kt=kt, ka=ka, ke=ke, om=om, r=r, t=t, Ra=Ra, Ta=Ta, denom=denom, tau=tau, kall=kall,
                int=tab, L8=band_window_means(five, L8_WINDOWS), ASTER=band_window_means(five, ASTER_WINDOWS))


# ----------------------------------------------------------------------------------------------
# LSM
# ----------------------------------------------------------------------------------------------
def lsm(reflectance, moisture):
    """Lambertian soil mix.  Reference: models.py:1916-1920 (+ float32 band means 1925-2006)."""
    reflectance, moisture = np.float64(reflectance), np.float64(moisture)
    T = load_tables()
    ref = reflectance * (moisture * T['rsoil1'] + (1 - moisture) * T['rsoil2'])    # :1917
    f32 = np.asarray(ref, dtype=np.float32)
    return dict(ref=ref, L8=band_window_means(f32, L8_WINDOWS), ASTER=band_window_means(f32, ASTER_WINDOWS))


# ----------------------------------------------------------------------------------------------
# Angles, LIDF, volume scattering
# ----------------------------------------------------------------------------------------------
def kernel_angles(iza, vza, raa, angle_unit='DEG'):
    """Angle normalisation for ONE geometry.  Reference: core/_core.py:116-173.

    Returns (iza, vza, raa) in radians and (izaDeg, vzaDeg, raaDeg).  Negative
    zeniths are flipped and pi is added to raa.  The *Deg attributes keep the
    user's values in 'DEG' mode (the flip happens after they are stored,
    _core.py:123-146) and are derived from the flipped radians in 'RAD' mode
    (_core.py:165-173).
    """
    if angle_unit not in ('DEG', 'RAD'):
        raise AssertionError("angle_unit must be 'DEG' or 'RAD'")
    iza, vza, raa = np.float64(iza), np.float64(vza), np.float64(raa)
    if angle_unit == 'DEG':
        izaDeg, vzaDeg, raaDeg = iza, vza, raa
        iza, vza, raa = izaDeg * np.pi / 180.0, vzaDeg * np.pi / 180.0, raaDeg * np.pi / 180.0   # auxiliary.py:182-192
    if vza < 0:
        vza = -vza
        raa = raa + np.pi
    if iza < 0:
        iza = -iza
        raa = raa + np.pi
    if angle_unit == 'RAD':
        izaDeg, vzaDeg, raaDeg = iza * 180. / np.pi, vza * 180. / np.pi, raa * 180. / np.pi
    return iza, vza, raa, izaDeg, vzaDeg, raaDeg


def lidf_campbell(a, n_elements=18):
    """Ellipsoidal LIDF.  Reference: models.py:283-332."""
    alpha = float(a)
    excent = np.exp(-1.6184e-5 * alpha ** 3. + 2.1145e-3 * alpha ** 2. - 1.2390e-1 * alpha + 3.2491)
    freq = np.zeros(n_elements)
    step = 90.0 / n_elements
    for i in range(n_elements):
        tl1 = (i * step) * np.pi / 180.0
        tl2 = ((i + 1.) * step) * np.pi / 180.0
        x1 = excent / (np.sqrt(1. + excent ** 2. * np.tan(tl1) ** 2.))
        x2 = excent / (np.sqrt(1. + excent ** 2. * np.tan(tl2) ** 2.))
        if excent == 1.:
            freq[i] = abs(np.cos(tl1) - np.cos(tl2))
        else:
            alph = excent / np.sqrt(abs(1. - excent ** 2.))
            alph2 = alph ** 2.
            x12 = x1 ** 2.
            x22 = x2 ** 2.
            if excent > 1.:
                alpx1 = np.sqrt(alph2 + x12)
                alpx2 = np.sqrt(alph2 + x22)
                dum = x1 * alpx1 + alph2 * np.log(x1 + alpx1)
                freq[i] = abs(dum - (x2 * alpx2 + alph2 * np.log(x2 + alpx2)))
            else:
                almx1 = np.sqrt(alph2 - x12)
                almx2 = np.sqrt(alph2 - x22)
                dum = x1 * almx1 + alph2 * np.arcsin(x1 / alph)
                freq[i] = abs(dum - (x2 * almx2 + alph2 * np.arcsin(x2 / alph)))
    sum0 = np.sum(freq)
    lidf = np.zeros(n_elements)
    for i in range(n_elements):
        lidf[i] = freq[i] / sum0
    return lidf


def lidf_verhoef(a, b, n_elements=18):
    """Bimodal LIDF (fixed-point iteration).  Reference: models.py:335-392."""
    freq = 1.0
    step = 90.0 / n_elements
    lidf = np.zeros(n_elements) * 1.
    angles = (np.arange(n_elements) * step)[::-1]
    i = 0
    for angle in angles:
        tl1 = np.radians(angle)
        if a > 1.0:
            f = 1.0 - np.cos(tl1)
        else:
            eps = 1e-8
            delx = 1.0
            x = 2.0 * tl1
            p = float(x)
            while delx >= eps:
                y = a * np.sin(x) + .5 * b * np.sin(2. * x)
                dx = .5 * (y - x + p)
                x = x + dx
                delx = abs(dx)
            f = (2. * y + p) / np.pi
        freq = freq - f
        lidf[i] = freq
        freq = float(f)
        i += 1
    return lidf[::-1]


def volscatt_volume(iza, vza, raa, lza):
    """chi_s, chi_o, frho, ftau for one leaf zenith ``lza`` [deg].  Reference: models.py:177-258."""
    cts = np.cos(iza)
    cto = np.cos(vza)
    sts = np.sin(iza)
    sto = np.sin(vza)
    cospsi = np.cos(raa)
    psir = raa
    clza = np.cos(np.radians(lza))
    slza = np.sin(np.radians(lza))
    cs = clza * cts
    co = clza * cto
    ss = slza * sts
    so = slza * sto
    cosbts = 5.
    if np.abs(ss) > 1e-6:
        cosbts = -cs / ss
    cosbto = 5.
    if np.abs(so) > 1e-6:
        cosbto = -co / so
    if np.abs(cosbts) < 1.0:
        bts = np.arccos(cosbts)
        ds = ss
    else:
        bts = np.pi
        ds = cs
    chi_s = 2. / np.pi * ((bts - np.pi * 0.5) * cs + np.sin(bts) * ss)
    if abs(cosbto) < 1.0:
        bto = np.arccos(cosbto)
        do_ = so
    else:
        if vza < 90. * np.pi / 180.0:
            bto = np.pi
            do_ = co
        else:
            bto = 0.0
            do_ = -co
    chi_o = 2.0 / np.pi * ((bto - np.pi * 0.5) * co + np.sin(bto) * so)
    btran1 = np.abs(bts - bto)
    btran2 = np.pi - np.abs(bts + bto - np.pi)
    if psir <= btran1:
        bt1, bt2, bt3 = psir, btran1, btran2
    else:
        bt1 = btran1
        if psir <= btran2:
            bt2, bt3 = psir, btran2
        else:
            bt2, bt3 = btran2, psir
    t1 = 2. * cs * co + ss * so * cospsi
    t2 = 0.
    if bt2 > 0.:
        t2 = np.sin(bt2) * (2. * ds * do_ + ss * so * np.cos(bt1) * np.cos(bt3))
    denom = 2. * np.pi ** 2
    frho = ((np.pi - bt2) * t1 + t2) / denom
    ftau = (-bt2 * t1 + t2) / denom
    if frho < 0.:
        frho = 0.
    if ftau < 0.:
        ftau = 0.
    return chi_s, chi_o, frho, ftau


def volscatt_coef(iza, vza, raa, lidf):
    """LIDF-weighted sums ks, ko, bf, Fs, Ft (angles in radians).  Reference: models.py:144-175."""
    ks = ko = bf = Fs = Ft = 0.
    n_angles = len(lidf)
    angle_step = float(90.0 / n_angles)
    litab = np.arange(n_angles) * angle_step + (angle_step * 0.5)
    for i, ili in enumerate(litab):
        ttl = 1. * ili
        cttl = np.cos(np.radians(ttl))
        chi_s, chi_o, frho, ftau = volscatt_volume(iza, vza, raa, ttl)
        ksli = chi_s / np.cos(iza)
        koli = chi_o / np.cos(vza)
        sobli = frho * np.pi / (np.cos(iza) * np.cos(vza))
        sofli = ftau * np.pi / (np.cos(iza) * np.cos(vza))
        bfli = cttl ** 2.
        w = float(lidf[i])
        ks += ksli * w
        ko += koli * w
        bf += bfli * w
        Fs += sobli * w
        Ft += sofli * w
    return ks, ko, bf, Fs, Ft


# ----------------------------------------------------------------------------------------------
# 4SAIL
# ----------------------------------------------------------------------------------------------
def _jfunc1(k, l, t):
    """Reference: models.py:765-785 (array branch)."""
    del_ = (k - l) * t
    big = np.abs(del_) > 1e-3
    result = np.zeros(len(l))
    result[big] = (np.exp(-l[big] * t) - np.exp(-k * t)) / (k - l[big])
    sm = ~big
    result[sm] = 0.5 * t * (np.exp(-k * t) + np.exp(-l[sm] * t)) * (1. - (del_[sm] ** 2.) / 12.)
    return result


def _jfunc2(k, l, t):
    """Reference: models.py:787-789."""
    return (1. - np.exp(-(k + l) * t)) / (k + l)


def _hotspot(alf, lai, ko, ks):
    """20-step exponential Simpson integral.  Reference: models.py:741-763."""
    fhot = lai * np.sqrt(ko * ks)
    x1 = 0.
    y1 = 0.
    f1 = 1.
    fint = (1. - np.exp(-alf)) * .05
    sumint = 0.
    with np.errstate(invalid='ignore', divide='ignore'):
        for istep in range(1, 21):
            if istep < 20:
                x2 = -np.log(1. - istep * fint) / alf
            else:
                x2 = 1.
            y2 = -(ko + ks) * lai * x2 + fhot * (1. - np.exp(-alf * x2)) / alf
            f2 = np.exp(y2)
            sumint = sumint + (f2 - f1) * (x2 - x1) / (y2 - y1)
            x1, y1, f1 = x2, y2, f2
    tsstoo = f1
    if np.isnan(sumint):
        sumint = 0.
    return tsstoo, sumint


def sail(iza, vza, raa, ks, kt, lai, hotspot, rho_surface, lidf_type='campbell', a=57, b=0, angle_unit='DEG'):
    """4SAIL for one parameter set and one geometry.  Reference: models.py:528-729.

    ``ks``/``kt`` are the leaf reflectance/transmittance spectra, ``rho_surface`` the soil spectrum
    (each float64[2101]).  Returns rsot/rddt/rsdt/rdot (BRF/BHR/DHR/HDR), the canopy-only terms,
    tss/too/tsstoo, ke (= m), the VolScatt sums, and float64 L8/ASTER means of the five products.
    """
    for name, v in (("ks", ks), ("kt", kt), ("rho_surface", rho_surface)):          # :534-547
        if len(v) != NBANDS:
            raise AssertionError("%s must have 2101 values" % name)
    iza_r, vza_r, raa_r, izaDeg, vzaDeg, raaDeg = kernel_angles(iza, vza, raa, angle_unit)
    if lidf_type == 'verhoef':                                                      # :560-565
        lidf = lidf_verhoef(a, b)
    elif lidf_type == 'campbell':
        lidf = lidf_campbell(a)
    else:
        raise AssertionError("The lidf_type must be 'verhoef' or 'campbell'")
    vk, vK, bf, Fs, Ft = volscatt_coef(iza_r, vza_r, raa_r, lidf)

    sdb = 0.5 * (vk + bf)                                                           # :583-588
    sdf = 0.5 * (vk - bf)
    dob = 0.5 * (vK + bf)
    dof = 0.5 * (vK - bf)
    ddb = 0.5 * (1.0 + bf)
    ddf = 0.5 * (1.0 - bf)
    sigb = ddb * ks + ddf * kt                                                      # :590-595
    sigf = ddf * ks + ddb * kt
    sigf[sigf == 0.0] = 1.e-36
    sigb[sigb == 0.0] = 1.0e-36
    att = 1. - sigf
    m = np.sqrt(att ** 2. - sigb ** 2.)                                             # :601
    sb = sdb * ks + sdf * kt
    sf = sdf * ks + sdb * kt
    vb = dob * ks + dof * kt
    vf = dof * ks + dob * kt
    w = Fs * ks + Ft * kt

    out = dict(vol_ks=vk, vol_ko=vK, bf=bf, Fs=Fs, Ft=Ft, lidf=lidf, ke=m)
    if lai <= 0:                                                                    # :609-634
        one = np.float64(1.0)
        out.update(tss=one, too=one, tsstoo=one, rdd=0, tdd=1, rsd=0, tsd=0, rdo=0, tdo=0, rso=0,
                   rddt=rho_surface, rsdt=rho_surface, rdot=rho_surface, rsot=rho_surface)
    else:
        e1 = np.exp(-m * lai)                                                       # :637-642
        e2 = e1 ** 2.
        rinf = (att - m) / sigb
        rinf2 = rinf ** 2.
        re = rinf * e1
        denom = 1. - rinf2 * e2
        J1ks = _jfunc1(vk, m, lai)
        J2ks = _jfunc2(vk, m, lai)
        J1ko = _jfunc1(vK, m, lai)
        J2ko = _jfunc2(vK, m, lai)
        Pss = (sf + sb * rinf) * J1ks                                               # :649-652
        Qss = (sf * rinf + sb) * J2ks
        Pv = (vf + vb * rinf) * J1ko
        Qv = (vf * rinf + vb) * J2ko
        tdd = (1. - rinf2) * e1 / denom                                             # :654-659
        rdd = rinf * (1. - e2) / denom
        tsd = (Pss - re * Qss) / denom
        rsd = (Qss - re * Pss) / denom
        tdo = (Pv - re * Qv) / denom
        rdo = (Qv - re * Pv) / denom
        tss = np.exp(-vk * lai)                                                     # :664-666
        too = np.exp(-vK * lai)
        z = _jfunc2(vk, vK, lai)
        g1 = (z - J1ks * too) / (vK + m)                                            # :668-669
        g2 = (z - J1ko * tss) / (vk + m)
        Tv1 = (vf * rinf + vb) * g1                                                 # :671-678
        Tv2 = (vf + vb * rinf) * g2
        T1 = Tv1 * (sf + sb * rinf)
        T2 = Tv2 * (sf * rinf + sb)
        T3 = (rdo * Qss + tdo * Pss) * rinf
        rsod = (T1 + T2 - T3) / (1. - rinf2)

        alf = 1e36                                                                  # :687-694
        tants = np.tan(np.radians(izaDeg))                                          # :731-739
        tanto = np.tan(np.radians(vzaDeg))
        cospsi = np.cos(np.radians(raaDeg))
        dso = np.sqrt(tants ** 2. + tanto ** 2. - 2. * tants * tanto * cospsi)
        if hotspot > 0.:
            alf = (dso / hotspot) * 2. / (vk + vK)
        if alf == 0.:                                                               # :696-702
            tsstoo = tss
            sumint = (1. - tss) / (vk * lai)
        else:
            tsstoo, sumint = _hotspot(alf, lai, vK, vk)
        rsos = w * lai * sumint                                                     # :706-710
        rso = rsos + rsod
        dn = 1. - rho_surface * rdd                                                 # :714-719
        dn[dn < 1e-36] = 1e-36
        rddt = rdd + tdd * rho_surface * tdd / dn                                   # :721-726
        rsdt = rsd + (tsd + tss) * rho_surface * tdd / dn
        rdot = rdo + tdd * rho_surface * (tdo + too) / dn
        rsodt = ((tss + tsd) * tdo + (tsd + tss * rho_surface * rdd) * too) * rho_surface / dn
        rsost = rso + tsstoo * rho_surface
        rsot = rsost + rsodt
        out.update(tss=tss, too=too, tsstoo=tsstoo, rdd=rdd, tdd=tdd, rsd=rsd, tsd=tsd, rdo=rdo, tdo=tdo,
                   rso=rso, rsos=rsos, rsod=rsod, rddt=rddt, rsdt=rsdt, rdot=rdot, rsot=rsot,
                   sumint=sumint, alf=alf, dso=dso)

    # products and their float64 band means, models.py:575-580
    prods = dict(BRF=out['rsot'], BRDF=out['rsot'] / np.pi, BHR=out['rddt'], DHR=out['rsdt'], HDR=out['rdot'])
    out['L8'] = {k: band_window_means(v, L8_WINDOWS) for k, v in prods.items()}
    out['ASTER'] = {k: band_window_means(v, ASTER_WINDOWS) for k, v in prods.items()}
    return out


def prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hotspot, iza, vza, raa, soil_reflectance, soil_moisture,
            Can=0, alpha=40, version='5', lidf_type='campbell', a=57, b=0):
    """PROSPECT -> LSM -> SAIL for one sample, as examples/prospect_sail.py:9-28 chains them."""
    leaf = prospect(N, Cab, Cxc, Cbr, Cw, Cm, Can=Can, alpha=alpha, version=version)
    soil = lsm(soil_reflectance, soil_moisture)
    can = sail(iza, vza, raa, leaf['ks'], leaf['kt'], lai, hotspot, soil['ref'], lidf_type=lidf_type, a=a, b=b)
    return leaf, soil, can
