#!/usr/bin/env python
"""Edge-of-domain inputs through the CUDA path (fp64 and fp32 mode) against the oracle: extreme lai / hotspot / angles / LIDF
parameters / pigment-free leaves / dark and bright soils.  Prints one line per case; test infrastructure (imports oracle/)."""
import os, sys, warnings
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
warnings.filterwarnings("ignore")
import pyrism_b200 as pb
from oracle import prosail_oracle as orc

base = dict(N=1.5, Cab=40.0, Cxc=8.0, Cbr=0.1, Cw=0.01, Cm=0.009, lai=3.0, hotspot=0.1, iza=35.0, vza=30.0, raa=50.0, sr=1.0, sm=0.5,
            lidf_type='campbell', a=57.0, b=0.0)
cases = [("base", {}), ("lai 50", dict(lai=50.0)), ("lai 1e-6", dict(lai=1e-6)), ("lai 300", dict(lai=300.0)), ("hotspot 1e-9", dict(hotspot=1e-9)),
         ("hotspot 10", dict(hotspot=10.0)), ("iza 89.9", dict(iza=89.9)), ("vza 89", dict(vza=89.0)), ("iza 0 vza 0", dict(iza=0.0, vza=0.0, raa=0.0)),
         ("raa 360", dict(raa=360.0)), ("raa 180 hot", dict(iza=30.0, vza=30.0, raa=180.0)), ("vza -20", dict(vza=-20.0)), ("campbell a 1", dict(a=1.0)),
         ("campbell a 89.9", dict(a=89.9)), ("verhoef -1 0", dict(lidf_type='verhoef', a=-1.0, b=0.0)), ("verhoef .99 0", dict(lidf_type='verhoef', a=0.99, b=0.0)),
         ("verhoef 0 1", dict(lidf_type='verhoef', a=0.0, b=-1.0)), ("N 1", dict(N=1.0)), ("N 6", dict(N=6.0)), ("no pigments", dict(Cab=0.0, Cxc=0.0, Cbr=0.0)),
         ("dry thin leaf", dict(Cw=1e-5, Cm=1e-4)), ("thick leaf", dict(Cw=0.2, Cm=0.1, Cab=200.0)), ("black soil", dict(sr=0.0)), ("bright wet soil", dict(sr=2.5, sm=1.0)),
         ("dry soil", dict(sm=0.0))]
worst64 = worst32 = 0.0
for name, over in cases:
    p = dict(base, **over)
    try:
        _, _, o = orc.prosail(p['N'], p['Cab'], p['Cxc'], p['Cbr'], p['Cw'], p['Cm'], p['lai'], p['hotspot'], p['iza'], p['vza'], p['raa'], p['sr'], p['sm'],
                              lidf_type=p['lidf_type'], a=p['a'], b=p['b'])
    except Exception as e:
        print("%-16s oracle raises %s" % (name, type(e).__name__))
        continue
    line = "%-16s" % name
    for dt in (np.float64, np.float32):
        m = pb.PROSAIL(N=p['N'], Cab=p['Cab'], Cxc=p['Cxc'], Cbr=p['Cbr'], Cw=p['Cw'], Cm=p['Cm'], lai=p['lai'], hotspot=p['hotspot'], iza=p['iza'], vza=p['vza'],
                       raa=p['raa'], soil_reflectance=p['sr'], soil_moisture=p['sm'], lidf_type=p['lidf_type'], a=p['a'], b=p['b'], dtype=dt)
        errs = []
        for got, key in ((m.BRF.ref, 'rsot'), (m.BHR.ref, 'rddt'), (m.DHR.ref, 'rsdt'), (m.HDR.ref, 'rdot')):
            ref = np.asarray(o[key], dtype=np.float64) * np.ones(2101)
            both_nan = np.isnan(got) & np.isnan(ref)
            d = np.abs(got.astype(np.float64) - ref)
            d[both_nan] = 0.0
            errs.append(float(np.nanmax(d)) if not np.all(np.isnan(d)) else float('nan'))
            if np.any(np.isnan(got) != np.isnan(ref)):
                errs[-1] = float('nan')
        e = max(errs) if not any(np.isnan(errs)) else float('nan')
        line += "  %s max|diff| %.2e (oracle BRF range %.3g..%.3g)" % (np.dtype(dt).name, e, np.nanmin(o['rsot']), np.nanmax(o['rsot']))
        if dt == np.float64:
            worst64 = max(worst64, e) if e == e else float('nan')
        else:
            worst32 = max(worst32, e) if e == e else float('nan')
    print(line)
print("worst fp64 %.3e   worst fp32 %.3e" % (worst64, worst32))
