"""The oracle against the reference ITSELF (oracle/_ref/pyrism, copied from /root/reference/pyrism by __graft_entry__.build()),
and the CPU arms of bench.py.  Skipped where oracle/_ref is absent (a fresh clone before build()); the committed golden
vectors of tests/golden/ (tests/test_oracle.py) pin the same thing without the reference."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from oracle import prosail_oracle as orc      # noqa: E402
from oracle import ref_runner as rr           # noqa: E402

needs_ref = pytest.mark.skipif(not rr.available(), reason="oracle/_ref/pyrism not built (run __graft_entry__.build() where /root/reference is mounted)")


@needs_ref
def test_reference_copy_is_the_reference():
    """oracle/_ref holds the reference package as it is: same files, byte for byte, when /root/reference is mounted."""
    src = "/root/reference/pyrism"
    if not os.path.isdir(src):
        pytest.skip("/root/reference is not mounted here")
    for dp, _, files in os.walk(src):
        for f in files:
            if f.endswith(".pyc"):
                continue
            a = os.path.join(dp, f)
            b = os.path.join(rr.REF_DIR, "pyrism", os.path.relpath(a, src))
            assert os.path.isfile(b), b
            assert open(a, "rb").read() == open(b, "rb").read(), b


@needs_ref
@pytest.mark.parametrize("lidf", ["campbell", "verhoef"])
def test_oracle_equals_live_reference(lidf):
    """max |oracle - reference| = 0.0 on random parameter sets, every quantity of the path (the pin of SURVEY.md 8(c),
    re-checked wherever the reference copy travelled -- here and on the GPU box)."""
    rng = np.random.default_rng(11 if lidf == "campbell" else 12)
    for _ in range(6):
        N, Cab, Cxc, Cbr = rng.uniform(1, 3), rng.uniform(5, 90), rng.uniform(1, 25), rng.uniform(0, 1)
        Cw, Cm, lai, hot = rng.uniform(0.001, 0.05), rng.uniform(0.001, 0.02), rng.uniform(0.1, 8), rng.uniform(0.01, 0.5)
        sr, sm = rng.uniform(0.3, 1.2), rng.uniform(0, 1)
        iza, vza, raa = rng.uniform(0, 70), rng.uniform(0, 60), rng.uniform(0, 180)
        a, b = (rng.uniform(20, 80), 0.0) if lidf == "campbell" else (rng.uniform(-0.6, 0.6), rng.uniform(-0.35, 0.35))
        leaf, soil, can = orc.prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, sr, sm, lidf_type=lidf, a=a, b=b)
        rl, rs, rc = rr.prosail(N, Cab, Cxc, Cbr, Cw, Cm, lai, hot, iza, vza, raa, sr, sm, lidf_type=lidf, a=a, b=b)
        assert np.array_equal(leaf["ks"], rl.ks) and np.array_equal(leaf["kt"], rl.kt)
        assert np.array_equal(soil["ref"], rs.ref)
        for q, ref in (("rsot", rc.BRF.ref), ("rddt", rc.BHR.ref), ("rsdt", rc.DHR.ref), ("rdot", rc.HDR.ref), ("ke", rc.ke)):
            assert np.array_equal(can[q], ref), q
        for i in range(2, 8):
            assert can["L8"]["BRF"]["B%d" % i] == getattr(rc.BRF.L8, "B%d" % i)
        for i in range(1, 10):
            assert can["ASTER"]["BRF"]["B%d" % i] == getattr(rc.BRF.ASTER, "B%d" % i)


def _bench(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, cwd="/tmp", timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads(out.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("cfg", [2, 1])
def test_bench_reference_arm_contract(cfg):
    """bench.py --impl reference: the contract keys, the same metric / unit as the GPU arm, and kind == "reference" whenever the
    reference copy is present (the port only where it is not)."""
    d = _bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--config", str(cfg))
    assert d["impl"] == "reference" and d["unit"] == "spectra/s" and d["metric"].startswith("PROSAIL spectra/s")
    assert d["value"] > 0 and d["higher_is_better"] is True and d["gpu_launches"] == 0
    assert d["config"]["index"] == cfg and d["config"]["workload"].startswith("BASELINE configs[%d]" % cfg)
    assert d["cpu_baseline"]["kind"] == ("reference" if rr.available() else "port")
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["e2e"] == {"value": d["value"], "unit": "spectra/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
