"""The drop-in boundary from the caller's side: include/prosail_b200.h is a plain C header, and a C99 program linked against
libprosail_b200.so (no Python, no C++ on its side) drives the PROSPECT path through host pointers."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")
LIBDIR = os.path.join(ROOT, "pyrism_b200", "lib")
SRC = os.path.join(ROOT, "tests", "c_abi", "prospect_demo.c")


def _build(tmp_path):
    import __graft_entry__ as ge
    ge.build()
    exe = str(tmp_path / "prospect_demo")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-O1", "-I", INC, SRC, "-o", exe,
                           "-L", LIBDIR, "-lprosail_b200", "-Wl,-rpath," + LIBDIR])
    return exe


def test_header_is_plain_c_and_cxx(tmp_path):
    for cc, std, name in (("gcc", "-std=c99", "t.c"), ("g++", "-std=c++11", "t.cpp")):
        src = tmp_path / name
        src.write_text('#include "prosail_b200.h"\nint main(void) { return PROSAIL_NOUT == 23 && sizeof(prosail_out_t) > 0 ? 0 : 1; }\n')
        subprocess.check_call([cc, std, "-pedantic", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", INC, str(src)])


def test_c_program_links_and_reports_missing_gpu_without_computing(tmp_path):
    """No GPU in the CPU test environment: prosail_create returns PROSAIL_E_NOGPU and the program stops there (no fallback)."""
    import torch
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 2 and "usage" in r.stderr          # a device exists: the program wants its arguments
    else:
        assert r.returncode == 0 and r.stdout.startswith("NOGPU"), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_c_program_reproduces_the_python_drop_in(tmp_path):
    import pyrism_b200 as pb
    from pyrism_b200 import constants as K
    from oracle import prosail_oracle as orc
    exe = _build(tmp_path)
    rng = np.random.default_rng(11)
    n = 5
    P = np.stack([rng.uniform(1, 3, n), rng.uniform(5, 90, n), rng.uniform(1, 25, n), rng.uniform(0, 1, n), rng.uniform(0.001, 0.05, n),
                  rng.uniform(0.001, 0.02, n)])
    t = K.lib.p5
    talf, t12, t21 = K.interface_constants('5', 40)
    tables = tmp_path / "tables.bin"
    with open(tables, "wb") as f:
        for a in (t.Kab, t.Kxc, t.Kbr, t.Kw, t.Km):
            f.write(np.ascontiguousarray(a, dtype=np.float32).tobytes())
        for a in (talf, t12, t21):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(np.int32(n).tobytes())
        f.write(np.ascontiguousarray(P, dtype=np.float64).tobytes())
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(tables), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (r.returncode, r.stdout, r.stderr)
    got = np.fromfile(out, dtype=np.float64).reshape(2, n, 2101)
    leaf = pb.PROSPECT(*P)
    assert np.array_equal(got[0], leaf.ks) and np.array_equal(got[1], leaf.kt)          # same library, same bits
    for i in range(n):
        o = orc.prospect(*P[:, i])
        assert np.max(np.abs(got[0, i] - o['ks'])) <= 1e-9 and np.max(np.abs(got[1, i] - o['kt'])) <= 1e-9
