import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.abspath(ROOT))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
warnings.filterwarnings("ignore", category=RuntimeWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def cases():
    """Inputs + outputs of the LIVE reference (tests/golden/make_golden.py)."""
    with np.load(os.path.join(GOLDEN, "ref_cases.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def fixtures():
    """The reference's own golden files for this path (prospect5_spectrum.txt, prospect_d_test.mat, REFL_CAN.txt, tav_alpha40.mat)."""
    with np.load(os.path.join(GOLDEN, "ref_fixtures.npz")) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def engine():
    import pyrism_b200
    return pyrism_b200.get_engine(0)
