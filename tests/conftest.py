import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden", "reference_cases.npz")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(GOLDEN)


def variant_params(variant, h=10):
    """Parameter variants used by oracle/gen_golden.py (kept in sync by hand)."""
    import numpy as np
    from oracle import reference_mpc as rm
    mpc, biped = rm.MPCParams(h=h), rm.BipedParams()
    if variant == 1:
        mpc.x_cmd = np.array([0, 0, 0, 0, 0, 0.55, 0.2, 0, 0, 0.3, -0.1, 0], dtype=float)
    elif variant == 2:
        biped.mu = 0.7
        biped.f_min = np.array([[-120.0], [-80.0], [0.0]])
        biped.f_max = np.array([[300.0], [300.0], [400.0]])
        biped.tau_max = np.array([[4.0], [50.0], [20.0]])
        biped.tau_min = np.array([[-3.0], [-45.0], [-20.0]])
        biped.m = 13.5
        biped.I = np.array([[0.9, 0.02, 0.01], [0.02, 0.95, -0.015], [0.01, -0.015, 0.08]])
        mpc.Q = np.array([400, 150, 120, 250, 320, 650, 2, 1.5, 1, 1, 2, 1, 1], dtype=float)
        mpc.R = np.array([1, 2, 1, 1, 2, 1, 3, 1, 2, 3, 1, 2], dtype=float) * 1e-4
        mpc.kv = 0.02
    return mpc, biped
