import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/libvc_oracle.so) -- the checker, never the product."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def vcb():
    import vcb200
    return vcb200


@pytest.fixture(scope="session")
def fixture_model():
    """The reference's real model test/models/clb_to_slt_gmm32_order40_diff.jld, re-stored as npz
    by tests/golden/make_golden.py (weights (32,), means (80,32), covars (80,80,32))."""
    z = np.load(os.path.join(GOLDEN, "gmm32_order40_diff.npz"))
    return (np.asfortranarray(z["weights"]), np.asfortranarray(z["means"]), np.asfortranarray(z["covars"]))


def tol_for(y_ref):
    """Parity bar of BASELINE.json: max abs error <= 1e-4 x feature scale (max |y_oracle|)."""
    return 1e-4 * float(np.abs(y_ref).max())
