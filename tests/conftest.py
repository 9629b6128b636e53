import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_GRAPHS = ["ba2000", "rmat10", "gnp600d", "weighted300"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import scipy.sparse as sp
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    n = len(z["indptr"]) - 1
    A = sp.csr_matrix((z["data"], z["indices"], z["indptr"]), shape=(n, n))
    return z, A, bool(z["directed"])


@pytest.fixture(params=GOLDEN_GRAPHS)
def golden(request):
    return (request.param,) + load_golden(request.param)


def rel_l1(a, b):
    """Relative L1 distance, the tolerance metric of BASELINE.json's north_star."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.abs(b).sum()
    return float(np.abs(a - b).sum() / (denom if denom != 0 else 1.0))
