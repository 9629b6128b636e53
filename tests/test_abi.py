"""CPU checks: the C-ABI library loads and exports every symbol include/pgb200.h declares, the
host-side mirror is importable without a GPU, and the product path fails loudly (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "pgb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pgb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import pygrank_b200
    from pygrank_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        pygrank_b200.build()
    handle = ctypes.CDLL(_capi.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(handle, name), f"{name} declared in pgb200.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == declared, "ctypes signatures out of sync with the header"
    handle.pgb_abi_version.restype = ctypes.c_int
    assert handle.pgb_abi_version() == _capi.ABI_VERSION == 4
    handle.pgb_tile_items.restype = ctypes.c_int
    assert handle.pgb_tile_items() % 2 == 0 and (handle.pgb_tile_items() // 256) % 2 == 1   # odd items/thread


def test_sass_is_sm100a_only():
    from pygrank_b200 import _capi
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import pygrank_b200
    import scipy.sparse as sp
    with pytest.raises(Exception, match="no CPU fallback"):
        pygrank_b200.DeviceGraph.from_scipy(sp.eye(4, format="csr"))
    from pygrank_b200 import backend
    with pytest.raises(Exception, match="no CPU fallback"):
        backend.backend_init()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pygrank_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "reference_port" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_host_generators_are_deterministic():
    from pygrank_b200 import synthetic
    a = synthetic.rmat_edges_host(10, 4, seed=3)
    b = synthetic.rmat_edges_host(10, 4, seed=3)
    c = synthetic.rmat_edges_host(10, 4, seed=3, first_edge=100, num_edges=50)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(a[0][100:150], c[0]) and np.array_equal(a[1][100:150], c[1])
    A = synthetic.rmat_graph_host(10, 4, seed=3)
    assert (A != A.T).nnz == 0 and A.diagonal().sum() == 0 and set(np.unique(A.data)) == {1.0}
    B = synthetic.ba_graph_host(3000, 4, seed=2)
    assert (B != B.T).nnz == 0 and np.diff(B.indptr).min() >= 1


def test_backend_registration_without_gpu():
    """install() wires "b200" into an unmodified pygrank; unknown names still raise
    (reference tests/test_core.py:89-97); without a GPU loading it fails loudly and leaves numpy active."""
    from refutil import import_pygrank
    pg = import_pygrank()
    if pg is None:
        pytest.skip("baseline/_ref not installed")
    import torch
    import pygrank_b200
    pygrank_b200.install(pg)
    with pytest.raises(Exception):
        pg.load_backend("unknown")
    assert pg.backend_name() == "numpy"
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="no CPU fallback"):
            pg.load_backend("b200")
        pg.load_backend("numpy")
        assert pg.backend_name() == "numpy"
        assert "matvec" not in pg.core.backend._imported_mods
