"""ClosedFormGraphFilter options beyond the fused Taylor/node-space case — Chebyshev coefficients, the Krylov space,
the optimization_dict power cache (/root/reference/pygrank/algorithms/filters/abstract_filters.py:196-246,
filters/krylov_space.py:16-83) — in the mirror classes, against the unmodified reference on its numpy backend."""
import warnings

import numpy as np
import pytest

from conftest import load_golden, rel_l1
from refutil import import_pygrank

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    mod = import_pygrank()
    if mod is None:
        pytest.skip("baseline/_ref is not installed on this box")
    mod.load_backend("numpy")
    return mod


def _ref(pg, A, directed, make, p, **kw):
    graph = pg.AdjacencyWrapper(A, directed=directed)
    pre = pg.preprocessor(normalization="auto", assume_immutability=True)
    alg = make(pg, preprocessor=pre)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = alg(pg.to_signal(graph, p.copy()), **kw)
    return np.asarray(r.np, dtype=np.float64), alg.convergence.iteration


@pytest.mark.parametrize("name", ["ba2000", "rmat10"])
def test_chebyshev_coefficients(pg, name):
    import pygrank_b200 as pgb
    z, A, directed = load_golden(name)
    g = pgb.DeviceGraph.from_scipy(A, directed=directed)
    for c in (0, 2):
        p = z["P"][:, c]
        for make in (lambda m, **k: m.HeatKernel(3, tol=1e-9, coefficient_type="chebyshev", **k),
                     lambda m, **k: m.GenericGraphFilter([0.5, 0.25, 0.125, 0.06], tol=1e-9, coefficient_type="chebyshev", **k)):
            ref, it = _ref(pg, A, directed, make, p)
            alg = make(pgb)
            got = alg(g, p.copy())
            assert alg.convergence.iteration == it
            assert rel_l1(got.numpy(), ref) <= 1e-10


def test_optimization_dict_caches_powers(pg):
    import pygrank_b200 as pgb
    from pygrank_b200 import _capi as C
    z, A, directed = load_golden("ba2000")
    g = pgb.DeviceGraph.from_scipy(A, directed=directed)
    p = z["P"][:, 1].copy()
    cache, ref_cache = dict(), dict()
    graph = pg.AdjacencyWrapper(A, directed=directed)
    pre = pg.preprocessor(normalization="auto", assume_immutability=True)
    sig = pg.to_signal(graph, p.copy())
    import torch
    pt = torch.from_numpy(p).cuda()
    convs = []
    orig = g.conv
    g.conv = lambda x: (convs.append(1), orig(x))[1]
    for t in (3, 5, 2):
        ref_alg = pg.HeatKernel(t, tol=1e-9, optimization_dict=ref_cache, preprocessor=pre)
        ref = ref_alg(sig)
        alg = pgb.HeatKernel(t, tol=1e-9, optimization_dict=cache)
        before = len(convs)
        got = alg(g, pt)
        assert alg.convergence.iteration == ref_alg.convergence.iteration
        assert rel_l1(got.numpy(), np.asarray(ref.np)) <= 1e-10
        if t == 2:   # fewer hops than an earlier run on the same personalization: every power comes from the cache
            assert len(convs) == before
    assert len(cache) == 1 and len(next(iter(cache.values()))) >= 10


def test_krylov_space(pg):
    """The reference's numpy backend cannot run its own Krylov code under numpy >= 2 (``np.array(obj, copy=False)`` in
    core/backend/numpy.py:50), so the reference algorithm is executed by the UNMODIFIED reference driver on the b200
    backend (every op eager) and the mirror class must agree with it."""
    import pygrank_b200 as pgb
    z, A, directed = load_golden("ba2000")
    g = pgb.DeviceGraph.from_scipy(A, directed=directed)
    p = z["P"][:, 0]
    make = lambda m, **k: m.HeatKernel(3, tol=1e-9, krylov_dims=8, **k)
    pgb.install(pg)
    try:
        with pg.Backend("b200"):
            graph = pg.AdjacencyWrapper(A, directed=directed)
            pre = pg.preprocessor(normalization="auto", assume_immutability=True)
            ref_alg = make(pg, preprocessor=pre)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                ref = ref_alg(pg.to_signal(graph, p.copy())).np.cpu().numpy()
    finally:
        pg.load_backend("numpy")
    alg = make(pgb)
    got = alg(g, p.copy())
    assert alg.convergence.iteration == ref_alg.convergence.iteration
    assert rel_l1(got.numpy(), ref) <= 1e-8          # Lanczos amplifies summation-order differences
    # and, up to the scale of its L2-normalised basis, the approximation points where the exact filter points
    exact = pgb.HeatKernel(3, tol=1e-9)(g, p.copy()).numpy()
    a = got.numpy()
    assert float(a @ exact) / (np.linalg.norm(a) * np.linalg.norm(exact)) >= 0.99
