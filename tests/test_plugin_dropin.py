"""Drop-in: the UNMODIFIED reference (baseline/_ref) runs its own filters on the b200 backend.

These read like the reference's tests (tests/test_core.py, tests/test_filters.py): a
``supported_backends()`` loop over ["numpy", "b200"], relations between algorithms, and direct
parity of pg.PageRank(...) outputs between the two backends on the same synthetic graph.
"""
import numpy as np
import pytest

from conftest import load_golden, rel_l1
from refutil import import_pygrank

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pg():
    mod = import_pygrank()
    if mod is None:
        pytest.skip("baseline/_ref (the unmodified reference) is not installed on this box")
    import pygrank_b200
    pygrank_b200.install(mod)
    yield mod
    mod.load_backend("numpy")


def supported_backends(pg):
    for backend in ["numpy", "b200"]:
        pg.load_backend(backend)
        yield backend
    pg.load_backend("numpy")


def test_backend_load_and_with(pg):
    pg.load_backend("b200")
    assert pg.backend_name() == "b200"
    pg.load_backend("numpy")
    assert pg.backend_name() == "numpy"
    with pytest.raises(Exception):
        pg.load_backend("unknown")
    with pg.Backend("b200") as backend:
        assert pg.backend_name() == "b200" and backend.backend_name() == "b200"
    assert pg.backend_name() == "numpy"


def test_primitive_conversion(pg):
    for backend in supported_backends(pg):
        assert pg.sum(pg.to_array([1, 2, 3])) == 6
        assert float(pg.sum(pg.dot(pg.exp(pg.log(pg.to_array([4, 5]))), pg.to_array([2, 2])))) == pytest.approx(18)
        if backend == "numpy":
            continue   # the reference's numpy to_primitive uses np.array(copy=False), which numpy >= 2 rejects
        primitive = pg.to_array([1, 2, 3])
        assert id(primitive) == id(pg.to_array(primitive, copy_array=False))
        assert id(primitive) != id(pg.to_array(primitive, copy_array=True))
        table = pg.to_primitive([[1, 2, 3], [4, 5, 6]])
        cols = pg.separate_cols(table)
        assert len(cols) == 3 and all(pg.length(c) == 2 for c in cols)
        assert pg.sum(pg.abs(table - pg.combine_cols(cols))) == 0


def test_signal_direct_operations(pg):
    import networkx as nx
    for _ in supported_backends(pg):
        graph = nx.DiGraph([(1, 2), (2, 3)])
        signal = pg.to_signal(graph, [1., 2., 3.])
        assert pg.sum(signal) == 6
        assert pg.sum(signal + 1) == 9
        assert pg.sum(signal ** 2) == 14
        assert pg.sum(signal / pg.to_signal(graph, [1., 2., 3.])) == 3
        signal *= 4
        assert pg.sum(signal) == 24
        assert signal[2] == 8.0
        del signal[2]
        assert signal[2] == 0


@pytest.mark.parametrize("name", ["ba2000", "rmat10", "gnp600d", "weighted300"])
def test_reference_filters_unchanged_on_b200(pg, name):
    z, A, directed = load_golden(name)
    P = z["P"]
    runs = {
        "ppr85": lambda pre: pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000),
        "ppr90_noq": lambda pre: pg.PageRank(0.9, tol=1e-9, use_quotient=False, preprocessor=pre, max_iters=1000),
        "heat3_tol9": lambda pre: pg.HeatKernel(3, tol=1e-9, preprocessor=pre),
        "gen40": lambda pre: pg.GenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters", max_iters=41,
                                                   preprocessor=pre),
        "absorb85": lambda pre: pg.AbsorbingWalks(0.85, tol=1e-9, preprocessor=pre, max_iters=1000),
    }
    with pg.Backend("b200"):
        graph = pg.AdjacencyWrapper(A, directed=directed)
        pre = pg.preprocessor(normalization="auto", assume_immutability=True)   # the reference's own preprocessor
        M = pre(graph)
        assert M.array.__class__.__name__ == "DeviceGraph"
        from pygrank_b200 import lazy
        for rname, make in runs.items():
            for c in (0, 3):
                alg = make(pre)
                lazy.reset_stats()
                r = alg(pg.to_signal(M, P[:, c].copy()))
                iters = int(z[f"run_{rname}_iters"][c])
                assert alg.convergence.iteration == iters, (name, rname, c)
                got = r.np.cpu().numpy()
                assert rel_l1(got, z[f"run_{rname}_scores"][:, c]) <= 1e-10, (name, rname, c)
                # the unmodified driver reached the FUSED kernels: every conv ran inside a fused step, and the number
                # of eager elementwise kernels / host synchronisations does not grow with the iteration count
                st = dict(lazy.STATS)
                assert st["eager_convs"] == 0 and st["runs"] == 1 and st["recomputed_runs"] == 0, (name, rname, c, st)
                assert st["fused_steps"] >= iters - 2, (name, rname, c, st)
                assert st["eager_ops"] <= 16 and st["syncs"] <= 8 + (iters if rname == "heat3_tol9" else 0), (name, rname, c, st)


def test_device_preprocessor_injected_into_reference_filter(pg):
    """`pg.PageRank(preprocessor=pygrank_b200.preprocessor(...))`: CSR built and normalised on the device."""
    import pygrank_b200
    z, A, directed = load_golden("ba2000")
    P = z["P"]
    with pg.Backend("b200"):
        pre = pygrank_b200.preprocessor(normalization="symmetric", assume_immutability=True)
        graph = pg.AdjacencyWrapper(A, directed=False)
        alg = pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000)
        r = alg(pg.to_signal(pre(graph), P[:, 0].copy()))
        assert alg.convergence.iteration == int(z["run_ppr85_sym_iters"][0])
        assert rel_l1(r.np.cpu().numpy(), z["run_ppr85_sym_scores"][:, 0]) <= 1e-10
        # AbsorbingWalks == PageRank under col normalisation (reference tests/test_filters.py:151-157)
        pre_col = pygrank_b200.preprocessor(normalization="col", assume_immutability=True)
        a = pg.AbsorbingWalks(0.85, tol=1e-12, preprocessor=pre_col, max_iters=1000)(pg.to_signal(pre_col(graph), P[:, 0].copy()))
        b = pg.PageRank(0.85, tol=1e-12, preprocessor=pre_col, max_iters=1000)(pg.to_signal(pre_col(graph), P[:, 0].copy()))
        assert float(pg.sum(pg.abs(a.np - b.np))) < 1e-9


def test_graph_dropout_semantics(pg):
    import torch
    z, A, directed = load_golden("ba2000")
    with pg.Backend("b200") as backend:
        M = backend.scipy_sparse_to_backend(A)
        assert backend.graph_dropout(M, 0) is M
        x = backend.to_array(np.ones(A.shape[0]))
        full = backend.conv(x, M)
        torch.manual_seed(0)
        dropped = backend.conv(x, backend.graph_dropout(M, 0.5))
        # survivors rescaled by 1/(1-p): the expectation is preserved
        assert abs(float(dropped.sum()) / float(full.sum()) - 1.0) < 0.05
        assert float((dropped - full).abs().sum()) > 0


def test_graph_dropout_on_a_directed_graph(pg):
    """graph_dropout (torch backends' semantics, core/backend/pytorch.py:34-38) also on directed graphs: the
    pull view conv reads is masked and rescaled; p = 0 is the identity; the expectation is preserved."""
    import torch
    import pygrank_b200
    z, A, directed = load_golden("gnp600d")
    assert directed
    g = pygrank_b200.DeviceGraph.from_scipy(A, directed=True, normalization="col")
    x = torch.ones(A.shape[0], dtype=torch.float64, device="cuda")
    full = g.conv(x)
    assert g.dropout(0) is g
    torch.manual_seed(1)
    dropped = g.dropout(0.3).conv(x)
    assert abs(float(dropped.sum()) / float(full.sum()) - 1.0) < 0.05
    assert float((dropped - full).abs().sum()) > 0
    assert float((g.conv(x) - full).abs().sum()) == 0          # the original graph is untouched


def test_plugin_route_fp32_dict_personalization_and_device_preprocessor(pg):
    """The route bench.py's e2e_plugin times: device preprocessor injected, dict personalization, fp32 vectors."""
    import torch
    import pygrank_b200
    from pygrank_b200 import backend as b200, lazy
    z, A, directed = load_golden("rmat10")
    P = z["P"]
    b200.configure(dtype=torch.float32)
    try:
        with pg.Backend("b200"):
            pre = pygrank_b200.preprocessor(normalization="symmetric", assume_immutability=True)
            graph = pre(pg.AdjacencyWrapper(A, directed=False))
            alg = pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000)
            seeds = {int(i): float(P[i, 0]) for i in np.nonzero(P[:, 0])[0]}
            lazy.reset_stats()
            r = alg(pg.to_signal(graph, seeds))
            assert abs(alg.convergence.iteration - int(z["run_ppr85_sym_iters"][0])) <= 1
            assert rel_l1(r.np.cpu().numpy(), z["run_ppr85_sym_scores"][:, 0]) <= 1e-5
            assert lazy.STATS["eager_convs"] == 0 and lazy.STATS["runs"] == 1
            assert isinstance(r[int(np.nonzero(P[:, 0])[0][0])], float)
    finally:
        b200.configure(dtype=torch.float64)


def test_plugin_route_eager_fallbacks_stay_correct(pg):
    """Shapes the engine does not fuse (measure classes it does not know, postprocessor quotients, eigenvector
    convergence, end_modulo > 1, max_iters cut-offs) still give the numpy backend's answer."""
    z, A, directed = load_golden("ba2000")
    P = z["P"]
    cases = {
        "end_modulo": lambda pre: pg.PageRank(0.85, tol=1e-6, end_modulo=3, preprocessor=pre),
        "iters20": lambda pre: pg.PageRank(0.85, error_type="iters", max_iters=20, preprocessor=pre),
        "msq": lambda pre: pg.PageRank(0.85, tol=1e-16, error_type=pg.MSQ, max_iters=1000, preprocessor=pre),
        "maxdiff": lambda pre: pg.PageRank(0.85, tol=1e-9, error_type=pg.MaxDifference, max_iters=1000, preprocessor=pre),
        "rmabs": lambda pre: pg.PageRank(0.85, tol=1e-7, error_type=pg.RMabs, max_iters=1000, preprocessor=pre),
        "eig": lambda pre: pg.PageRank(0.85, tol=1e-9, converge_to_eigenvectors=True, max_iters=1000, preprocessor=pre),
        "cheby": lambda pre: pg.HeatKernel(3, tol=1e-9, coefficient_type="chebyshev", preprocessor=pre),
        "gen3": lambda pre: pg.GenericGraphFilter([0.5, 0.25, 0.125], tol=1e-9, preprocessor=pre),
        "closed": lambda pre: pg.PageRankClosed(0.85, tol=1e-9, max_iters=1000, preprocessor=pre),
        "absorb_custom": lambda pre: pg.AbsorbingWalks(0.85, tol=1e-9, max_iters=1000, preprocessor=pre),
        # measures a polynomial run does not keep on the device: consecutive results are read back and compared eagerly
        "heat_rmabs": lambda pre: pg.HeatKernel(3, tol=1e-7, error_type=pg.RMabs, preprocessor=pre),
        "heat_maxdiff": lambda pre: pg.HeatKernel(3, tol=1e-9, error_type=pg.MaxDifference, preprocessor=pre),
        "heat_msq": lambda pre: pg.HeatKernel(3, tol=1e-16, error_type=pg.MSQ, preprocessor=pre),
        "heat_modulo": lambda pre: pg.HeatKernel(3, tol=1e-9, end_modulo=2, preprocessor=pre),
    }
    results = {}
    for backend in ["numpy", "b200"]:
        pg.load_backend(backend)
        graph = pg.AdjacencyWrapper(A, directed=directed)
        pre = pg.preprocessor(normalization="auto", assume_immutability=True)
        for cname, make in cases.items():
            alg = make(pre)
            kwargs = {"absorption": {i: 1.0 + (i % 3) for i in range(A.shape[0])}} if cname == "absorb_custom" else {}
            r = alg(pg.to_signal(graph, P[:, 1].copy()), **kwargs)
            out = r.np if backend == "numpy" else r.np.cpu().numpy()
            results.setdefault(cname, []).append((np.asarray(out, dtype=np.float64), alg.convergence.iteration))
    pg.load_backend("numpy")
    for cname, ((ref, it_ref), (got, it_got)) in results.items():
        assert it_ref == it_got, (cname, it_ref, it_got)
        assert rel_l1(got, ref) <= 1e-10, cname


def test_parameter_tuner_runs_unchanged_on_b200(pg):
    """pg.ParameterTuner (algorithms/autotune/parameterized.py:117-167 with optimization.py's line search) evaluates
    its candidates through the fused runs on the b200 backend and lands on the parameters the numpy backend finds."""
    z, A, directed = load_golden("ba2000")
    rng = np.random.default_rng(0)
    n = A.shape[0]
    members = rng.choice(n, 200, replace=False)
    found = {}
    for backend in ["numpy", "b200"]:
        pg.load_backend(backend)
        graph = pg.AdjacencyWrapper(A, directed=directed)
        signal = pg.to_signal(graph, {int(v): 1.0 for v in members})
        pre = pg.preprocessor(normalization="symmetric", assume_immutability=True)
        tuner = pg.ParameterTuner(lambda params: pg.PageRank(params[0], preprocessor=pre, tol=1e-9, max_iters=1000),
                                  max_vals=[0.99], min_vals=[0.5], measure=pg.AUC, deviation_tol=0.01,
                                  tuning_backend=backend)
        ranks = tuner(graph, signal)
        found[backend] = (list(tuner.last_params), np.asarray(ranks.np if backend == "numpy" else ranks.np.cpu().numpy()))
    pg.load_backend("numpy")
    assert found["numpy"][0] == pytest.approx(found["b200"][0], abs=1e-9)
    assert rel_l1(found["b200"][1], found["numpy"][1]) <= 1e-9


def test_parameter_tuner_candidates_run_as_panels(pg):
    """pg.ParameterTuner, unmodified, with pygrank_b200.AlphaSweep as its ranker generator and optimizer: every round's
    candidates are columns of one panel sweep (pgb_affine_steps_panel, per-column alpha), and the tuner lands on the
    parameters and scores the stock numpy run finds."""
    import pygrank_b200
    z, A, directed = load_golden("ba2000")
    rng = np.random.default_rng(0)
    n = A.shape[0]
    members = rng.choice(n, 200, replace=False)
    args = dict(max_vals=[0.99], min_vals=[0.5], measure=pg.AUC, deviation_tol=0.01)
    pg.load_backend("numpy")
    graph = pg.AdjacencyWrapper(A, directed=directed)
    pre = pg.preprocessor(normalization="symmetric", assume_immutability=True)
    stock = pg.ParameterTuner(lambda params: pg.PageRank(params[0], preprocessor=pre, tol=1e-9, max_iters=1000),
                              tuning_backend="numpy", **args)
    ref = stock(graph, pg.to_signal(graph, {int(v): 1.0 for v in members}))
    ref_scores = np.asarray(ref.np, dtype=np.float64)
    pg.load_backend("b200")
    try:
        sweep = pygrank_b200.AlphaSweep(tol=1e-9, max_iters=1000, normalization="symmetric")
        tuner = pg.ParameterTuner(sweep.ranker, optimizer=sweep.optimizer, tuning_backend="b200", **args)
        graph = pg.AdjacencyWrapper(A, directed=directed)
        got = tuner(graph, pg.to_signal(graph, {int(v): 1.0 for v in members}))
        got_scores = got.np.cpu().numpy()
    finally:
        pg.load_backend("numpy")
    assert list(tuner.last_params) == pytest.approx(list(stock.last_params), abs=1e-9)
    assert rel_l1(got_scores, ref_scores) <= 1e-9
    # 5 candidates per round (the stock default): one sweep per round carries the round's new candidates (those of an
    # earlier round are served from the cache), plus one single-column solve for the final ranking
    assert sweep.stats["sweeps"] >= 2 and sweep.stats["columns"] >= 5 + (sweep.stats["sweeps"] - 1)
    assert sweep.stats["served"] >= 2 * 5 + 1 > sweep.stats["sweeps"]
