"""The oracle port must reproduce the real reference's outputs (tests/golden) bit for bit."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import reference_port as orc

NORMS = [("auto", 0), ("auto", 1), ("symmetric", 0), ("symmetric", 1), ("col", 0), ("laplacian", 0), ("both", 0),
         ("none", 0)]


def test_c_scatter_matches_python_loop_and_scipy():
    rng = np.random.default_rng(0)
    M = sp.random(40, 40, density=0.2, random_state=3, format="csr")
    x = rng.uniform(-1, 1, 40)
    y = orc.conv(x, M)
    ref = orc.conv_python(list(x), M.indptr.tolist(), M.indices.tolist(), M.data.tolist(), 40)
    assert np.array_equal(y, np.array(ref))
    assert np.array_equal(y, x @ M)


def test_conv_matches_reference(golden):
    name, z, A, directed = golden
    M = orc.to_sparse_matrix(A, "auto", directed)
    assert np.array_equal(orc.conv(z["conv_x"], M), z["conv_y"])
    assert np.array_equal(z["conv_x"] @ M, z["conv_y"])


@pytest.mark.parametrize("norm,renorm", NORMS)
def test_normalisation_bit_exact(golden, norm, renorm):
    name, z, A, directed = golden
    M = orc.to_sparse_matrix(A, norm, directed, renormalize=bool(renorm))
    key = f"norm_{norm}_{renorm}"
    assert np.array_equal(M.indptr, z[key + "_indptr"])
    assert np.array_equal(M.indices, z[key + "_indices"])
    assert np.array_equal(M.data, z[key + "_data"])
    R = orc.to_sparse_matrix(A, norm, directed, renormalize=bool(renorm), reference_storage_order=True)
    assert np.array_equal(orc.degrees(R), z[key + "_degrees"])


RUNS = {
    "ppr85": ("auto", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)),
    "ppr90_noq": ("auto", lambda M, p: orc.pagerank(M, p, 0.9, tol=1e-9, use_quotient=False, max_iters=1000)),
    "ppr85_sym": ("symmetric", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)),
    "ppr85_col": ("col", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)),
    "ppr85_tol6_mod3": ("auto", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-6, end_modulo=3)),
    "ppr85_iters20": ("auto", lambda M, p: orc.pagerank(M, p, 0.85, error_type="iters", max_iters=20)),
    "ppr85_l1": ("auto", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-7, error_type="l1", max_iters=1000)),
    "ppr85_msq": ("auto", lambda M, p: orc.pagerank(M, p, 0.85, tol=1e-16, error_type="msq", max_iters=1000)),
    "heat3": ("auto", lambda M, p: orc.heat_kernel(M, p, 3)),
    "heat3_tol9": ("auto", lambda M, p: orc.heat_kernel(M, p, 3, tol=1e-9)),
    "heat5_sym": ("symmetric", lambda M, p: orc.heat_kernel(M, p, 5, tol=1e-9)),
    "gen40": ("auto", lambda M, p: orc.generic_filter(M, p, [0.9 ** k for k in range(40)], error_type="iters",
                                                      max_iters=41)),
    "gen3_tol": ("auto", lambda M, p: orc.generic_filter(M, p, [0.5, 0.25, 0.125], tol=1e-9)),
    "pprclosed": ("auto", lambda M, p: orc.pagerank_closed(M, p, 0.85, tol=1e-9, max_iters=1000)),
    "absorb": ("auto", lambda M, p: orc.absorbing_walks(M, p, tol=1e-9, max_iters=1000)),
    "absorb85": ("auto", lambda M, p: orc.absorbing_walks(M, p, 0.85, tol=1e-9, max_iters=1000)),
    "absorb85_col": ("col", lambda M, p: orc.absorbing_walks(M, p, 0.85, tol=1e-9, max_iters=1000)),
}


@pytest.mark.parametrize("run", sorted(RUNS))
def test_filters_bit_exact(golden, run):
    name, z, A, directed = golden
    norm, fn = RUNS[run]
    M = orc.to_sparse_matrix(A, norm, directed, reference_storage_order=True)
    P = z["P"]
    for c in range(P.shape[1]):
        scores, iters, _ = fn(M, P[:, c])
        assert iters == int(z[f"run_{run}_iters"][c]), (name, run, c)
        assert np.array_equal(scores, z[f"run_{run}_scores"][:, c]), (name, run, c)


def test_custom_absorption_and_propagate(golden):
    name, z, A, directed = golden
    M = orc.to_sparse_matrix(A, "auto", directed)
    scores, iters, _ = orc.absorbing_walks(M, z["P"][:, 0], 0.9, absorption=z["absorption"], tol=1e-9, max_iters=1000)
    assert iters == int(z["run_absorb90_custom_iters"][0])
    assert np.array_equal(scores, z["run_absorb90_custom_scores"])
    out, _ = orc.propagate(orc.pagerank, M, z["P"], alpha=0.85, tol=1e-9, max_iters=1000)
    assert np.array_equal(out, z["run_propagate_ppr85"])


def test_max_iters_raises_and_zero_personalization():
    _, A, directed = __import__("conftest").load_golden("ba2000")
    M = orc.to_sparse_matrix(A, "auto", directed)
    p = np.zeros(A.shape[0]); p[3] = 1
    with pytest.raises(Exception):
        orc.pagerank(M, p, 0.99, tol=1e-14, max_iters=5)
    r, iters, _ = orc.pagerank(M, np.zeros(A.shape[0]))
    assert iters == 0 and not r.any()


def test_fast_c_generator_equals_numpy_recipe():
    """oracle/rmat_gen.c (bench CPU legs) == pygrank_b200.synthetic (the documented recipe), bit for bit."""
    from oracle import fastgen
    from pygrank_b200 import synthetic
    for scale, seed in ((9, 1), (12, 7), (13, 1)):
        s0, d0 = synthetic.rmat_edges_host(scale, 16, seed)
        s1, d1 = fastgen.rmat_edges(scale, 16, seed)
        assert np.array_equal(s0, s1) and np.array_equal(d0, d1)
        A0 = synthetic.rmat_graph_host(scale, 16, seed)
        A1 = fastgen.rmat_graph(scale, 16, seed)
        assert np.array_equal(A0.indptr, A1.indptr) and np.array_equal(A0.indices, A1.indices)
        assert np.array_equal(A0.data, A1.data)
