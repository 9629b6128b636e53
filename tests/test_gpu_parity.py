"""Parity of the CUDA path (through the C-ABI) against the oracle and the golden vectors.

Bars (BASELINE.json north_star): CSR in canonical form, graph degrees and iteration counts
bit-exact; fp64 scores <= 1e-10 relative L1; fp32 mode <= 1e-5 relative L1.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import GOLDEN_GRAPHS, load_golden, rel_l1

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-10
FP32_TOL = 1e-5


@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch


@pytest.fixture(scope="module")
def pgb(torch_cuda):
    import pygrank_b200
    pygrank_b200.lib()     # raises if the extension is missing: no fallback
    return pygrank_b200


@pytest.fixture(scope="module")
def orc():
    from oracle import reference_port
    return reference_port


def _graph(pgb, A, directed, norm="auto", relabel="degree", renormalize=False):
    return pgb.DeviceGraph.from_scipy(A, directed=directed, normalization=norm, relabel=relabel,
                                      renormalize=renormalize)


# ---------------------------------------------------------------------------- generators / build
def test_generators_bit_identical_to_host(pgb, torch_cuda):
    from pygrank_b200 import synthetic, device_synthetic
    src_h, dst_h = synthetic.rmat_edges_host(12, 8, seed=5)
    src_d, dst_d = device_synthetic.rmat_edges_device(12, 8, seed=5)
    assert np.array_equal(src_h, src_d.cpu().numpy()) and np.array_equal(dst_h, dst_d.cpu().numpy())
    src_h, dst_h = synthetic.ba_edges_host(5000, 4, seed=9)
    src_d, dst_d = device_synthetic.ba_edges_device(5000, 4, seed=9)
    assert np.array_equal(src_h, src_d.cpu().numpy()) and np.array_equal(dst_h, dst_d.cpu().numpy())


@pytest.mark.parametrize("relabel", ["none", "degree"])
def test_csr_build_matches_scipy(pgb, torch_cuda, relabel):
    from pygrank_b200 import synthetic, device_synthetic
    scale = 13
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=2)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="none", relabel=relabel)
    A = synthetic.rmat_graph_host(scale, 16, seed=2)
    M = g.to_scipy_normalized()
    assert M.nnz == A.nnz
    assert np.array_equal(M.indptr, A.indptr) and np.array_equal(M.indices, A.indices)
    assert np.array_equal(M.data, A.data)
    rowsum, colsum = g.graph_degrees()
    assert np.array_equal(rowsum.cpu().numpy(), np.diff(A.indptr).astype(np.float64))
    assert np.array_equal(colsum.cpu().numpy(), np.diff(A.indptr).astype(np.float64))


def test_directed_build_and_transpose(pgb, torch_cuda):
    torch = torch_cuda
    rng = np.random.default_rng(3)
    n, m = 700, 9000
    src = rng.integers(0, n, m).astype(np.int32)
    dst = rng.integers(0, n, m).astype(np.int32)
    w = rng.uniform(0.5, 2.0, m)
    g = pgb.DeviceGraph.from_edges(n, torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda(),
                                   torch.from_numpy(w).cuda(), directed=True, normalization="none", relabel="none")
    A = sp.coo_matrix((w, (src, dst)), shape=(n, n)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    M = g.to_scipy_normalized()
    assert np.array_equal(M.indptr, A.indptr) and np.array_equal(M.indices, A.indices)
    assert np.allclose(M.data, A.data, rtol=1e-15, atol=0)
    x = rng.uniform(-1, 1, n)
    y = g.conv(torch.from_numpy(x).cuda()).cpu().numpy()
    assert rel_l1(y, x @ A) < 1e-13


NORMS = [("auto", 0), ("auto", 1), ("symmetric", 0), ("symmetric", 1), ("col", 0), ("both", 0), ("none", 0)]


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("norm,renorm", NORMS)
@pytest.mark.parametrize("relabel", ["none", "degree"])
def test_normalised_csr_bit_exact(pgb, name, norm, renorm, relabel):
    z, A, directed = load_golden(name)
    g = _graph(pgb, A, directed, norm, relabel, renormalize=bool(renorm))
    M = g.to_scipy_normalized()
    key = f"norm_{norm}_{renorm}"
    assert np.array_equal(M.indptr, z[key + "_indptr"])
    assert np.array_equal(M.indices, z[key + "_indices"])
    if name == "weighted300":
        # float-weighted degrees depend on numpy's pairwise order -> scales may differ in the last bit
        assert np.allclose(M.data, z[key + "_data"], rtol=2e-15, atol=0)
    else:
        assert np.array_equal(M.data, z[key + "_data"])
    deg = g.degrees().cpu().numpy()
    if relabel == "none" and name != "weighted300":
        assert np.array_equal(deg, z[key + "_degrees"])
    else:
        assert np.allclose(deg, z[key + "_degrees"], rtol=1e-14, atol=1e-300)


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
def test_laplacian_identity(pgb, torch_cuda, name):
    """tests/test_preprocessor.py:6-15 of the reference: conv(s, laplacian) + conv(s, symmetric) - s == 0."""
    z, A, directed = load_golden(name)
    x = torch_cuda.from_numpy(z["conv_x"]).cuda()
    lap = _graph(pgb, A, directed, "laplacian").conv(x)
    sym = _graph(pgb, A, directed, "symmetric").conv(x)
    assert float((lap + sym - x).abs().sum()) <= 1e-13 * float(x.abs().sum())


# ---------------------------------------------------------------------------- conv
@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("relabel", ["none", "degree", "hub"])
def test_conv_matches_reference(pgb, torch_cuda, name, relabel):
    torch = torch_cuda
    z, A, directed = load_golden(name)
    g = _graph(pgb, A, directed, "auto", relabel)
    x = torch.from_numpy(z["conv_x"]).cuda()
    assert rel_l1(g.conv(x).cpu().numpy(), z["conv_y"]) < 1e-14
    y32 = g.conv(x.float()).cpu().numpy()
    assert rel_l1(y32, z["conv_y"]) < FP32_TOL


def test_conv_linearity_and_mass_rmat18(pgb, torch_cuda, orc):
    """Size-independent properties on a graph with hubs, empty rows and rows cut by tile boundaries."""
    torch = torch_cuda
    from pygrank_b200 import device_synthetic
    scale = 18
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=1)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="symmetric")
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    y = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    lhs = g.conv(2.5 * x - 0.75 * y)
    rhs = 2.5 * g.conv(x) - 0.75 * g.conv(y)
    assert float((lhs - rhs).abs().sum()) <= 1e-13 * float(rhs.abs().sum())
    # sum(x @ M) == x . rowsum(M)
    assert abs(float(g.conv(x).sum()) - float((x * g.degrees()).sum())) <= 1e-11 * float(x.sum())
    # against the oracle's scatter on the same matrix
    M = g.to_scipy_normalized()
    assert rel_l1(g.conv(x).cpu().numpy(), orc.conv(x.cpu().numpy(), M)) < 1e-13


# ---------------------------------------------------------------------------- fused filters
def _runs(pgb):
    P = pgb
    return {
        "ppr85": ("auto", lambda kw: P.PageRank(0.85, tol=1e-9, max_iters=1000, **kw), {}),
        "ppr90_noq": ("auto", lambda kw: P.PageRank(0.9, tol=1e-9, use_quotient=False, max_iters=1000, **kw), {}),
        "ppr85_sym": ("symmetric", lambda kw: P.PageRank(0.85, tol=1e-9, max_iters=1000, **kw), {}),
        "ppr85_col": ("col", lambda kw: P.PageRank(0.85, tol=1e-9, max_iters=1000, **kw), {}),
        "ppr85_tol6_mod3": ("auto", lambda kw: P.PageRank(0.85, tol=1e-6, end_modulo=3, **kw), {}),
        "ppr85_iters20": ("auto", lambda kw: P.PageRank(0.85, error_type="iters", max_iters=20, **kw), {}),
        "ppr85_l1": ("auto", lambda kw: P.PageRank(0.85, tol=1e-7, error_type="L1", max_iters=1000, **kw), {}),
        "ppr85_msq": ("auto", lambda kw: P.PageRank(0.85, tol=1e-16, error_type="MSQ", max_iters=1000, **kw), {}),
        "heat3": ("auto", lambda kw: P.HeatKernel(3, **kw), {}),
        "heat3_tol9": ("auto", lambda kw: P.HeatKernel(3, tol=1e-9, **kw), {}),
        "heat5_sym": ("symmetric", lambda kw: P.HeatKernel(5, tol=1e-9, **kw), {}),
        "gen40": ("auto", lambda kw: P.GenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters",
                                                          max_iters=41, **kw), {}),
        "gen3_tol": ("auto", lambda kw: P.GenericGraphFilter([0.5, 0.25, 0.125], tol=1e-9, **kw), {}),
        "pprclosed": ("auto", lambda kw: P.PageRankClosed(0.85, tol=1e-9, max_iters=1000, **kw), {}),
        "absorb": ("auto", lambda kw: P.AbsorbingWalks(tol=1e-9, max_iters=1000, **kw), {}),
        "absorb85": ("auto", lambda kw: P.AbsorbingWalks(0.85, tol=1e-9, max_iters=1000, **kw), {}),
        "absorb85_col": ("col", lambda kw: P.AbsorbingWalks(0.85, tol=1e-9, max_iters=1000, **kw), {}),
    }


RUN_NAMES = ["ppr85", "ppr90_noq", "ppr85_sym", "ppr85_col", "ppr85_tol6_mod3", "ppr85_iters20", "ppr85_l1",
             "ppr85_msq", "heat3", "heat3_tol9", "heat5_sym", "gen40", "gen3_tol", "pprclosed", "absorb", "absorb85",
             "absorb85_col"]


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("run", RUN_NAMES)
@pytest.mark.parametrize("relabel", ["hub", "degree", "none"])
def test_filters_fp64_match_golden(pgb, torch_cuda, name, run, relabel):
    torch = torch_cuda
    z, A, directed = load_golden(name)
    norm, make, _ = _runs(pgb)[run]
    g = _graph(pgb, A, directed, norm, relabel)
    P = z["P"]
    for c in range(P.shape[1]):
        alg = make({"dtype": torch.float64})
        r = alg(g, P[:, c])
        assert alg.convergence.iteration == int(z[f"run_{run}_iters"][c]), (name, run, c)
        assert rel_l1(r.numpy(), z[f"run_{run}_scores"][:, c]) <= FP64_TOL, (name, run, c)


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("run", ["ppr85", "ppr85_col", "heat3_tol9", "gen40", "absorb85"])
def test_filters_fp32_mode(pgb, torch_cuda, name, run):
    torch = torch_cuda
    z, A, directed = load_golden(name)
    norm, make, _ = _runs(pgb)[run]
    g = _graph(pgb, A, directed, norm)
    P = z["P"]
    for c in range(P.shape[1]):
        alg = make({"dtype": torch.float32})
        r = alg(g, P[:, c])
        assert abs(alg.convergence.iteration - int(z[f"run_{run}_iters"][c])) <= 1, (name, run, c)
        assert rel_l1(r.numpy(), z[f"run_{run}_scores"][:, c]) <= FP32_TOL, (name, run, c)


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
def test_error_sequence_matches_oracle(pgb, torch_cuda, orc, name):
    """The device-side ConvergenceManager sees the same error sequence as the reference's Mabs."""
    z, A, directed = load_golden(name)
    M = orc.to_sparse_matrix(A, "auto", directed)
    p = z["P"][:, 0]
    _, iters, errs = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
    alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
    alg(_graph(pgb, A, directed), p)
    got = alg.convergence.errors.cpu().numpy()
    assert alg.convergence.iteration == iters
    assert len(got) == len(errs)
    assert np.allclose(got, errs, rtol=1e-9, atol=0)


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
def test_custom_absorption_and_propagate(pgb, torch_cuda, name):
    z, A, directed = load_golden(name)
    g = _graph(pgb, A, directed)
    alg = pgb.AbsorbingWalks(0.9, tol=1e-9, max_iters=1000)
    r = alg(g, z["P"][:, 0], absorption=z["absorption"])
    assert alg.convergence.iteration == int(z["run_absorb90_custom_iters"][0])
    assert rel_l1(r.numpy(), z["run_absorb90_custom_scores"]) <= FP64_TOL
    out = pgb.PageRank(0.85, tol=1e-9, max_iters=1000).propagate(g, z["P"])
    assert rel_l1(out.cpu().numpy(), z["run_propagate_ppr85"]) <= FP64_TOL


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("run", ["ppr85", "ppr90_noq", "ppr85_col", "ppr85_l1", "ppr85_msq", "ppr85_tol6_mod3",
                                 "absorb85", "absorb85_col"])
@pytest.mark.parametrize("relabel", ["hub", "degree", "none"])
@pytest.mark.parametrize("family", ["hsell", "csr"])
def test_batched_propagate_matches_golden(pgb, torch_cuda, monkeypatch, name, run, relabel, family):
    """propagate() through the panel kernels (hsell: pgb_affine_steps_panel on the hub-blocked form; csr:
    pgb_affine_steps_batched on the item stream): every column must stop at the reference's own iteration count and
    match its scores (signals.py:225-226 runs them one by one)."""
    torch = torch_cuda
    monkeypatch.setenv("PGB_PANEL", "1" if family == "hsell" else "csr")
    if family == "hsell":
        monkeypatch.setenv("PGB_HSELL_BLOCK_COLS", "64")      # golden graphs are small: several hub blocks and a tail
        monkeypatch.setenv("PGB_HSELL_BLOCKS", "5")
        monkeypatch.setenv("PGB_HSELL_MIN_ENTRIES", "4")
    z, A, directed = load_golden(name)
    norm, make, _ = _runs(pgb)[run]
    g = _graph(pgb, A, directed, norm, relabel)
    P = z["P"]
    alg = make({"dtype": torch.float64})
    assert alg._panel_family(g) == ("csr" if g.in_view.weighted else family)   # weighted graphs: item-stream panels
    out = alg.propagate(g, P).cpu().numpy()
    assert list(alg.convergence.iterations) == [int(v) for v in z[f"run_{run}_iters"]], (name, run)
    for c in range(P.shape[1]):
        assert rel_l1(out[:, c], z[f"run_{run}_scores"][:, c]) <= FP64_TOL, (name, run, c)
    if run not in ("ppr85", "ppr85_col", "absorb85"):
        return          # tolerances below the fp32 noise floor (L1 1e-7, MSQ 1e-16) are fp64-only cases
    alg32 = make({"dtype": torch.float32})
    out32 = alg32.propagate(g, P).cpu().numpy()
    for c in range(P.shape[1]):
        assert abs(alg32.convergence.iterations[c] - int(z[f"run_{run}_iters"][c])) <= 1
        assert rel_l1(out32[:, c], z[f"run_{run}_scores"][:, c]) <= FP32_TOL, (name, run, c)


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("run", ["heat3", "heat3_tol9", "heat5_sym", "gen40", "gen3_tol", "pprclosed"])
@pytest.mark.parametrize("relabel", ["hub", "none"])
def test_closed_form_propagate_as_panels_matches_golden(pgb, torch_cuda, monkeypatch, name, run, relabel):
    """propagate() of the closed-form filters through the polynomial mode of the hub-blocked panel kernel: every slot
    accumulates coef[k] * power_k at its own step (abstract_filters.py:225-256); golden iteration counts and scores per
    column.  Weighted graphs (no panel form) run column by column and must give the same."""
    torch = torch_cuda
    monkeypatch.setenv("PGB_HSELL_BLOCK_COLS", "64")          # several hub blocks and a tail on the small golden graphs
    monkeypatch.setenv("PGB_HSELL_BLOCKS", "5")
    monkeypatch.setenv("PGB_HSELL_MIN_ENTRIES", "4")
    z, A, directed = load_golden(name)
    norm, make, _ = _runs(pgb)[run]
    g = _graph(pgb, A, directed, norm, relabel)
    P = z["P"]
    alg = make({"dtype": torch.float64})
    assert alg._can_batch(g, n_columns=P.shape[1]) == (not g.in_view.weighted)
    out = alg.propagate(g, P).cpu().numpy()
    assert list(alg.convergence.iterations) == [int(v) for v in z[f"run_{run}_iters"]], (name, run)
    for c in range(P.shape[1]):
        assert rel_l1(out[:, c], z[f"run_{run}_scores"][:, c]) <= FP64_TOL, (name, run, c)
    alg32 = make({"dtype": torch.float32})
    out32 = alg32.propagate(g, P).cpu().numpy()
    for c in range(P.shape[1]):
        assert abs(alg32.convergence.iterations[c] - int(z[f"run_{run}_iters"][c])) <= 1
        assert rel_l1(out32[:, c], z[f"run_{run}_scores"][:, c]) <= FP32_TOL, (name, run, c)


@pytest.mark.parametrize("dtype_name,tol", [("float64", 1e-12), ("float32", 2e-6)])
@pytest.mark.parametrize("family", ["hsell", "hsell_small_blocks", "hsell_groups", "csr"])
def test_batched_propagate_ragged_panels_rmat17(pgb, torch_cuda, monkeypatch, dtype_name, tol, family):
    """11 columns (full panels + a ragged one), one all-zero column, columns that converge at
    different iterations; the batched result must equal the single-column fused path."""
    torch = torch_cuda
    monkeypatch.setenv("PGB_PANEL", "csr" if family == "csr" else "1")
    if family == "hsell_small_blocks":                           # 80 hub blocks of 256 nodes + a long tail
        monkeypatch.setenv("PGB_HSELL_BLOCK_COLS", "256")
    from pygrank_b200 import synthetic, device_synthetic
    dtype = getattr(torch, dtype_name)
    scale = 17
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=4)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="symmetric")
    assert not g.in_view.weighted
    rng = np.random.default_rng(5)
    B = 11
    P = np.zeros((n, B))
    for c in range(B):
        if c == 4:
            continue                                    # zero personalization: returned untouched, 0 iterations
        k = [1, 10, 1000, n // 2][c % 4]
        P[rng.choice(n, k, replace=False), c] = rng.uniform(0.5, 2.0, k)
    alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=dtype)
    if family == "hsell_groups":
        alg.panel_group, alg.panel_chunk = 5, 3             # columns staged in three groups, polled every 3 steps
    out = alg.propagate(g, torch.from_numpy(P).cuda())
    its = list(alg.convergence.iterations)
    assert out.shape == (n, B) and its[4] == 0 and not bool(out[:, 4].any())
    assert len(set(its)) > 2                            # the panel really freezes columns at different steps
    for c in range(B):
        one = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=dtype)
        ref = one(g, P[:, c]).np
        assert abs(one.convergence.iteration - its[c]) <= (0 if dtype == torch.float64 else 1), (c, its)
        assert rel_l1(out[:, c].cpu().numpy(), ref.cpu().numpy()) <= tol, c
    with pytest.raises(Exception, match="Could not converge within 4 iterations"):
        pgb.PageRank(0.99, tol=1e-14, max_iters=4, dtype=dtype).propagate(g, torch.from_numpy(P).cuda())


@pytest.mark.parametrize("dtype_name,tol", [("float64", 1e-12), ("float32", 2e-6)])
def test_alpha_sweep_as_panels_matches_single_solves(pgb, torch_cuda, dtype_name, tol):
    """PageRank.sweep: the candidates of a parameter search (autotune/optimization.py:160-180 evaluates them one by
    one) as columns of the hub-blocked panel kernel, every column with its own alpha, normaliser and stop decision."""
    torch = torch_cuda
    from pygrank_b200 import device_synthetic
    dtype = getattr(torch, dtype_name)
    scale = 16
    n = 1 << scale
    g = device_synthetic.rmat_graph_device(scale, 16, seed=9)
    rng = np.random.default_rng(2)
    p = np.zeros(n)
    p[rng.choice(n, 50, replace=False)] = 1.0
    alphas = [0.5, 0.6, 0.7, 0.8, 0.85, 0.9, 0.95]             # ragged: 7 columns
    for quotient in (True, False):
        alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=dtype, use_quotient=quotient)
        out = alg.sweep(g, p, alphas)
        its = list(alg.convergence.iterations)
        assert out.shape == (n, len(alphas)) and its == sorted(its) and its[0] < its[-1]
        for j, a in enumerate(alphas):
            one = pgb.PageRank(a, tol=1e-9, max_iters=1000, dtype=dtype, use_quotient=quotient)
            ref = one(g, p).np
            assert abs(one.convergence.iteration - its[j]) <= (0 if dtype == torch.float64 else 1), (a, its)
            assert rel_l1(out[:, j].cpu().numpy(), ref.cpu().numpy()) <= tol, a


def test_panel_max_difference_matches_single_solves(pgb, torch_cuda, monkeypatch):
    """error_type=MaxDifference through the hub-blocked panel kernel (a max reduction per column)."""
    torch = torch_cuda
    from pygrank_b200 import device_synthetic
    monkeypatch.setenv("PGB_PANEL", "1")

    class MaxDifference:
        pass

    n = 1 << 15
    g = device_synthetic.rmat_graph_device(15, 16, seed=3)
    rng = np.random.default_rng(1)
    P = np.zeros((n, 5))
    for c in range(5):
        P[rng.choice(n, 10 ** (c % 3 + 1), replace=False), c] = 1.0
    alg = pgb.PageRank(0.85, tol=1e-8, max_iters=1000, error_type=MaxDifference)
    assert alg._panel_family(g) == "hsell"
    out = alg.propagate(g, torch.from_numpy(P).cuda())
    for c in range(5):
        one = pgb.PageRank(0.85, tol=1e-8, max_iters=1000, error_type=MaxDifference)
        ref = one(g, P[:, c]).np
        assert one.convergence.iteration == alg.convergence.iterations[c]
        assert rel_l1(out[:, c].cpu().numpy(), ref.cpu().numpy()) <= 1e-12


@pytest.mark.parametrize("name", GOLDEN_GRAPHS)
@pytest.mark.parametrize("variant", ["hsell_or_stream", "item_stream"])
def test_max_difference_criterion_matches_oracle(pgb, torch_cuda, orc, name, variant):
    """error_type=MaxDifference (measures/supervised.py:93-98): a max reduction instead of a sum, on both
    per-iteration kernels; iteration counts and the error sequence must equal the oracle's."""
    torch = torch_cuda
    from pygrank_b200 import _capi as C
    z, A, directed = load_golden(name)
    M = orc.to_sparse_matrix(A, "auto", directed)
    g = _graph(pgb, A, directed)

    class MaxDifference:      # selected like the reference's class (its __name__ is what the engine maps)
        pass

    lib = C.lib()
    try:
        if variant == "item_stream":
            C.check(lib.pgb_set_kernel_variant(3))
        for c in range(2):
            p = z["P"][:, c]
            ref, iters, errs = orc.pagerank(M, p, 0.85, tol=1e-8, max_iters=1000, error_type="max")
            for et in ("max", MaxDifference):
                alg = pgb.PageRank(0.85, tol=1e-8, max_iters=1000, error_type=et)
                r = alg(g, p)
                assert alg.convergence.iteration == iters, (name, c)
                assert rel_l1(r.numpy(), ref) <= FP64_TOL
                # the maximum sits on the largest scores, whose fp64 rounding noise (1e-17 absolute) is ~1e-9 of
                # an error near the tolerance: compare the sequences to 1e-6, the iteration counts exactly
                assert np.allclose(alg.convergence.errors.cpu().numpy(), errs, rtol=1e-6, atol=1e-300)
            ref, iters, _ = orc.heat_kernel(M, p, 3, tol=1e-8, error_type="max")
            alg = pgb.HeatKernel(3, tol=1e-8, error_type="max")
            assert rel_l1(alg(g, p).numpy(), ref) <= FP64_TOL and alg.convergence.iteration == iters
            a32 = pgb.PageRank(0.85, tol=1e-6, max_iters=1000, error_type="max", dtype=torch.float32)
            ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-6, max_iters=1000, error_type="max")
            r32 = a32(g, p)
            assert abs(a32.convergence.iteration - iters) <= 1 and rel_l1(r32.numpy(), ref) <= FP32_TOL
    finally:
        C.check(lib.pgb_set_kernel_variant(4))


def test_personalization_forms_and_edge_cases(pgb, torch_cuda):
    torch = torch_cuda
    z, A, directed = load_golden("ba2000")
    g = _graph(pgb, A, directed)
    n = A.shape[0]
    seeds = [3, 17, 256]
    dense = np.zeros(n)
    dense[seeds] = 1.0
    a = pgb.PageRank(0.85, tol=1e-9)(g, dense).numpy()
    b = pgb.PageRank(0.85, tol=1e-9)(g, {s: 1 for s in seeds}).numpy()
    c = pgb.PageRank(0.85, tol=1e-9)(g, seeds).numpy()
    d = pgb.PageRank(0.85, tol=1e-9)(g, torch.from_numpy(dense).cuda()).numpy()
    # (grid-level fp64 atomics make the last bit of the normaliser run-dependent)
    assert max(rel_l1(b, a), rel_l1(c, a), rel_l1(d, a)) < 1e-14
    # zero personalization returns it untouched with 0 iterations (abstract_filters.py:53-54)
    alg = pgb.PageRank(0.85)
    r = alg(g, np.zeros(n))
    assert alg.convergence.iteration == 0 and not r.numpy().any()
    # max_iters raises like convergence.py:90
    with pytest.raises(Exception, match="Could not converge within 5 iterations"):
        pgb.PageRank(0.99, tol=1e-14, max_iters=5)(g, dense)
    # preserve_norm scales the output by sum|p| (abstract_filters.py:63-64)
    r3 = pgb.PageRank(0.85, tol=1e-9)(g, dense * 3).numpy()
    assert rel_l1(r3, 3 * a) < 1e-14
    # warm start from the converged answer stops at iteration 2
    alg = pgb.PageRank(0.85, tol=1e-9)
    alg(g, dense, warm_start=a / 3.0)
    assert alg.convergence.iteration == 2
    # wrong length is rejected like signals.py:55-57
    with pytest.raises(Exception, match="should be equal to graph nodes"):
        pgb.PageRank()(g, np.ones(n + 1))


def test_large_rmat_against_oracle(pgb, torch_cuda, orc):
    """RMAT scale 17 built on the device, filters vs the oracle on the same CSR."""
    from pygrank_b200 import synthetic, device_synthetic
    scale = 17
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=4)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="symmetric")
    A = synthetic.rmat_graph_host(scale, 16, seed=4)
    M = orc.to_sparse_matrix(A, "symmetric", False)
    assert np.array_equal(g.to_scipy_normalized().data, M.data)
    seeds = synthetic.seed_sets(n, 2, 10, seed=0)
    for s in seeds:
        p = np.zeros(n)
        p[s] = 1.0
        ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
        alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
        got = alg(g, p).numpy()
        assert alg.convergence.iteration == iters
        assert rel_l1(got, ref) <= FP64_TOL
        ref, iters, _ = orc.heat_kernel(M, p, 3)
        alg = pgb.HeatKernel(3)
        got = alg(g, p).numpy()
        assert alg.convergence.iteration == iters
        assert rel_l1(got, ref) <= FP64_TOL
        alg32 = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch_cuda.float32)
        ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
        got = alg32(g, p).numpy()
        assert abs(alg32.convergence.iteration - iters) <= 1
        assert rel_l1(got, ref) <= FP32_TOL


def test_c_abi_reports_errors(pgb, torch_cuda):
    from pygrank_b200 import _capi as C
    lib = C.lib()
    assert lib.pgb_make_scales(4, 0, 99, 0, 0) != 0
    assert b"unknown kind" in lib.pgb_last_error()
    with pytest.raises(Exception, match="pgb200"):
        C.check(lib.pgb_scale(8, 7, 0, 0, 1.0, 0, 0, 0))
