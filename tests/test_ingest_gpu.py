"""Ingestion straight to device COO (pygrank_b200/ingest.py) against the reference's own path: the SNAP pairs loader
(/root/reference/pygrank/benchmarks/loader.py:18-88) feeding fastgraph (fastgraph/fastgraph.py:40-78) feeding
pg.preprocessor (core/utils/preprocessing.py:99-145).  The reference is the unmodified install in baseline/_ref."""
import os

import numpy as np
import pytest

from conftest import rel_l1
from refutil import import_pygrank

pytestmark = pytest.mark.gpu


def _write_dataset(tmp_path, rng, n=300, m=2500, directed=False):
    d = tmp_path / "toy"
    d.mkdir()
    names = ["n%03d" % i for i in rng.permutation(n)]
    lines = ["# toy dataset", ""]
    for _ in range(m):
        u, v = rng.integers(0, n, 2)
        lines.append(f"{names[u]} {names[v]}" if rng.random() < 0.8 else f"{names[u]}\t{names[v]}  ")
    lines += ["lonely", f"{names[0]} {names[1]}", f"{names[0]} {names[1]}", f"{names[5]} {names[5]}"]   # duplicates, a self loop
    (d / "pairs.txt").write_text("\n".join(lines) + "\n", encoding="utf-8")
    groups = ["# groups"]
    for gsz in (40, 2, 25):
        groups.append(" ".join(names[i] for i in rng.choice(n, gsz, replace=False)) + " unknown_node")
    (d / "groups.txt").write_text("\n".join(groups) + "\n", encoding="utf-8")
    return str(tmp_path)


@pytest.mark.parametrize("directed", [False, True])
@pytest.mark.parametrize("norm", ["symmetric", "col", "none"])
def test_snap_loader_matches_reference(tmp_path, directed, norm):
    pg = import_pygrank()
    if pg is None:
        pytest.skip("baseline/_ref is not installed on this box")
    import pygrank_b200
    from pygrank_b200 import ingest
    rng = np.random.default_rng(3)
    root = _write_dataset(tmp_path, rng, directed=directed)
    from pygrank.benchmarks import loader
    orig = loader.download_dataset
    loader.download_dataset = lambda *a, **k: None           # no network: the dataset is already on disk
    try:
        G, groups_ref = loader.import_snap_format_dataset("toy", path=root, directed=directed, min_group_size=3, verbose=False)
    finally:
        loader.download_dataset = orig
    pg.load_backend("numpy")
    M = pg.preprocessor(normalization=norm)(G).array.tocsr()
    M.sort_indices()
    g, groups = ingest.import_snap_format_dataset("toy", path=root, directed=directed, min_group_size=3, normalization=norm)
    assert groups == groups_ref
    assert list(g._pygrank_node2id) == list(G.node_map) and dict(g._pygrank_node2id) == dict(G.node_map)
    got = g.to_scipy_normalized()
    assert np.array_equal(got.indptr, M.indptr) and np.array_equal(got.indices, M.indices)
    assert np.abs(got.data - M.data).max() <= 4e-16 * max(1.0, np.abs(M.data).max())
    # the same fastgraph object handed to the device preprocessor takes the COO route too
    g2 = pygrank_b200.preprocessor(normalization=norm)(G)
    got2 = g2.to_scipy_normalized()
    assert np.array_equal(got2.indptr, M.indptr) and np.array_equal(got2.indices, M.indices)
    assert np.abs(got2.data - M.data).max() <= 4e-16 * max(1.0, np.abs(M.data).max())
    if norm == "none":
        return
    # and ranks through it equal the reference's, node names as keys
    seeds = {list(G.node_map)[i]: 1.0 for i in (3, 10, 77)}
    ref_alg = pg.PageRank(0.85, tol=1e-9, max_iters=1000, normalization=norm)
    ref = ref_alg(G, seeds)
    alg = pygrank_b200.PageRank(0.85, tol=1e-9, max_iters=1000, normalization=norm)
    mine = alg(g, seeds)
    assert alg.convergence.iteration == ref_alg.convergence.iteration
    assert rel_l1(mine.numpy(), np.asarray(ref.np)) <= 1e-10
    some = list(G.node_map)[42]
    assert mine[some] == pytest.approx(ref[some], rel=1e-9)


def test_masked_fastgraph_edges_stay_as_zeros():
    pg = import_pygrank()
    if pg is None:
        pytest.skip("baseline/_ref is not installed on this box")
    from pygrank.fastgraph import fastgraph
    import pygrank_b200
    G = fastgraph.Graph()
    for u, v in [(0, 1), (1, 2), (2, 3), (3, 0), (0, 2), (1, 3)]:
        G.add_edge(u, v)
    G.remove_edge(0, 2)
    A = G.to_scipy_sparse_array().tocsr()
    g = pygrank_b200.preprocessor(normalization="none")(G)
    got = g.to_scipy_normalized()
    assert np.array_equal(got.toarray(), A.toarray())
