"""Hub-signature node order (pgb_hub_order, include/pgb200.h): a numpy restatement (CPU) and the device
result against it (GPU).  The order is the engine's own choice (the reference keeps the user's labels,
/root/reference/pygrank/core/utils/preprocessing.py:151); what must hold is that it is a permutation that keeps
every node inside its degree region and sorts regions by the presence bits of the hub regions a row touches."""
import numpy as np
import pytest


def hub_order_np(indptr, indices, G, span):
    """perm (new position -> node) of pgb_hub_order for a CSR in original labels."""
    n = len(indptr) - 1
    deg = np.diff(indptr)
    perm0 = np.argsort(-deg, kind="stable")                    # pgb_degree_order: descending degree, stable
    rank = np.empty(n, dtype=np.int64)
    rank[perm0] = np.arange(n)
    bits = -(-span // G)
    words = -(-bits // 64)
    sig = np.zeros((max(words, 1), n), dtype=np.uint64)
    rows = np.repeat(np.arange(n), deg)
    c = rank[indices]
    keep = c < span
    b = c[keep] // G
    key = np.unique(rows[keep] * bits + b) if bits else np.zeros(0, dtype=np.int64)
    r, bb = key // max(bits, 1), key % max(bits, 1)
    np.bitwise_or.at(sig, (bb // 64, r), np.uint64(1) << (np.uint64(63) - (bb % 64).astype(np.uint64)))
    region = np.minimum(rank, span) // G
    keys = [rank]
    for w in range(words - 1, -1, -1):
        keys.append(~sig[w])
    keys.append(region)
    return np.lexsort(keys)


def test_numpy_model_properties():
    from pygrank_b200 import synthetic
    A = synthetic.rmat_graph_host(12, 16, seed=3)
    n = A.shape[0]
    G, span = 256, 1024
    perm = hub_order_np(A.indptr, A.indices, G, span)
    assert sorted(perm.tolist()) == list(range(n))
    deg = np.diff(A.indptr)
    rank = np.empty(n, dtype=np.int64)
    rank[np.argsort(-deg, kind="stable")] = np.arange(n)
    new_pos = np.empty(n, dtype=np.int64)
    new_pos[perm] = np.arange(n)
    # every node stays inside its degree region (hub block membership of the columns is unchanged)
    assert np.array_equal(np.minimum(rank, span) // G, np.minimum(new_pos, span) // G)
    # the slices of the tail are denser per hub region than in plain degree order
    def sparse_pairs(pos):
        rows = np.repeat(pos, deg)
        cols = pos[A.indices]
        m = cols < span
        unit = (rows[m] // 32) * 64 + cols[m] // G
        _, cnt = np.unique(unit, return_counts=True)
        return int(cnt[cnt < 32].sum())
    assert sparse_pairs(new_pos) < 0.6 * sparse_pairs(rank)


@pytest.mark.gpu
@pytest.mark.parametrize("scale,block_cols,blocks", [(12, 256, 4), (14, 1024, 6), (13, 64, 70)])
def test_device_order_matches_numpy_model(monkeypatch, scale, block_cols, blocks):
    import torch
    import pygrank_b200 as pgb
    from pygrank_b200 import synthetic
    from pygrank_b200.graph import hub_signature_shape
    monkeypatch.setenv("PGB_HSELL_BLOCK_COLS", str(block_cols))
    monkeypatch.setenv("PGB_HSELL_BLOCKS", str(blocks))
    A = synthetic.rmat_graph_host(scale, 16, seed=5)
    n = A.shape[0]
    G, span = hub_signature_shape(n)
    assert G == block_cols
    g = pgb.DeviceGraph.from_scipy(A, directed=False, normalization="symmetric", relabel="hub")
    want = hub_order_np(A.indptr, A.indices, G, span)
    assert np.array_equal(g.perm.cpu().numpy().astype(np.int64), want)
    assert np.array_equal(g.iperm.cpu().numpy()[want], np.arange(n))
    # and the relabelled operator is the same operator: conv agrees with scipy in user labels
    x = np.random.default_rng(0).random(n)
    M = g.to_scipy_normalized()
    y = g.conv(torch.from_numpy(x).cuda()).cpu().numpy()
    assert np.abs(y - x @ M).sum() <= 1e-12 * np.abs(y).sum()
