"""Device postprocessors (pygrank_b200/postprocess.py) against the reference's own classes
(/root/reference/pygrank/algorithms/postprocess/postprocess.py:106-352) on the same vectors, ties included.  The
classes are plain tensor code, so the comparison runs on CPU tensors here and on CUDA tensors on the GPU box."""
import numpy as np
import pytest
import torch

from refutil import import_pygrank


class _G:
    def __init__(self, n):
        from pygrank_b200.graph import IdentityNodeMap
        self.n = n
        self._pygrank_node2id = IdentityNodeMap(n)


def _cases():
    rng = np.random.default_rng(5)
    x = rng.random(257)
    ties = np.round(rng.random(300) * 8) / 8
    return {"random": x, "ties": ties, "constant": np.full(40, 0.25), "with_zero": np.concatenate([[0.0, 0.0], x[:30]])}


@pytest.mark.parametrize("device", ["cpu", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_postprocessors_match_reference(device):
    pg = import_pygrank()
    if pg is None:
        pytest.skip("baseline/_ref is not installed on this box")
    import networkx as nx
    from pygrank_b200 import postprocess as pp
    from pygrank_b200.filters import RankResult
    pg.load_backend("numpy")
    for cname, x in _cases().items():
        n = len(x)
        graph = nx.empty_graph(n)
        sig = pg.to_signal(graph, x.copy())
        mine = RankResult(_G(n), torch.from_numpy(x.copy()).to(device))
        pairs = [
            (pg.Normalize(), pp.Normalize()), (pg.Normalize("sum"), pp.Normalize("sum")),
            (pg.Normalize("range"), pp.Normalize(method="range")), (pg.Normalize("L2"), pp.Normalize("L2")),
            (pg.Ordinals(), pp.Ordinals()), (pg.Top(5), pp.Top(5)), (pg.Top(0.3), pp.Top(0.3)),
            (pg.Top(fraction_of_training=1), pp.Top()), (pg.Threshold(0.5), pp.Threshold(0.5)),
            (pg.Threshold(0.5, inclusive=True), pp.Threshold(0.5, inclusive=True)),
            (pg.Threshold("gap"), pp.Threshold("gap")),
        ]
        for ref_pp, my_pp in pairs:
            ref = ref_pp.transform(sig)
            want = np.array([ref.get(i, 0.0) for i in range(n)], dtype=np.float64)
            got = my_pp.transform(mine).np.cpu().numpy()
            assert np.allclose(got, want, rtol=1e-15, atol=0), (cname, type(my_pp).__name__, getattr(my_pp, "method", ""))


def test_argument_juggling_like_the_reference():
    from pygrank_b200 import postprocess as pp

    class Ranker:
        def rank(self, *a, **k):
            pass
    r = Ranker()
    assert pp.Normalize("sum", r).ranker is r and pp.Normalize("sum", r).method == "sum"
    assert pp.Normalize(r, "sum").ranker is r
    assert pp.Top(3).ranker is None and pp.Top(3).fraction_of_training == 3
    assert pp.Threshold(r, 0.2).threshold == 0.2 and pp.Threshold(0.2, r).ranker is r
    with pytest.raises(Exception):
        pp.Normalize("median").transform(__import__("pygrank_b200").filters.RankResult(_G(3), torch.ones(3)))
