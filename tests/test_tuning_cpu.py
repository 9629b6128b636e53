"""AlphaSweep.optimizer restates the reference's coordinate line search (autotune/optimization.py:118-199) so that it
can announce each round's candidates; on any loss it must visit the same candidates in the same order and return the
same point as ``pg.optimize``."""
import pytest

from refutil import import_pygrank


@pytest.fixture(scope="module")
def pg():
    mod = import_pygrank()
    if mod is None:
        pytest.skip("baseline/_ref (the unmodified reference) is not installed on this box")
    return mod


def _beale(p):
    return (1.5 - p[0] + p[0] * p[1]) ** 2 + (2.25 - p[0] + p[0] * p[1] ** 2) ** 2 + (2.625 - p[0] + p[0] * p[1] ** 3) ** 2


@pytest.mark.parametrize("case", [
    dict(max_vals=[4.5, 4.5], min_vals=[-4.5, -4.5]),
    dict(max_vals=[0.99], min_vals=[0.5], deviation_tol=0.01, divide_range=1.5, partitions=7),
    dict(max_vals=[4.5, 4.5], min_vals=[-4.5, -4.5], depth=2, divide_range=2, partitions=4, parameter_tol=1e-3),
    dict(max_vals=[1, 1, 1], min_vals=[0, 0, 0], coarse=0.05, divide_range=1.3, deviation_tol=1e-4),
])
def test_line_search_visits_the_same_candidates(pg, case):
    from pygrank_b200.tuning import AlphaSweep
    sweep = AlphaSweep.__new__(AlphaSweep)            # the optimizer needs no engine (and no GPU)
    sweep._announced = []
    announced = []
    sweep.announce = lambda cands: announced.append([list(c) for c in cands])

    def loss_of(trace):
        def loss(p):
            trace.append(list(p))
            q = list(p) + [0.3] * (2 - len(p))
            return _beale(q[:2]) + sum((v - 0.4) ** 2 for v in p[2:])
        return loss

    ref_trace, got_trace = [], []
    ref = pg.optimize(loss_of(ref_trace), verbose=False, **case)
    got = sweep.optimizer(loss_of(got_trace), verbose=False, **case)
    assert got_trace == ref_trace and len(ref_trace) > 10
    assert list(got) == list(ref)
    flat = [c for grid in announced for c in grid]
    assert flat == got_trace                          # every evaluated candidate was announced first, in order


def test_unsupported_strategies_are_refused():
    from pygrank_b200.tuning import AlphaSweep
    sweep = AlphaSweep.__new__(AlphaSweep)
    sweep._announced = []
    for kw in (dict(shrink_strategy="shrinking"), dict(partition_strategy="step"), dict(randomize=True)):
        with pytest.raises(Exception, match="default search"):
            sweep.optimizer(lambda p: 0.0, max_vals=[1], **kw)
    with pytest.raises(Exception, match="divide_range"):
        sweep.optimizer(lambda p: 0.0, max_vals=[1], divide_range=1)


class _Signal:
    """The three attributes of a pg.GraphSignal the sweep's rankers touch (core/signals.py:38-60)."""

    def __init__(self, graph, obj, node2id=None):
        self.graph, self.np, self.node2id = graph, obj, node2id


class _Filter:
    def __init__(self):
        import torch
        self.dtype = torch.float64
        self.calls = []

    def _device_graph(self, graph):
        return graph

    def sweep(self, g, p, alphas, **kwargs):
        import torch
        self.calls.append(list(alphas))
        return p[:, None] * torch.tensor(alphas, dtype=p.dtype)[None, :]


def test_candidates_of_a_round_share_one_sweep():
    """Announced candidates are evaluated by ONE sweep on the first miss and served from the cache afterwards; a new
    personalization, or an in-place write to the old one, starts over; the final single-alpha ranking is its own solve."""
    import torch
    from pygrank_b200.lazy import LazyVec
    from pygrank_b200.tuning import AlphaSweep
    sweep = AlphaSweep.__new__(AlphaSweep)
    sweep.filter, sweep._announced, sweep.stats = _Filter(), [], {"sweeps": 0, "columns": 0, "served": 0}
    from collections import OrderedDict
    sweep._cache = OrderedDict()
    graph = object()
    p = torch.arange(1.0, 6.0, dtype=torch.float64)
    sig = _Signal(graph, LazyVec.wrap(p), {"a": 0})
    grid = [[0.5], [0.6], [0.7]]
    sweep.announce(grid)
    outs = [sweep.ranker(c).rank(sig) for c in grid]
    assert sweep.filter.calls == [[0.5, 0.6, 0.7]] and sweep.stats == {"sweeps": 1, "columns": 3, "served": 3}
    for c, o in zip(grid, outs):
        assert isinstance(o, _Signal) and o.graph is graph and o.node2id == {"a": 0}
        assert torch.equal(o.np.materialize(), p * c[0])
    # next round: two old candidates, two new ones -> one sweep with the new ones only (the missed one first)
    sweep.announce([[0.6], [0.65], [0.7], [0.75]])
    sweep.ranker([0.6]).rank(sig)
    sweep.ranker([0.65]).rank(sig)
    assert sweep.filter.calls[-1] == [0.65, 0.75] and sweep.stats["sweeps"] == 2
    # the tuner's final ranking: a parameter nobody announced
    sweep.announce([])
    sweep.ranker([0.9]).rank(graph, sig)
    assert sweep.filter.calls[-1] == [0.9]
    # an in-place write to the personalization invalidates what was cached for it
    p[0] = 100.0
    sweep.announce([[0.5]])
    out = sweep.ranker([0.5]).rank(sig)
    assert sweep.filter.calls[-1] == [0.5] and float(out.np.materialize()[0]) == 50.0
    with pytest.raises(Exception, match="graph signal"):
        sweep.ranker([0.5]).rank(graph, [1.0, 2.0])
