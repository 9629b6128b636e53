"""Tuner batching: the candidates of a parameter search as columns of the hub-blocked panel kernel.

``pg.ParameterTuner`` (/root/reference/pygrank/algorithms/autotune/parameterized.py:117-167) asks its optimizer for the
best parameters of ``ranker_generator(params)``; the stock optimizer (autotune/optimization.py:118-199) walks the
parameters one at a time and, per round, evaluates ``partitions`` equally spaced candidates ONE AFTER THE OTHER — for
PageRank's alpha that is ``partitions`` full solves on the same personalization per round, each streaming the graph.

:class:`AlphaSweep` plugs into the UNMODIFIED tuner through the two hooks it already has::

    sweep = pygrank_b200.AlphaSweep(tol=1e-9, max_iters=1000)
    tuner = pg.ParameterTuner(sweep.ranker, optimizer=sweep.optimizer, measure=pg.AUC,
                              max_vals=[0.99], min_vals=[0.5], deviation_tol=0.01)

* ``sweep.optimizer`` is the same coordinate line search as the stock one (same candidates, same order, same stop rule),
  except that it announces every round's candidate list before evaluating it;
* ``sweep.ranker(params)`` returns a ranker whose ``rank`` serves the announced candidates from ONE
  ``PageRank.sweep`` call per (round, personalization): every candidate is a column of the panel kernel
  (pgb_affine_steps_panel) with its own alpha, normaliser and stop decision, so the graph is streamed once per panel
  iteration instead of once per candidate.  The measure, the training/validation split and the final ranking stay the
  reference's own code.
"""
from __future__ import annotations

from collections import OrderedDict
import torch

from .filters import PageRank


class _Candidate:
    """What ``ranker_generator(params)`` returns: ``rank`` like NodeRanking.rank (core/signals.py:196-223), served from
    the sweep's cache."""

    def __init__(self, owner: "AlphaSweep", alpha: float):
        self.owner, self.alpha = owner, float(alpha)

    def rank(self, graph=None, personalization=None, *args, **kwargs):
        if personalization is None and hasattr(graph, "graph") and hasattr(graph, "np"):
            graph, personalization = graph.graph, graph                      # rank(signal), the tuner's call
        if not (hasattr(personalization, "graph") and hasattr(personalization, "np")):
            raise Exception("AlphaSweep rankers take a graph signal (pg.to_signal(graph, data))")
        if graph is None:
            graph = personalization.graph
        scores = self.owner._scores(graph, personalization, self.alpha, **kwargs)
        from .lazy import LazyVec
        return personalization.__class__(personalization.graph, LazyVec.wrap(scores), personalization.node2id)

    __call__ = rank

    def propagate(self, *args, **kwargs):
        raise Exception("AlphaSweep rankers only rank")

    def cite(self):
        return "personalized PageRank \\cite{page1999pagerank} with restart probability " + str(1 - self.alpha)

    def references(self):
        return [self.cite()]


class AlphaSweep:
    """Ranker generator + optimizer for ``pg.ParameterTuner`` over PageRank's ``alpha`` (module docstring).  Keyword
    arguments go to :class:`pygrank_b200.PageRank` (tol, max_iters, error_type, dtype, normalization, ...)."""

    def __init__(self, **filter_kwargs):
        filter_kwargs.setdefault("assume_immutability", True)     # like the tuner's own default preprocessor
        self.filter = PageRank(0.85, **filter_kwargs)
        self._announced: list = []
        self._cache: "OrderedDict[tuple, dict]" = OrderedDict()
        self.stats = {"sweeps": 0, "columns": 0, "served": 0}

    # -- hook 1: ranker_generator -------------------------------------------------------------------------------------
    def ranker(self, params) -> _Candidate:
        return _Candidate(self, params[0])

    # -- hook 2: optimizer --------------------------------------------------------------------------------------------
    def optimizer(self, loss, max_vals=(1,), min_vals=None, deviation_tol: float = 1.E-9, divide_range: float = 1.01,
                  partitions: int = 5, parameter_tol: float = float("inf"), depth: int = 1, coarse: float = 0,
                  shrink_strategy: str = "divide", partition_strategy: str = "split", randomize: bool = False,
                  weights=None, verbose: bool = False, validation_loss=None):
        """The stock coordinate line search (autotune/optimization.py:118-199: per round the range of one parameter
        shrinks by ``divide_range``, ``partitions`` equally spaced candidates around the current point are evaluated,
        the search moves to the best; it ends when the loss moved by at most ``deviation_tol`` and every range is within
        ``parameter_tol``), announcing each round's candidates to the sweep before they are evaluated."""
        if shrink_strategy != "divide" or partition_strategy != "split" or randomize or validation_loss is not None:
            raise Exception("AlphaSweep.optimizer implements the default search (shrink_strategy='divide', "
                            "partition_strategy='split', cyclic order, no validation loss); use pg.optimize otherwise")
        hi = [float(v) for v in max_vals]
        lo = [0.0] * len(hi) if min_vals is None else [float(v) for v in min_vals]
        for a, b in zip(lo, hi):
            if a > b:
                raise Exception("Empty parameter range [" + str(a) + "," + str(b) + "]")
        if divide_range <= 1:
            raise Exception("divide_range should be greater than 1, otherwise the search space never shrinks.")
        point = [(a + b) / 2 for a, b in zip(lo, hi)] if weights is None else list(weights)
        radius = [(b - a) / 2 for a, b in zip(lo, hi)]
        moved = [float("inf")] * len(hi)
        best = float("inf")
        var = 0
        while max(radius) != 0:
            radius[var] /= divide_range
            if radius[var] == 0:
                moved[var] = 0
                var = (var + 1) % len(hi)
                continue
            grid = []
            for part in range(partitions):
                cand = list(point)
                cand[var] = min(hi[var], max(lo[var], point[var] + radius[var] * (part * 2. / (partitions - 1) - 1)))
                if coarse != 0:
                    cand[var] = round(cand[var] / coarse) * coarse
                grid.append(cand)
            self.announce(grid)
            losses = [loss(cand) for cand in grid]
            k = min(range(len(grid)), key=lambda j: losses[j])
            point, previous, best = grid[k], best, losses[k]
            moved[var] = abs(previous - best)
            if max(moved) <= deviation_tol and max(radius) <= parameter_tol:
                break
            var = (var + 1) % len(hi)
        self.announce([])
        if depth > 1:
            return self.optimizer(loss, max_vals, min_vals, deviation_tol, divide_range, partitions, parameter_tol,
                                  depth - 1, coarse, shrink_strategy, partition_strategy, randomize, point, verbose,
                                  validation_loss)
        return point

    def announce(self, candidates) -> None:
        """The parameter vectors about to be evaluated (alpha is params[0])."""
        self._announced = [float(c[0]) for c in candidates]

    # -- the batched evaluation ---------------------------------------------------------------------------------------
    def _scores(self, graph, personalization, alpha: float, **kwargs) -> torch.Tensor:
        from . import backend
        p = backend.to_tensor(personalization.np)
        key = (id(graph), p.data_ptr(), int(p._version), tuple(sorted(kwargs)))
        bucket = self._cache.get(key)
        if bucket is None:
            bucket = {"keep": (graph, p), "cols": {}}
            self._cache[key] = bucket
            while len(self._cache) > 4:
                self._cache.popitem(last=False)
        cols = bucket["cols"]
        if alpha not in cols:
            todo = [alpha] + [a for a in self._announced if a != alpha and a not in cols]
            todo = list(OrderedDict.fromkeys(todo))
            g = self.filter._device_graph(graph)
            out = self.filter.sweep(g, p.to(self.filter.dtype), todo, **kwargs)
            self.stats["sweeps"] += 1
            self.stats["columns"] += len(todo)
            for j, a in enumerate(todo):
                cols[a] = out[:, j].contiguous()
        self.stats["served"] += 1
        return cols[alpha]
