"""Device-side result extraction: the postprocessors that sit right after the filters, as O(n) /
O(n log n) CUDA passes on the score vector instead of per-element host loops over ``float(x[i])``.

Same names, argument juggling and outcomes as the reference (paths under
/root/reference/pygrank/algorithms/postprocess/postprocess.py): ``Normalize`` :106-160, ``Ordinals`` :163-192,
``Top`` :237-285, ``Threshold`` :288-352.  They wrap the fused filters of :mod:`pygrank_b200.filters` (or nothing,
acting as ``Tautology``) and return :class:`RankResult` objects; ties are broken like Python's stable
``sorted(..., reverse=True)`` over node order.
"""
from __future__ import annotations

import torch

from .filters import RankResult


def _split_args(ranker, other):
    """The reference's constructors accept (ranker, parameter) in either order (postprocess.py:132-135)."""
    if ranker is not None and not callable(getattr(ranker, "rank", None)):
        ranker, other = other, ranker
        if not callable(getattr(ranker, "rank", None)):
            ranker = None
    return ranker, other


class Postprocessor:
    def __init__(self, ranker=None):
        self.ranker = ranker

    def rank(self, *args, **kwargs) -> RankResult:
        if self.ranker is None:
            raise Exception("this postprocessor wraps no ranker: use transform(result)")
        return self.transform(self.ranker.rank(*args, **kwargs))

    __call__ = rank

    def transform(self, ranks: RankResult) -> RankResult:
        if not isinstance(ranks, RankResult):
            raise Exception("device postprocessors transform the RankResult of a pygrank_b200 filter")
        return RankResult(ranks.graph, self._transform(ranks.np))

    def _transform(self, x: torch.Tensor) -> torch.Tensor:
        raise Exception("_transform method not implemented for the class " + self.__class__.__name__)

    @property
    def convergence(self):
        return self.ranker.convergence


class Normalize(Postprocessor):
    """Divide by the max (default), the sum, the L2 norm, or map to the [0, 1] range (postprocess.py:137-156)."""

    def __init__(self, ranker=None, method="max"):
        ranker, method = _split_args(ranker, method)
        super().__init__(ranker)
        self.method = method

    def _transform(self, x):
        min_rank = 0.0
        if self.method == "range":
            max_rank, min_rank = float(x.max()), float(x.min())
        elif self.method == "max":
            max_rank = float(x.max())
        elif self.method == "sum":
            max_rank = float(x.sum(dtype=torch.float64))
        elif self.method == "L2":
            max_rank = float((x.to(torch.float64) ** 2).sum()) ** 0.5
        else:
            raise Exception("Can only normalize towards max, sum, range, or L2")
        if min_rank == max_rank:
            return x
        return (x - min_rank) / (max_rank - min_rank)


def _descending_order(x: torch.Tensor) -> torch.Tensor:
    """Node indices from the highest score down; equal scores keep node order, like sorted(ranks, key=ranks.get,
    reverse=True) over a signal's node iteration order."""
    return torch.sort(x, descending=True, stable=True).indices


class Ordinals(Postprocessor):
    """Highest score -> 1, second highest -> 2, ... (postprocess.py:186-188)."""

    def _transform(self, x):
        order = _descending_order(x)
        out = torch.empty_like(x)
        out[order] = torch.arange(1, x.numel() + 1, device=x.device, dtype=x.dtype)
        return out


class Top(Postprocessor):
    """1 for the scores at or above the k-th highest, 0 below (postprocess.py:266-277; k = fraction * n when the
    argument is below 1)."""

    def __init__(self, ranker=None, fraction_of_training=1):
        ranker, fraction_of_training = _split_args(ranker, fraction_of_training)
        super().__init__(ranker)
        self.fraction_of_training = fraction_of_training

    def _transform(self, x):
        k = self.fraction_of_training * x.numel() if self.fraction_of_training < 1 else self.fraction_of_training
        k = int(k)
        threshold = 0.0
        if 1 <= k <= x.numel():
            threshold = torch.topk(x, k).values[-1]
        return (x >= threshold).to(x.dtype)


class Threshold(Postprocessor):
    """Binary scores around a threshold; "gap" places it at the largest relative drop between consecutive sorted
    scores (postprocess.py:330-351)."""

    def __init__(self, ranker=None, threshold=0, inclusive: bool = False):
        ranker, threshold = _split_args(ranker, threshold)
        super().__init__(ranker)
        self.threshold = threshold
        self.inclusive = inclusive

    def _transform(self, x):
        threshold = self.threshold
        if threshold == "gap":
            s = torch.sort(x.to(torch.float64), descending=True, stable=True).values
            prev, cur = s[:-1], s[1:]
            diff = torch.where(prev > 0, (prev - cur) / torch.where(prev > 0, prev, torch.ones_like(prev)),
                               torch.zeros_like(prev))
            threshold = 0.0
            if diff.numel() and float(diff.max()) > 0:
                threshold = float(cur[int(torch.argmax(diff))])        # argmax: the first of equal maxima, as the loop keeps it
        if self.inclusive:
            return (x >= threshold).to(x.dtype)
        return (x > threshold).to(x.dtype)
