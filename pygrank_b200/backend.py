"""The ``b200`` backend module for pygrank's plugin surface.

Implements every function of the backend contract
(/root/reference/pygrank/core/backend/specification.py:5-117) on CUDA tensors, with ``conv``,
``degrees`` and ``scipy_sparse_to_backend`` routed through libpgb200 (hand-written sm_100a
kernels, include/pgb200.h).  ``pygrank_b200.install()`` registers it so that
``pg.load_backend("b200")`` / ``with pg.Backend("b200")`` work and ``pg.PageRank``,
``pg.HeatKernel``, ``pg.GenericGraphFilter``, ``pg.AbsorbingWalks`` and ``ConvergenceManager``
run unchanged on top of it.  The module also works stand-alone (it does not import pygrank).

Vectors are :class:`pygrank_b200.lazy.LazyVec` — deferred expressions over 1-D CUDA tensors (fp64 by default — the
numpy backend's precision, numpy.py:84-86 — or fp32 after ``configure(dtype=torch.float32)``) — and reductions are
:class:`pygrank_b200.lazy.LazyScalar`: the drivers' op-by-op arithmetic is recorded, matched against the iteration
shapes the engine fuses and run as ``pgb_affine_steps`` / ``pgb_poly_steps`` (see lazy.py); whatever does not match
is evaluated eagerly with thin torch calls.  ``configure(lazy=False)`` restores plain tensors everywhere (round-1
behaviour: every op its own kernel).  Graphs are :class:`pygrank_b200.graph.DeviceGraph`; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi as C
from . import lazy as _lazy
from .graph import DeviceGraph, _dev
from .lazy import LazyScalar, LazyVec

_config = {"dtype": torch.float64, "relabel": "hub", "lazy": True}


def configure(dtype=None, relabel=None, lazy=None):
    """Engine options (the spec has no per-call option channel)."""
    if lazy is not None:
        _config["lazy"] = bool(lazy)
    if dtype is not None:
        if dtype not in (torch.float32, torch.float64):
            raise Exception("dtype must be torch.float32 or torch.float64")
        _config["dtype"] = dtype
    if relabel is not None:
        from .graph import RELABELS
        if relabel not in RELABELS:
            raise Exception("relabel must be one of " + ", ".join(RELABELS))
        _config["relabel"] = relabel


def _dtype():
    return _config["dtype"]


# ---- identity / lifecycle (specification.py:5-10) ------------------------------------------
def backend_name():
    return "b200"


def backend_init():
    C.lib()          # fail loudly here if the CUDA library is missing
    _dev()


# ---- graph side ---------------------------------------------------------------------------------
def _structural_normalization(M):
    """Name of the reference normalisation (preprocessing.py:109-138) that turns the 0/1 pattern of ``M`` into exactly
    ``M`` — bit for bit — or None.  ``pg.preprocessor`` hands this backend the already normalised matrix; when its
    values are nothing but degree scalings of an unweighted graph the engine keeps the normalisation factorised
    (graph.py) and the gather streams no edge values: the hub-blocked kernels instead of the weighted item stream."""
    n = M.shape[0]
    if M.nnz == 0 or M.shape[0] != M.shape[1]:
        return None
    data = np.asarray(M.data, dtype=np.float64)
    if bool(np.all(data == 1.0)):
        return None                                     # genuinely unweighted: uploaded as is
    rowsum = np.diff(M.indptr).astype(np.float64)
    colsum = np.bincount(M.indices, minlength=n).astype(np.float64)
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(M.indptr))

    def inv(s, root):
        s = np.sqrt(s) if root else s.copy()
        s[s != 0] = 1.0 / s[s != 0]
        return s

    for name, left, right in (("symmetric", inv(rowsum, True), inv(colsum, True)), ("col", inv(rowsum, False), None),
                              ("both", inv(rowsum, False), inv(colsum, False))):
        cand = left[rows] * 1.0
        if right is not None:
            cand = cand * right[M.indices]
        if np.array_equal(cand, data):
            return name
    return None


def scipy_sparse_to_backend(M):
    """Upload point of the unchanged ``pg.preprocessor`` route (preprocessing.py:144): M is the already normalised
    scipy matrix.  Degree normalisations of an unweighted graph are recognised (bit-exactly) and kept factorised;
    anything else keeps its explicit values (normalization 'none')."""
    if isinstance(M, DeviceGraph):
        return M
    import scipy.sparse as sp
    M = sp.csr_matrix(M)
    name = _structural_normalization(M)
    if name is not None:
        pattern = sp.csr_matrix((np.ones(M.nnz, dtype=np.float64), M.indices, M.indptr), shape=M.shape)
        return DeviceGraph.from_scipy(pattern, directed=True, normalization=name, relabel=_config["relabel"])
    return DeviceGraph.from_scipy(M, directed=True, normalization="none", relabel=_config["relabel"])


def graph_dropout(M, dropout):
    if dropout == 0:
        return M
    return M.dropout(dropout)


def _vec(t):
    """Backend vector for a 1-D tensor."""
    return LazyVec.wrap(t) if _config["lazy"] else t


def _t(x):
    """Plain tensor of a backend vector / tensor."""
    return x.materialize() if isinstance(x, LazyVec) else x


def to_tensor(x) -> torch.Tensor:
    """The plain CUDA tensor behind a backend vector (forces a deferred expression); not part of the spec."""
    return _t(to_array(x))


def conv(signal, M):
    """``signal @ M`` (numpy.py:64-65): recorded; runs fused with the surrounding update when the driver's loop is one
    the engine knows (lazy.py), else on the gather kernels by itself."""
    x = to_array(signal)
    if isinstance(x, LazyVec):
        return LazyVec("conv", (x, M), x.n, x.dtype)
    return M.conv(x)


def degrees(M):
    """Row sums of the (normalised) matrix (numpy.py:76-77)."""
    return _vec(M.degrees(_dtype()))


# ---- conversions ----------------------------------------------------------------------------------
def to_array(obj, copy_array=False):
    if isinstance(obj, LazyVec):
        if obj.dtype != _dtype():
            return _vec(obj.materialize().to(_dtype()))
        return _vec(obj.materialize().clone()) if copy_array else obj
    if isinstance(obj, LazyScalar):
        obj = [obj.value()]
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda and obj.dtype == _dtype() and (obj.dim() == 1 or (obj.dim() == 2 and obj.shape[1] == 1)):
            if copy_array:
                return _vec(obj.clone().reshape(-1))
            return _vec(obj if obj.dim() == 1 else obj.reshape(-1))
        return _vec(obj.to(device=_dev(), dtype=_dtype()).reshape(-1))
    return _vec(torch.as_tensor(np.array(obj, dtype=np.float64), dtype=_dtype()).to(_dev()).reshape(-1))


def to_primitive(obj):
    if isinstance(obj, LazyVec):
        return obj
    if isinstance(obj, LazyScalar):
        return obj.value()
    if isinstance(obj, torch.Tensor):
        return obj.to(device=_dev(), dtype=_dtype())
    return torch.as_tensor(np.array(obj, dtype=np.float64), dtype=_dtype()).to(_dev())


def cast(x):
    if isinstance(x, LazyVec):
        return x if x.dtype == _dtype() else _vec(x.materialize().to(_dtype()))
    return x.to(_dtype()) if isinstance(x, torch.Tensor) else x


def is_array(obj):
    return isinstance(obj, (list, np.ndarray, torch.Tensor, LazyVec))


def length(x):
    if isinstance(x, LazyVec):
        return x.n
    if isinstance(x, torch.Tensor) or isinstance(x, np.ndarray):
        return int(x.shape[0]) if x.ndim == 1 else int(np.prod(x.shape))
    return len(x)


# ---- elementwise / reductions (numpy.py:2) -------------------------------------------------------
def _eager1(fn, x):
    if isinstance(x, LazyVec):
        _lazy.STATS["eager_ops"] += 1
        return _vec(fn(x.materialize()))
    if isinstance(x, LazyScalar):
        return float(fn(torch.tensor(x.value(), dtype=torch.float64)))
    return fn(x)


def abs(x):
    if isinstance(x, (LazyVec, LazyScalar)):
        return x.__abs__()
    return torch.abs(x)


def exp(x):
    return _eager1(torch.exp, x)


def log(x):
    return _eager1(torch.log, x)


def copy(x):
    if isinstance(x, LazyVec):
        return _vec(x.materialize().clone())
    return x.clone()


def sum(x, axis=None):
    if isinstance(x, DeviceGraph):
        raise Exception("sum over a device graph is not part of the hot path")
    if isinstance(x, LazyVec):
        return LazyScalar(("sum", x))
    if isinstance(x, LazyScalar) or not isinstance(x, torch.Tensor):
        return x
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


def mean(x, axis=None):
    if isinstance(x, LazyVec):
        return LazyScalar(("mean", x))
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis)


def min(x, axis=None):
    if isinstance(x, LazyVec):
        return LazyScalar(("min", x))
    return torch.min(x) if axis is None else torch.min(x, dim=axis).values


def max(x, axis=None):
    if isinstance(x, LazyVec):
        return LazyScalar(("max", x))
    return torch.max(x) if axis is None else torch.max(x, dim=axis).values


def dot(x, y):
    if isinstance(x, LazyVec) and isinstance(y, LazyVec):
        return LazyScalar(("dot", x, y))
    return torch.sum(_t(x) * _t(y))


def ones(dims):
    t = torch.ones(dims, dtype=_dtype(), device=_dev())
    return _vec(t) if t.dim() == 1 else t


def eye(dims):
    return torch.eye(dims, dtype=_dtype(), device=_dev())


def diag(diagonal, offset=0):
    return torch.diagflat(_t(diagonal), offset=offset)


def repeat(value, times):
    return _vec(torch.full((int(times),), float(value), dtype=_dtype(), device=_dev()))


def self_normalize(obj):
    t = _t(obj)
    s = torch.sum(torch.abs(t))
    return _vec(t / s if s != 0 else t)


def filter_out(x, exclude):
    x, exclude = _t(x), _t(exclude)
    return _vec(x[exclude == 0])


def separate_cols(x):
    return [_vec(x[:, c].contiguous()) for c in range(x.shape[1])]


def combine_cols(cols):
    cols = [_t(c) for c in cols]
    cols = [c.reshape(-1, 1) if c.dim() < 2 else c for c in cols]
    return torch.cat(cols, dim=1)


def epsilon():
    # fp64 eps in both modes: returning the fp32 eps would snap tol=1e-9 up to 1.19e-7 through
    # convergence.py:101 and change iteration counts (as it does for the reference's torch backends)
    return float(np.finfo(float).eps)
