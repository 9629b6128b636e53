"""The ``b200`` backend module for pygrank's plugin surface.

Implements every function of the backend contract
(/root/reference/pygrank/core/backend/specification.py:5-117) on CUDA tensors, with ``conv``,
``degrees`` and ``scipy_sparse_to_backend`` routed through libpgb200 (hand-written sm_100a
kernels, include/pgb200.h).  ``pygrank_b200.install()`` registers it so that
``pg.load_backend("b200")`` / ``with pg.Backend("b200")`` work and ``pg.PageRank``,
``pg.HeatKernel``, ``pg.GenericGraphFilter``, ``pg.AbsorbingWalks`` and ``ConvergenceManager``
run unchanged on top of it.  The module also works stand-alone (it does not import pygrank).

Vectors are 1-D ``torch.Tensor`` on the current CUDA device (fp64 by default — the numpy
backend's precision, numpy.py:84-86 — or fp32 after ``configure(dtype=torch.float32)``);
graphs are :class:`pygrank_b200.graph.DeviceGraph`.  Elementwise helpers are thin torch calls
(K6 of SURVEY §2.2: not on the roofline path); there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _capi as C
from .graph import DeviceGraph, _dev

_config = {"dtype": torch.float64, "relabel": "hub"}


def configure(dtype=None, relabel=None):
    """Engine options (the spec has no per-call option channel)."""
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
def scipy_sparse_to_backend(M):
    """Upload point of the unchanged ``pg.preprocessor`` route (preprocessing.py:144): M is the
    already-normalised scipy matrix, kept with its explicit values (normalization 'none')."""
    if isinstance(M, DeviceGraph):
        return M
    return DeviceGraph.from_scipy(M, directed=True, normalization="none", relabel=_config["relabel"])


def graph_dropout(M, dropout):
    if dropout == 0:
        return M
    return M.dropout(dropout)


def conv(signal, M):
    """``signal @ M`` (numpy.py:64-65) on the fused merge-path gather kernel."""
    return M.conv(to_array(signal))


def degrees(M):
    """Row sums of the (normalised) matrix (numpy.py:76-77)."""
    return M.degrees(_dtype())


# ---- conversions ----------------------------------------------------------------------------------
def to_array(obj, copy_array=False):
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda and obj.dtype == _dtype() and (obj.dim() == 1 or (obj.dim() == 2 and obj.shape[1] == 1)):
            if copy_array:
                return obj.clone().reshape(-1)
            return obj if obj.dim() == 1 else obj.reshape(-1)
        return obj.to(device=_dev(), dtype=_dtype()).reshape(-1)
    return torch.as_tensor(np.array(obj, dtype=np.float64), dtype=_dtype()).to(_dev()).reshape(-1)


def to_primitive(obj):
    if isinstance(obj, torch.Tensor):
        return obj.to(device=_dev(), dtype=_dtype())
    return torch.as_tensor(np.array(obj, dtype=np.float64), dtype=_dtype()).to(_dev())


def cast(x):
    return x.to(_dtype()) if isinstance(x, torch.Tensor) else x


def is_array(obj):
    return isinstance(obj, (list, np.ndarray, torch.Tensor))


def length(x):
    if isinstance(x, torch.Tensor) or isinstance(x, np.ndarray):
        return int(x.shape[0]) if x.ndim == 1 else int(np.prod(x.shape))
    return len(x)


# ---- elementwise / reductions (numpy.py:2) -------------------------------------------------------
def abs(x):
    return torch.abs(x)


def exp(x):
    return torch.exp(x)


def log(x):
    return torch.log(x)


def copy(x):
    return x.clone()


def sum(x, axis=None):
    if isinstance(x, DeviceGraph):
        raise Exception("sum over a device graph is not part of the hot path")
    return torch.sum(x) if axis is None else torch.sum(x, dim=axis)


def mean(x, axis=None):
    return torch.mean(x) if axis is None else torch.mean(x, dim=axis)


def min(x, axis=None):
    return torch.min(x) if axis is None else torch.min(x, dim=axis).values


def max(x, axis=None):
    return torch.max(x) if axis is None else torch.max(x, dim=axis).values


def dot(x, y):
    return torch.sum(x * y)


def ones(dims):
    return torch.ones(dims, dtype=_dtype(), device=_dev())


def eye(dims):
    return torch.eye(dims, dtype=_dtype(), device=_dev())


def diag(diagonal, offset=0):
    return torch.diagflat(diagonal, offset=offset)


def repeat(value, times):
    return torch.full((int(times),), float(value), dtype=_dtype(), device=_dev())


def self_normalize(obj):
    s = torch.sum(torch.abs(obj))
    return obj / s if s != 0 else obj


def filter_out(x, exclude):
    return x[exclude == 0]


def separate_cols(x):
    return [x[:, c] for c in range(x.shape[1])]


def combine_cols(cols):
    cols = [c.reshape(-1, 1) if c.dim() < 2 else c for c in cols]
    return torch.cat(cols, dim=1)


def epsilon():
    # fp64 eps in both modes: returning the fp32 eps would snap tol=1e-9 up to 1.19e-7 through
    # convergence.py:101 and change iteration counts (as it does for the reference's torch backends)
    return float(np.finfo(float).eps)
