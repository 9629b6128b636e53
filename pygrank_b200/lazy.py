"""Deferred vectors: how the UNMODIFIED reference drivers reach the fused kernels through the backend plugin.

pygrank drives a filter one backend call at a time — ``conv``, ``*``, ``+``, ``sum``, ``/``, then ``Mabs`` as
``sum(abs(prev - cur)) / length`` and a host comparison (/root/reference/pygrank/algorithms/filters/adhoc.py:34-36,
abstract_filters.py:126-136,225-256, algorithms/convergence.py:77-101, measures/supervised.py:93-130,
core/signals.py:114-178).  Executed eagerly that is ~8 elementwise kernels and two host synchronisations per
iteration around one gather.  Here every backend vector is a :class:`LazyVec` — an expression node that records
the operation instead of running it — and every reduction a :class:`LazyScalar`.  Nothing runs until the driver needs
a host value (a comparison, ``float()``, an element), and at that point the expression is matched against the two
iteration shapes the engine fuses:

* affine recursion ``x' = (A o conv(x, M) + B) [/ sum(.)]`` with ``A`` a scalar and/or a loop-invariant vector and
  ``B`` a loop-invariant vector — PageRank (adhoc.py:34-36), AbsorbingWalks (adhoc.py:166-169), with or without
  RecursiveGraphFilter's quotient — which becomes an :class:`AffineRun` on ``pgb_affine_steps``: once the driver's
  first convergence test has shown its measure and threshold, the run goes AHEAD of the driver with the stop decision
  on the device, and later tests are answered from the device-side error history without touching the GPU;
* polynomial accumulation ``res' = res + c_k * pow_k ; pow_{k+1} = conv(pow_k, M)`` — HeatKernel, PageRankClosed,
  GenericGraphFilter (abstract_filters.py:225-256) — a :class:`PolyRun` on ``pgb_poly_steps``: one fused launch per
  tested iteration, or all iterations in one call when the driver never tests (``error_type="iters"``).

Anything else (Chebyshev/Krylov recursions, postprocessor quotients, measures other than Mabs/L1/MSQ/MaxDifference,
laplacian operators, user arithmetic) is evaluated eagerly, node by node, with the same torch calls round 1 used:
correct, just not fused.  ``STATS`` counts both kinds so the tests can assert which path ran.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np
import torch

from . import _capi as C

FUSE = True      # False: never start a fused run (every node evaluates eagerly); debugging / CPU tests of the eager path
STATS = {"eager_ops": 0, "eager_convs": 0, "fused_steps": 0, "runs": 0, "recomputed_runs": 0, "syncs": 0,
         "step_waits": 0}   # syncs: full stream synchronisations; step_waits: waits for the event of ONE step


def reset_stats():
    for k in STATS:
        STATS[k] = 0


def _is_number(x) -> bool:
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool)


# ------------------------------------------------------------------------------------------------------------------
# scalars
# ------------------------------------------------------------------------------------------------------------------
class LazyScalar:
    """A reduction (or arithmetic on reductions) that has not been computed yet.  ``tree`` is a nested tuple:
    ("sum", vec) | ("max", vec) | ("min", vec) | ("dot", a, b) | ("mul"|"div"|"add"|"sub"|"pow", x, y) | ("neg", x)
    with floats and LazyVec leaves."""
    __slots__ = ("tree", "_value")

    def __init__(self, tree):
        self.tree = tree
        self._value = None

    # -- evaluation -------------------------------------------------------------------------------------------
    def value(self) -> float:
        if self._value is None:
            self._value = float(_eval_scalar(self.tree))
        return self._value

    def __float__(self):
        return self.value()

    def __bool__(self):
        return self.value() != 0

    def item(self):
        return self.value()

    # -- arithmetic stays symbolic ------------------------------------------------------------------------------
    def _bin(self, op, other, swap=False):
        if isinstance(other, LazyScalar):
            other = other.tree
        elif isinstance(other, torch.Tensor) and other.dim() == 0:
            other = float(other)
        elif not _is_number(other):
            return NotImplemented
        return LazyScalar((op, other, self.tree) if swap else (op, self.tree, other))

    def __add__(self, o): return self._bin("add", o)
    def __radd__(self, o): return self._bin("add", o, True)
    def __sub__(self, o): return self._bin("sub", o)
    def __rsub__(self, o): return self._bin("sub", o, True)
    def __mul__(self, o):
        if isinstance(o, LazyVec):
            return o * self
        return self._bin("mul", o)
    def __rmul__(self, o): return self._bin("mul", o, True)
    def __truediv__(self, o): return self._bin("div", o)
    def __rtruediv__(self, o): return self._bin("div", o, True)
    def __pow__(self, o): return self._bin("pow", o)
    def __neg__(self): return LazyScalar(("neg", self.tree))
    def __abs__(self): return LazyScalar(("abs", self.tree))

    # -- host decisions: the points where the driver forces the device ---------------------------------------------
    def _cmp(self, other, op):
        if isinstance(other, LazyScalar):
            other = other.value()
        other = float(other)
        err = _match_error(self.tree)
        if err is not None and self._value is None:
            self._value = err.resolve(hint_threshold=other if op in ("le", "lt") else None)
        v = self.value()
        return {"le": v <= other, "lt": v < other, "ge": v >= other, "gt": v > other, "eq": v == other,
                "ne": v != other}[op]

    def __le__(self, o): return self._cmp(o, "le")
    def __lt__(self, o): return self._cmp(o, "lt")
    def __ge__(self, o): return self._cmp(o, "ge")
    def __gt__(self, o): return self._cmp(o, "gt")

    def __eq__(self, o):
        if _is_number(o) and float(o) == 0.0 and self._value is None and _speculate_nonzero(self.tree):
            return False     # safe_div's ``denom == 0`` (backend/__init__.py:14-17): verified at the next real sync
        return self._cmp(o, "eq")

    def __ne__(self, o):
        r = self.__eq__(o)
        return r if r is NotImplemented else not r

    __hash__ = None

    def __repr__(self):
        return "LazyScalar(" + (repr(self._value) if self._value is not None else self.tree[0]) + ")"


def _eval_scalar(t):
    if isinstance(t, LazyScalar):
        return t.value()
    if not isinstance(t, tuple):
        return float(t)
    op = t[0]
    if op == "sum":
        return _sync(t[1].materialize().sum(dtype=torch.float64))
    if op == "max":
        return _sync(t[1].materialize().max())
    if op == "min":
        return _sync(t[1].materialize().min())
    if op == "mean":
        return _sync(t[1].materialize().mean(dtype=torch.float64))
    if op == "dot":
        return _sync((t[1].materialize().to(torch.float64) * t[2].materialize().to(torch.float64)).sum())
    if op == "neg":
        return -_eval_scalar(t[1])
    if op == "abs":
        return abs(_eval_scalar(t[1]))
    a, b = _eval_scalar(t[1]), _eval_scalar(t[2])
    if op == "add": return a + b
    if op == "sub": return a - b
    if op == "mul": return a * b
    # numpy's scalar semantics, like the reference's numpy backend: x/0 is inf or nan (RMabs against an all-zero
    # previous iterate, measures/supervised.py:114), never ZeroDivisionError
    with np.errstate(all="ignore"):
        if op == "div": return float(np.float64(a) / np.float64(b))
        if op == "pow": return float(np.float64(a) ** np.float64(b))
    raise Exception("pygrank_b200.lazy: unknown scalar node " + str(op))


def _sync(t: torch.Tensor) -> float:
    STATS["syncs"] += 1
    STATS["eager_ops"] += 1
    return float(t)


# ------------------------------------------------------------------------------------------------------------------
# vectors
# ------------------------------------------------------------------------------------------------------------------
class LazyVec:
    """A backend vector: either a materialised 1-D CUDA tensor (op "tensor") or a recorded operation.

    ops: tensor | conv(x, M) | scale(x, float) | divs(x, LazyScalar) | muls(x, LazyScalar) | add | sub | mulv | divv |
    abs | neg | rsubs(float, x) | rdivs(float, x) | pows(x, float) | iter(run, k) | res(run, k)"""
    __slots__ = ("op", "args", "_val", "n", "dtype", "_resolved", "__weakref__")

    def __init__(self, op, args, n, dtype, val=None):
        self.op, self.args, self.n, self.dtype = op, args, int(n), dtype
        self._val = val
        self._resolved = None

    @staticmethod
    def wrap(t: torch.Tensor) -> "LazyVec":
        return LazyVec("tensor", (), t.shape[0], t.dtype, t)

    # -- shape protocol ---------------------------------------------------------------------------------------------
    @property
    def shape(self):
        return (self.n,)

    @property
    def ndim(self):
        return 1

    def dim(self):
        return 1

    def __len__(self):
        return self.n

    def numel(self):
        return self.n

    @property
    def device(self):
        return self.materialize().device

    @property
    def is_cuda(self):
        return True

    # -- evaluation -------------------------------------------------------------------------------------------------
    def materialize(self) -> torch.Tensor:
        if self._val is None:
            _materialize(self)
        if self.op != "tensor":
            self.op, self.args = "tensor", ()         # drop the expression: the value is the node now
        return self._val

    @property
    def tensor(self) -> torch.Tensor:
        return self.materialize()

    def __array__(self, dtype=None, copy=None):
        a = self.materialize().detach().cpu().numpy()
        return a.astype(dtype) if dtype is not None else a

    def __getattr__(self, name):          # .cpu(), .to(), .tolist(), .sum() ... act on the value
        # only names a tensor really has: the reference probes its arguments with hasattr(arg, "array")
        # (backend/__init__.py:69-73), which must not force anything
        if name.startswith("__") or not hasattr(torch.Tensor, name):
            raise AttributeError(name)
        return getattr(self.materialize(), name)

    def __getitem__(self, key):
        if isinstance(key, LazyVec):
            key = key.materialize()
        out = self.materialize()[key]
        return LazyVec.wrap(out) if isinstance(out, torch.Tensor) and out.dim() == 1 else out

    def __setitem__(self, key, value):
        if isinstance(key, LazyVec):
            key = key.materialize()
        if isinstance(value, LazyVec):
            value = value.materialize()
        self.materialize()[key] = value

    def __iter__(self):
        return iter(self.materialize())

    def __float__(self):
        return float(self.materialize())

    # -- arithmetic: recorded, not executed -----------------------------------------------------------------------
    def _new(self, op, *args):
        return LazyVec(op, args, self.n, self.dtype)

    @staticmethod
    def _operand(o):
        if isinstance(o, LazyVec):
            return "vec", o
        if isinstance(o, LazyScalar):
            return "lazy", o
        if _is_number(o):
            return "num", float(o)
        if isinstance(o, torch.Tensor):
            if o.dim() == 0:
                return "num", float(o)
            return "vec", LazyVec.wrap(o.reshape(-1))
        if isinstance(o, (np.ndarray, list, tuple)):
            from . import backend
            return "vec", backend.to_array(o)
        return "other", o

    def __mul__(self, o):
        kind, v = self._operand(o)
        if kind == "num":
            return self._new("scale", self, v)
        if kind == "lazy":
            return self._new("muls", self, v)
        if kind == "vec":
            return self._new("mulv", self, v)
        return NotImplemented

    __rmul__ = __mul__

    def __truediv__(self, o):
        kind, v = self._operand(o)
        if kind == "num":
            return self._new("scale", self, 1.0 / v) if v != 0 else self._new("divn", self, v)
        if kind == "lazy":
            return self._new("divs", self, v)
        if kind == "vec":
            return self._new("divv", self, v)
        return NotImplemented

    def __rtruediv__(self, o):
        kind, v = self._operand(o)
        if kind == "num":
            return self._new("rdivs", v, self)
        if kind == "vec":
            return v._new("divv", v, self)
        if kind == "lazy":
            return self._new("rdivs", v, self)
        return NotImplemented

    def __add__(self, o):
        kind, v = self._operand(o)
        if kind == "vec":
            return self._new("add", self, v)
        if kind in ("num", "lazy"):
            return self._new("adds", self, v)
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, o):
        kind, v = self._operand(o)
        if kind == "vec":
            return self._new("sub", self, v)
        if kind in ("num", "lazy"):
            return self._new("adds", self, -v)
        return NotImplemented

    def __rsub__(self, o):
        kind, v = self._operand(o)
        if kind == "vec":
            return v._new("sub", v, self)
        if kind in ("num", "lazy"):
            return self._new("rsubs", v, self)
        return NotImplemented

    def __neg__(self):
        return self._new("scale", self, -1.0)

    def __pos__(self):
        return self

    def __abs__(self):
        return self._new("abs", self)

    def __pow__(self, o):
        kind, v = self._operand(o)
        if kind == "num":
            return self._new("pows", self, v)
        if kind == "vec":
            return self._new("powv", self, v)
        if kind == "lazy":
            return self._new("pows", self, v.value())
        return NotImplemented

    def __rpow__(self, o):
        kind, v = self._operand(o)
        if kind == "num":
            return self._new("rpows", v, self)
        return NotImplemented

    def __matmul__(self, o):                      # dense products (krylov_space.py) act on the value
        other = o.materialize() if isinstance(o, LazyVec) else o
        return self.materialize() @ other

    def __rmatmul__(self, o):
        other = o.materialize() if isinstance(o, LazyVec) else o
        return other @ self.materialize()

    # comparisons give masks (tensors): they are host-side decisions or index sets in the reference
    def _cmpv(self, o, fn):
        if isinstance(o, LazyVec):
            o = o.materialize()
        elif isinstance(o, LazyScalar):
            o = o.value()
        STATS["eager_ops"] += 1
        return fn(self.materialize(), o)

    def __eq__(self, o): return self._cmpv(o, torch.Tensor.__eq__)
    def __ne__(self, o): return self._cmpv(o, torch.Tensor.__ne__)
    def __lt__(self, o): return self._cmpv(o, torch.Tensor.__lt__)
    def __le__(self, o): return self._cmpv(o, torch.Tensor.__le__)
    def __gt__(self, o): return self._cmpv(o, torch.Tensor.__gt__)
    def __ge__(self, o): return self._cmpv(o, torch.Tensor.__ge__)

    __hash__ = object.__hash__

    def __repr__(self):
        return f"LazyVec({self.op}, n={self.n}" + (", materialised" if self._val is not None else "") + ")"


def _scale_of(v: LazyVec):
    """(inner, factor) when v is ``inner * factor`` with a known factor (the final ``ranks * personalization_norm`` of
    abstract_filters.py:63-64), else None."""
    if v.op == "scale":
        return v.args[0], float(v.args[1])
    if v.op in ("muls", "divs") and isinstance(v.args[1], LazyScalar) and v.args[1]._value is not None:
        f = v.args[1]._value
        return v.args[0], (f if v.op == "muls" else 1.0 / f)
    return None


def _materialize(v: LazyVec) -> torch.Tensor:
    """Value of a node.  Fused shapes first; otherwise eager, one torch call per node (an explicit stack: driver loops
    build chains as deep as max_iters)."""
    stack = [v]
    while stack:
        node = stack[-1]
        if node._val is not None:
            stack.pop()
            continue
        it = _resolve_iter(node)
        if it is not None:
            node._val = it.args[0].materialize_handle(it, 1.0)
            stack.pop()
            continue
        sc = _scale_of(node)
        if sc is not None and isinstance(sc[0], LazyVec) and sc[0]._val is None:
            it = _resolve_iter(sc[0])
            if it is not None:                      # rescaling folded into the read-out kernel
                node._val = it.args[0].materialize_handle(it, sc[1])
                stack.pop()
                continue
        pending = [a for a in node.args if isinstance(a, LazyVec) and a._val is None]
        if pending:
            if len(pending) > 1:      # iterates of one run are read in step order (prev before cur): no recomputation
                step = lambda a: (lambda h: h.args[1] if h is not None else -1)(_resolve_iter(a))
                pending.sort(key=step, reverse=True)
            stack.extend(pending)
            continue
        node._val = _eager(node)
        stack.pop()
    return v._val


def _eager(node: LazyVec) -> torch.Tensor:
    op, a = node.op, node.args
    val = lambda x: x._val if isinstance(x, LazyVec) else (x.value() if isinstance(x, LazyScalar) else x)
    if op == "conv":
        STATS["eager_convs"] += 1
        return a[1].conv(val(a[0]))
    STATS["eager_ops"] += 1
    if op == "scale": return val(a[0]) * a[1]
    if op == "divn": return val(a[0]) / a[1]
    if op == "muls": return val(a[0]) * val(a[1])
    if op == "divs": return val(a[0]) / val(a[1])
    if op == "add": return val(a[0]) + val(a[1])
    if op == "adds": return val(a[0]) + val(a[1])
    if op == "sub": return val(a[0]) - val(a[1])
    if op == "rsubs": return val(a[0]) - val(a[1])
    if op == "rdivs": return val(a[0]) / val(a[1])
    if op == "mulv": return val(a[0]) * val(a[1])
    if op == "divv": return val(a[0]) / val(a[1])
    if op == "abs": return torch.abs(val(a[0]))
    if op == "pows": return val(a[0]) ** a[1]
    if op == "powv": return val(a[0]) ** val(a[1])
    if op == "rpows": return a[0] ** val(a[1])
    if op in ("iter", "res"):
        return a[0].materialize_handle(node, 1.0)
    raise Exception("pygrank_b200.lazy: unknown vector node " + op)


# ------------------------------------------------------------------------------------------------------------------
# loop-invariant ("pure") expressions: elementwise trees over materialised tensors, evaluated once and memoised by
# structure, so that e.g. ``personalization * (1 - alpha)`` — rebuilt by the driver every iteration — costs nothing
# ------------------------------------------------------------------------------------------------------------------
_PURE_OPS = ("scale", "add", "sub", "mulv", "divv", "abs", "adds", "rsubs", "rdivs", "pows", "divn")
_pure_cache: dict = {}


def _pure_key(v, depth=0):
    """Structural key of a conv-free expression over tensors (None when it is not one, or deeper than any
    loop-invariant term of the reference's formulas)."""
    if not isinstance(v, LazyVec):
        return ("num", float(v)) if _is_number(v) else None
    if v.op == "tensor":
        return ("t", id(v._val))
    if v.op not in _PURE_OPS or depth > 24:
        return None
    parts = []
    for a in v.args:
        k = _pure_key(a, depth + 1)
        if k is None:
            return None
        parts.append(k)
    return (v.op,) + tuple(parts)


def _pure_value(v: LazyVec, key) -> torch.Tensor:
    """Value of a conv-free expression, memoised by structure + identity of its leaf tensors.  A hit is only valid while
    no leaf has been written in place since (``signal[v] = x`` mutates the backend vector of a GraphSignal,
    core/signals.py:92-93): torch's per-tensor version counters tell."""
    hit = _pure_cache.get(key)
    if hit is not None and all(t._version == ver for t, ver in zip(hit[1], hit[2])):
        return hit[0]
    leaves = []

    def collect(x):
        if isinstance(x, LazyVec):
            if x.op == "tensor":
                leaves.append(x._val)
            else:
                for a in x.args:
                    collect(a)
    collect(v)
    out = _materialize_copy(v)
    if len(_pure_cache) > 64:
        _pure_cache.pop(next(iter(_pure_cache)))
    # the leaves stay referenced so their ids cannot be recycled
    _pure_cache[key] = (out, leaves, [t._version for t in leaves])
    return out


def _materialize_copy(v: LazyVec) -> torch.Tensor:
    """Evaluate without turning the node into a tensor node (its structure is still being matched)."""
    if v._val is not None:
        return v._val
    tmp = LazyVec(v.op, v.args, v.n, v.dtype)
    return tmp.materialize()


# ------------------------------------------------------------------------------------------------------------------
# shape matching
# ------------------------------------------------------------------------------------------------------------------
class _Affine:
    """v = (a_s * a_v) o conv(x, M) + b  with a_v / b conv-free expressions (or None)."""
    __slots__ = ("a_s", "a_v", "x", "M", "b")

    def __init__(self, a_s, a_v, x, M, b):
        self.a_s, self.a_v, self.x, self.M, self.b = a_s, a_v, x, M, b


def _mul_pure(p, q, n, dtype, op="mulv"):
    if p is None:
        return q if op == "mulv" else LazyVec("rdivs", (1.0, q), n, dtype)
    if q is None:
        return p
    return LazyVec(op, (p, q), n, dtype)


def _as_affine(v, depth=0) -> Optional[_Affine]:
    if not isinstance(v, LazyVec) or v._val is not None or depth > 24:
        return None
    op, a = v.op, v.args
    if op == "conv":
        M = a[1]
        if getattr(M, "normalization", None) == "laplacian" or getattr(M, "pathological", False):
            return None
        return _Affine(1.0, None, a[0], M, None)
    if op == "scale":
        f = _as_affine(a[0], depth + 1)
        if f is None:
            return None
        b = None if f.b is None else LazyVec("scale", (f.b, a[1]), v.n, v.dtype)
        return _Affine(f.a_s * a[1], f.a_v, f.x, f.M, b)
    if op in ("mulv", "divv"):
        left = _as_affine(a[0], depth + 1)
        other = a[1]
        if left is None and op == "mulv":
            left, other = _as_affine(a[1], depth + 1), a[0]
        if left is None or _pure_key(other) is None:
            return None
        a_v = _mul_pure(left.a_v, other, v.n, v.dtype, op)
        b = None if left.b is None else LazyVec(op, (left.b, other), v.n, v.dtype)
        return _Affine(left.a_s, a_v, left.x, left.M, b)
    if op in ("add", "sub"):
        fa, fb = _as_affine(a[0], depth + 1), _as_affine(a[1], depth + 1)
        if fa is not None and fb is None and _pure_key(a[1]) is not None:
            other = a[1] if op == "add" else LazyVec("scale", (a[1], -1.0), v.n, v.dtype)
            b = other if fa.b is None else LazyVec("add", (fa.b, other), v.n, v.dtype)
            return _Affine(fa.a_s, fa.a_v, fa.x, fa.M, b)
        if fb is not None and fa is None and _pure_key(a[0]) is not None:
            sign = 1.0 if op == "add" else -1.0
            nb = fb.b if sign > 0 or fb.b is None else LazyVec("scale", (fb.b, -1.0), v.n, v.dtype)
            b = a[0] if nb is None else LazyVec("add", (a[0], nb), v.n, v.dtype)
            return _Affine(fb.a_s * sign, fb.a_v, fb.x, fb.M, b)
    return None


def _resolve_iter(v) -> Optional[LazyVec]:
    """The iterate handle (op "iter" / "res") a node stands for, creating or extending a fused run when the node is
    one more step of a recognised recursion.  Memoised on the node; iterative over chains of steps the driver never
    forced (error_type="iters" builds max_iters of them before anything is read)."""
    if not isinstance(v, LazyVec):
        return None
    if v.op in ("iter", "res"):
        return v
    if v._resolved is not None:
        return v._resolved
    if not FUSE:
        return None
    chain, node = [], v
    while True:
        if node.op in ("iter", "res"):
            base = node
            break
        if node._resolved is not None:
            base = node._resolved
            break
        if node._val is not None:
            base = None
            break
        step = _step_of(node)
        if step is None:
            base = None
            break
        chain.append((node, step))
        node = step[1]
        if not isinstance(node, LazyVec):
            return None
    operand = node                       # innermost operand: a handle, or a plain vector a new run can start from
    for holder, (kind, _, spec) in reversed(chain):
        if kind == "affine":
            nxt = AffineRun.extend(base, operand, spec[1], spec[0])
        elif kind == "poly-res":
            nxt = PolyRun.extend_result(base, operand, spec[0], spec[1])
        else:
            nxt = PolyRun.extend_power(base, operand, spec)
        if nxt is None:
            return None
        holder._resolved = nxt
        base, operand = nxt, holder
    return base


def _step_of(node: LazyVec):
    """(kind, operand, spec): ("affine", x, (quotient, form)) | ("poly-res", prev_result, (power_node, coef)) |
    ("poly-pow", prev_power, M) when the node is one step applied to ``operand``; None otherwise."""
    from .graph import DeviceGraph
    if node.op == "divs":                                   # numerator / sum(numerator): the quotient
        s = node.args[1]
        if isinstance(s, LazyScalar) and isinstance(s.tree, tuple) and s.tree[0] == "sum" and s.tree[1] is node.args[0]:
            f = _as_affine(node.args[0])
            if f is not None and isinstance(f.M, DeviceGraph):
                return "affine", f.x, (True, f)
        return None
    if node.op == "add":                                    # res + pow * c  (abstract_filters.py:226-228)
        r, t = node.args
        if isinstance(t, LazyVec) and t.op == "scale" and isinstance(t.args[0], LazyVec) and isinstance(r, LazyVec):
            pw = t.args[0]
            if _power_chain(pw) is not None:
                return "poly-res", r, (pw, float(t.args[1]))
    if node.op == "conv":
        if isinstance(node.args[1], DeviceGraph) and _power_chain(node) is not None and _power_chain(node)[0] == "run":
            return "poly-pow", node.args[0], node.args[1]
        return None
    f = _as_affine(node)
    if f is not None and f.b is not None and isinstance(f.M, DeviceGraph):
        return "affine", f.x, (False, f)
    return None


def _power_chain(node):
    """("run", handle) when node is conv^j of a power handle of a PolyRun (j >= 0), ("seed", seed_node, M, j) when it is
    conv^j(seed, M) of a materialised seed with j >= 1; None otherwise."""
    from .graph import DeviceGraph
    j, M = 0, None
    while isinstance(node, LazyVec):
        h = node if node.op == "iter" else node._resolved
        if h is not None and h.op == "iter" and isinstance(h.args[0], PolyRun) and (M is None or h.args[0].g is M):
            return "run", _handle(h.args[0], h.args[1] + j) if j else h
        if node.op == "conv" and node._val is None and isinstance(node.args[1], DeviceGraph) and \
                (M is None or node.args[1] is M):
            M = node.args[1]
            if M.normalization == "laplacian" or M.pathological:
                return None
            j += 1
            node = node.args[0]
            continue
        if node._val is not None and j >= 1:
            return "seed", node, M, j
        return None
    return None


def _speculate_nonzero(tree) -> bool:
    """``sum(numerator) == 0`` of RecursiveGraphFilter._step's safe_div (backend/__init__.py:14-17): when the numerator
    is one more step of an affine recursion the answer is taken to be False without computing anything; a zero (or
    non-finite) normaliser surfaces as a non-finite error at the run's next synchronisation and raises there."""
    if not (isinstance(tree, tuple) and tree[0] == "sum" and isinstance(tree[1], LazyVec)):
        return False
    from .graph import DeviceGraph
    f = _as_affine(tree[1])
    return f is not None and isinstance(f.M, DeviceGraph) and isinstance(f.x, LazyVec)


class _ErrorMatch:
    def __init__(self, mode, prev, cur, divisor):
        self.mode, self.prev, self.cur, self.divisor = mode, prev, cur, divisor

    def resolve(self, hint_threshold=None) -> Optional[float]:
        if self.prev is self.cur:
            return 0.0
        cur = _resolve_iter(self.cur)
        if cur is None:
            return None
        run, k = cur.args
        if isinstance(run, PolyRun) and cur.op != "res":
            return None
        prev = _resolve_iter(self.prev) if self.prev._val is None else None
        ok_prev = (prev is not None and prev.args[0] is run and prev.op == cur.op and prev.args[1] == k - 1) \
            or run.is_start(self.prev, k - 1)
        if not ok_prev:
            return None
        return run.error(k, self.mode, self.divisor, hint_threshold)


def _match_error(tree) -> Optional[_ErrorMatch]:
    """Mabs / L1 / MSQ / MaxDifference between two vectors (measures/supervised.py:93-130)."""
    div, t = 1.0, tree
    if isinstance(t, tuple) and t[0] == "div" and _is_number(t[2]):
        div, t = float(t[2]), t[1]
    if not isinstance(t, tuple) or len(t) != 2 or not isinstance(t[1], LazyVec):
        return None
    v = t[1]
    if t[0] == "sum" and v.op == "abs" and isinstance(v.args[0], LazyVec) and v.args[0].op == "sub":
        p, q = v.args[0].args
        return _ErrorMatch(C.ERR_MABS if div != 1.0 else C.ERR_L1, p, q, div)
    if t[0] == "max" and v.op == "abs" and isinstance(v.args[0], LazyVec) and v.args[0].op == "sub" and div == 1.0:
        p, q = v.args[0].args
        return _ErrorMatch(C.ERR_MAX, p, q, 1.0)
    if t[0] == "sum" and v.op == "mulv" and all(isinstance(x, LazyVec) and x.op == "sub" for x in v.args):
        (p, q), (p2, q2) = v.args[0].args, v.args[1].args
        if p is p2 and q is q2:
            return _ErrorMatch(C.ERR_MSQ, p, q, div)
    return None


# ------------------------------------------------------------------------------------------------------------------
# fused runs
# ------------------------------------------------------------------------------------------------------------------
def _handle(run, k, op="iter") -> LazyVec:
    return LazyVec(op, (run, k), run.n, run.dtype)


_KIND = {C.ERR_MABS: 0, C.ERR_L1: 0, C.ERR_MSQ: 1, C.ERR_MAX: 2}


def _check_finite(host_hist):
    if not np.all(np.isfinite(host_hist)):
        raise Exception("pygrank_b200: a fused iteration produced a non-finite error (zero normaliser in "
                        "RecursiveGraphFilter's quotient, or non-finite input)")


_SNAP_ROWS = 1024
_snap_pool: list = []      # pinned (state_f64 rows, state_i32 rows) snapshot buffers, reused across runs


def _take_snapshots():
    if _snap_pool:
        return _snap_pool.pop()
    return (torch.empty((_SNAP_ROWS, C.STATE_LEN), dtype=torch.float64).pin_memory(),
            torch.empty((_SNAP_ROWS, C.STATE_LEN), dtype=torch.int32).pin_memory())


class AffineRun:
    """x_k = (A o conv(x_{k-1}, M) + B) [/ sum] on pgb_affine_steps: resumable, and — once the driver's first convergence
    test has shown its measure and threshold — ahead of the driver with the stop decision on the device.

    The driver's test of iteration k only needs step k: after every step the device state (error of the step, stop
    flag) is copied to pinned host memory behind an event, and a test waits for THAT event, not for the steps enqueued
    after it — the driver's own Python work of iteration k+1 overlaps the device's step k+1."""

    def __init__(self, g, f: _Affine, x0: torch.Tensor, quotient: bool, x0_node):
        from .graph import dtype_code
        lib = C.lib()
        self.g, self.quotient = g, bool(quotient)
        self.n, self.dtype = g.n, x0.dtype
        self.x0_node = x0_node
        self.sig = self.signature(f, quotient)
        dtype, code = self.dtype, dtype_code(self.dtype)
        dev = x0.device
        f64 = torch.float64
        self.code = code
        a_v = None if f.a_v is None else _pure_value(f.a_v, _pure_key(f.a_v))
        b = _pure_value(f.b, _pure_key(f.b)).to(dtype).contiguous() if f.b is not None else \
            torch.zeros(g.n, dtype=dtype, device=dev)
        self.alpha = float(f.a_s)
        self.sq = g.vec("sq", dtype)
        if a_v is None:
            self.w_run = None
            self.c = g.vec("c", dtype)
        else:   # coefficient vector in internal order: w' = w o a; next normaliser through rowsum(M diag(a))
            a_int = a_v.to(f64) if g.perm is None else a_v.to(f64)[g.perm.long()]
            self.w_run = (g.vec("w", f64) * a_int).to(dtype)
            gsum = g._spmv_raw(g.out_view, g.R * a_int, g.L, f64)
            self.c = (g.vec("sq", f64) * gsum).to(dtype)
        sf = [0.0] * C.STATE_LEN
        si = [0] * C.STATE_LEN
        sf[C.SF_ALPHA], sf[C.SF_INVS], sf[C.SF_MEAN], sf[C.SF_NORM] = self.alpha, 1.0, float(g.n), 1.0
        si[C.SI_MAX_ITERS], si[C.SI_END_MODULO], si[C.SI_ERR_MODE] = 2 ** 30, 1, C.ERR_ITERS
        si[C.SI_QUOTIENT] = int(quotient)
        self.state_f64 = torch.tensor(sf, dtype=f64, device=dev)
        self.state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
        self.err_hist = torch.zeros(256, dtype=f64, device=dev)
        self.zbuf = [torch.empty(g.n, dtype=dtype, device=dev), torch.empty(g.n, dtype=dtype, device=dev)]
        self.q = torch.empty(g.n, dtype=dtype, device=dev)
        self.b, self.x0 = b, x0.to(dtype).contiguous()
        self.view = g.in_view
        self.cs = self.view.cstruct(dtype)
        self.ws = self.view.new_span_ws(dtype)
        self.symdeg = g.symdeg and self.w_run is None
        self.w_arg = None if self.symdeg else (self.w_run if self.w_run is not None else g.vec("w", dtype))
        self.sq_arg = None if self.symdeg else self.sq
        self.mean = float(g.n)        # divisor the device applies to the error sums (SF_MEAN)
        self.mode = None              # measure the device accumulates: None = |delta| sums without a stop rule
        self.threshold = None         # set: the device stops by itself and the run goes ahead of the driver
        self.chunk = 8
        self._init_device()
        STATS["runs"] += 1

    def _init_device(self):
        lib, st = C.lib(), C.stream_ptr()
        C.check(lib.pgb_affine_init(self.n, self.code, C.ptr(self.b), C.ptr(self.x0), C.ptr(self.sq), C.ptr(self.c), 1.0,
                                    None, C.ptr(self.g.perm), 0, C.ptr(self.zbuf[0]), C.ptr(self.q),
                                    C.ptr(self.state_f64), st))
        C.check(lib.pgb_affine_init_finish(C.ptr(self.state_f64), C.ptr(self.state_i32), st))
        C.count_launches(2)
        self.done = 0                 # steps enqueued on the device
        self.known = 0                # steps whose error is on the host
        self.host_err = np.zeros(1)
        self.stopped_at = None        # step at which the device-side rule stopped the run
        self.events = {}              # step -> event recorded after its state snapshot

    # -- matching ---------------------------------------------------------------------------------------------------
    @staticmethod
    def signature(f: _Affine, quotient):
        return (id(f.M), float(f.a_s), None if f.a_v is None else _pure_key(f.a_v),
                None if f.b is None else _pure_key(f.b), bool(quotient))

    @staticmethod
    def extend(base, operand, f: _Affine, quotient):
        if base is not None:
            if base.op == "iter" and isinstance(base.args[0], AffineRun) and \
                    base.args[0].sig == AffineRun.signature(f, quotient):
                return _handle(base.args[0], base.args[1] + 1)
            return None
        x0 = operand.materialize()                     # a new recursion starts from a plain vector
        if x0.dim() != 1 or x0.shape[0] != f.M.n or x0.dtype not in (torch.float32, torch.float64):
            return None
        return _handle(AffineRun(f.M, f, x0, quotient, operand), 1)

    def is_start(self, node, k) -> bool:
        return k == 0 and isinstance(node, LazyVec) and (node is self.x0_node or (
            node._val is not None and node._val is self.x0_node._val))

    # -- device ---------------------------------------------------------------------------------------------------
    def _launch(self, count):
        from .graph import span_struct
        lib = C.lib()
        if self.done + count + 2 > self.err_hist.numel():
            grown = torch.zeros(max(2 * self.err_hist.numel(), self.done + count + 2), dtype=torch.float64,
                                device=self.err_hist.device)
            grown[:self.err_hist.numel()] = self.err_hist
            self.err_hist = grown
        if getattr(self, "snap", None) is None:
            self.snap = _take_snapshots()
        st = C.stream_ptr()
        ws = span_struct(self.ws)
        for _ in range(count):
            C.check(lib.pgb_affine_steps(ctypes.byref(self.cs), self.code, self.alpha, C.ptr(self.w_arg),
                                         C.ptr(self.sq_arg), C.ptr(self.c), C.ptr(self.q), C.ptr(self.zbuf[0]),
                                         C.ptr(self.zbuf[1]), 0, C.ptr(self.state_f64), C.ptr(self.state_i32),
                                         C.ptr(self.err_hist), ws, self.done + 1, 1, 1, st))
            self.done += 1
            if self.done < _SNAP_ROWS:                       # state after this step -> pinned host row, behind an event
                self.snap[0][self.done].copy_(self.state_f64, non_blocking=True)
                self.snap[1][self.done].copy_(self.state_i32, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                self.events[self.done] = ev
        C.count_launches(count * self.view.kernels_per_step(self.dtype))
        STATS["fused_steps"] += count

    def __del__(self):
        snap = getattr(self, "snap", None)
        if snap is not None and len(_snap_pool) < 8:
            try:
                for ev in getattr(self, "events", {}).values():   # nothing may still be writing into the rows
                    ev.synchronize()
                _snap_pool.append(snap)
            except Exception:
                pass

    def _advance(self, k) -> bool:
        """Learn the outcome of steps known+1 .. k from their snapshots, waiting only for the events of those steps.
        False when a step has no snapshot (very long runs): the caller falls back to a full read."""
        while self.known < k and self.known < self.done:
            j = self.known + 1
            ev = self.events.get(j)
            if ev is None:
                return False
            ev.synchronize()
            STATS["step_waits"] += 1
            steps, stop = int(self.snap[1][j][C.SI_STEPS]), int(self.snap[1][j][C.SI_STOP])
            if steps < j:                                    # the device had already stopped: launch j was a no-op
                self.stopped_at, self.done = steps, steps
                return True
            if j >= self.host_err.shape[0]:
                grown = np.zeros(max(2 * self.host_err.shape[0], j + 8))
                grown[:self.host_err.shape[0]] = self.host_err
                self.host_err = grown
            self.host_err[j] = float(self.snap[0][j][C.SF_LASTERR])
            _check_finite(self.host_err[j:j + 1])
            self.known = j
            if stop != C.RUNNING:
                self.stopped_at, self.done = j, j            # launches enqueued after the stop were no-ops
                return True
        return True

    def _configure(self, mode, divisor, threshold):
        """Device-side measure and stop rule = the driver's test (convergence.py:97-101), adopted when its first test
        shows it (before any step has run)."""
        self.mode, self.threshold, self.mean = mode, threshold, float(divisor)
        self.state_i32[C.SI_ERR_MODE] = int(mode) if threshold is not None else (
            int(mode) if mode in (C.ERR_MSQ, C.ERR_MAX) else C.ERR_ITERS)
        if threshold is None and mode in (C.ERR_MSQ, C.ERR_MAX):
            threshold = -1.0                             # accumulate in that measure, never stop (errors are >= 0)
        self.state_f64[C.SF_MEAN] = float(divisor)
        self.state_f64[C.SF_TOL] = float(threshold) if threshold is not None else 0.0

    def _manual(self):
        """Step-by-step mode: no device-side stop (the driver's rule is not the one assumed, or it reads an iterate
        past the stop)."""
        self.state_i32[C.SI_STOP] = C.RUNNING
        if self.mode in (C.ERR_MSQ, C.ERR_MAX):
            self.state_f64[C.SF_TOL] = -1.0
        else:
            self.state_i32[C.SI_ERR_MODE] = C.ERR_ITERS
        self.stopped_at, self.threshold = None, None

    def _read(self):
        STATS["syncs"] += 1
        host_i = self.state_i32.cpu()
        steps = int(host_i[C.SI_STEPS])
        self.host_err = self.err_hist[:steps + 1].cpu().numpy()
        _check_finite(self.host_err[1:])
        self.known = steps
        if int(host_i[C.SI_STOP]) != C.RUNNING:
            self.stopped_at = steps
            self.done = steps                          # launches enqueued after the stop were no-ops
        return steps

    def error(self, k, mode, divisor, hint_threshold):
        """err(x_{k-1}, x_k) in the driver's measure, advancing the device as needed."""
        if self.done == 0 and self.mode is None:
            self._configure(mode, divisor, hint_threshold)
        elif _KIND[mode] != _KIND[self.mode if self.mode is not None else C.ERR_MABS]:
            return None                                  # another measure mid-run: the eager path answers
        while self.known < k:
            if self.stopped_at is not None:
                self._manual()
            if self.done < k or self.done == self.known:
                self._launch(max(k - self.done, self.chunk if self.threshold is not None else 1))
                if self.threshold is not None:
                    self.chunk = min(self.chunk * 2, 64)
            if not self._advance(k):
                self._read()
        scale = self.mean / float(divisor) if _KIND[mode] != 2 else 1.0
        return float(self.host_err[k]) * scale

    def materialize_handle(self, handle, scale) -> torch.Tensor:
        """x_k * scale in the user's node order (one read-out kernel)."""
        k = handle.args[1]
        if self.threshold is not None and self.stopped_at is None and self.done > self.known:
            if not self._advance(self.done):
                self._read()
        final = self.stopped_at if self.stopped_at is not None else self.done
        if final > k:
            # the device went past the iterate the driver ends on (its stop rule was not the assumed one): redo k steps
            STATS["recomputed_runs"] += 1
            self.state_f64[C.SF_TACC] = 0.0
            self.state_f64[C.SF_EACC] = 0.0
            self.state_f64[C.SF_BIAS] = 0.0
            self.state_i32[C.SI_STEPS] = 0
            self.state_i32[C.SI_TICKET] = 0
            self._manual()
            self._init_device()
            final = 0
        if final < k:
            self._manual()
            self._launch(k - final)
        out = torch.empty(self.n, dtype=self.dtype, device=self.zbuf[0].device)
        C.check(C.lib().pgb_unscale(self.n, self.code, C.ptr(self.zbuf[k & 1]), C.ptr(self.sq), None, float(scale),
                                    C.ptr(self.g.perm), C.ptr(out), C.stream_ptr()))
        C.count_launches(1)
        return out


class PolyRun:
    """res_t = res_{t-1} + c_t * pow_t ; pow_{t+1} = conv(pow_t, M) on pgb_poly_steps.  Handles: ("iter", t) is pow_t
    (pow_1 = the seed), ("res", t) the accumulated result after step t.  A run starts when the driver's result first
    adds a power that went through the operator: its seed is the materialised power below that chain, its result starts
    from whatever the driver had accumulated eagerly until then."""

    def __init__(self, g, seed_node: LazyVec, res0_node: LazyVec):
        from .graph import dtype_code
        lib, st = C.lib(), C.stream_ptr()
        seed = seed_node.materialize()
        self.g, self.n, self.dtype = g, g.n, seed.dtype
        self.code = dtype_code(self.dtype)
        self.seed_node, self.res0_node = seed_node, res0_node
        dev = seed.device
        f64 = torch.float64
        sf = [0.0] * C.STATE_LEN
        si = [0] * C.STATE_LEN
        sf[C.SF_ALPHA], sf[C.SF_INVS], sf[C.SF_MEAN], sf[C.SF_NORM] = 1.0, 1.0, float(g.n), 1.0
        si[C.SI_MAX_ITERS], si[C.SI_END_MODULO], si[C.SI_ERR_MODE] = 2 ** 30, 1, C.ERR_ITERS
        self.state_f64 = torch.tensor(sf, dtype=f64, device=dev)
        self.state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
        self.cap = 128
        self.coef_host = np.zeros(self.cap)
        self.coef_dev = torch.zeros(self.cap, dtype=f64, device=dev)
        self.err_hist = torch.zeros(self.cap + 2, dtype=f64, device=dev)
        self.sq = g.vec("sq", self.dtype)
        self.zbuf = [torch.empty(g.n, dtype=self.dtype, device=dev), torch.empty(g.n, dtype=self.dtype, device=dev)]
        res0 = res0_node.materialize().to(self.dtype)
        self.ranks = (res0 if g.perm is None else res0[g.perm.long()]).contiguous().clone()   # internal order
        C.check(lib.pgb_affine_init(g.n, self.code, C.ptr(seed.contiguous()), None, C.ptr(self.sq), None, 0.0, None,
                                    C.ptr(g.perm), 0, C.ptr(self.zbuf[0]), None, C.ptr(self.state_f64), st))
        C.check(lib.pgb_affine_init_finish(C.ptr(self.state_f64), C.ptr(self.state_i32), st))
        C.count_launches(2)
        self._seed_keep = seed
        self.view = g.in_view
        self.cs = self.view.cstruct(self.dtype)
        self.ws = self.view.new_span_ws(self.dtype)
        self.symdeg = g.symdeg
        self.w_arg = None if self.symdeg else g.vec("w", self.dtype)
        self.sq_arg = None if self.symdeg else self.sq
        self.done = 0            # steps executed: res_done accumulated, pow_{done+1} in zbuf[done & 1]
        self.coefs = {}          # t -> c_t as the driver declared them (steps without an entry only advance the power)
        self.host_err = np.zeros(1)
        self.known = 0
        self.mean = float(g.n)
        STATS["runs"] += 1

    # -- matching ---------------------------------------------------------------------------------------------------
    @staticmethod
    def extend_power(base, operand, M):
        if base is not None and base.op == "iter" and isinstance(base.args[0], PolyRun) and base.args[0].g is M:
            return _handle(base.args[0], base.args[1] + 1)
        return None

    @staticmethod
    def extend_result(base, operand, pw, coef):
        chain = _power_chain(pw)
        if chain is None:
            return None
        if base is not None:                             # res_{t} = res_{t-1} + c * pow_t of the same run
            if base.op != "res" or not isinstance(base.args[0], PolyRun) or chain[0] != "run":
                return None
            run, t = base.args
            h = chain[1]
            if h.args[0] is not run or h.args[1] != t + 1 or not run.declare(t + 1, coef):
                return None
            pw._resolved = h
            return _handle(run, t + 1, "res")
        if chain[0] != "seed":
            return None
        _, seed_node, M, j = chain                       # pw = conv^j(seed): pow_{j+1} of a new run
        seed = seed_node.materialize()
        if seed.dim() != 1 or seed.shape[0] != M.n or seed.dtype not in (torch.float32, torch.float64):
            return None
        run = PolyRun(M, seed_node, operand)
        run.declare(j + 1, coef)
        pw._resolved = _handle(run, j + 1)
        return _handle(run, j + 1, "res")

    def declare(self, t, coef) -> bool:
        if t in self.coefs:
            return self.coefs[t] == float(coef)
        if t <= self.done:
            return False
        self.coefs[t] = float(coef)
        return True

    def is_start(self, node, k) -> bool:
        first = min(self.coefs) if self.coefs else 1
        return k == first - 1 and node is self.res0_node

    # -- device ---------------------------------------------------------------------------------------------------
    def _run_to(self, t):
        """Execute steps done+1 .. t (step j adds c_j * pow_j to the result and advances the power)."""
        from .graph import span_struct
        lib = C.lib()
        if t + 2 >= self.cap:
            cap = max(2 * self.cap, t + 4)
            ch = np.zeros(cap)
            ch[:self.cap] = self.coef_host
            self.coef_host, self.cap = ch, cap
            self.coef_dev = torch.zeros(cap, dtype=torch.float64, device=self.coef_dev.device)
            eh = torch.zeros(cap + 2, dtype=torch.float64, device=self.coef_dev.device)
            eh[:self.err_hist.numel()] = self.err_hist
            self.err_hist = eh
        for j in range(self.done + 1, t + 1):
            self.coef_host[j] = self.coefs.get(j, 0.0)
        self.coef_dev.copy_(torch.from_numpy(self.coef_host))
        count = t - self.done
        C.check(lib.pgb_poly_steps(ctypes.byref(self.cs), self.code, C.ptr(self.w_arg), C.ptr(self.sq_arg),
                                   C.ptr(self.coef_dev), C.ptr(self.ranks), C.ptr(self.zbuf[0]), C.ptr(self.zbuf[1]), 0,
                                   C.ptr(self.state_f64), C.ptr(self.state_i32), C.ptr(self.err_hist),
                                   span_struct(self.ws), self.done + 1, count, 1, C.stream_ptr()))
        C.count_launches(count * self.view.kernels_per_step(self.dtype))
        STATS["fused_steps"] += count
        self.done = t

    def error(self, t, mode, divisor, hint_threshold):
        if _KIND[mode] != 0 or t not in self.coefs:
            return None
        if self.done < t:
            self._run_to(t)
        if self.known < t:
            STATS["syncs"] += 1
            self.host_err = self.err_hist[:self.done + 1].cpu().numpy()
            _check_finite(self.host_err[1:])
            self.known = self.done
        return float(self.host_err[t]) * self.mean / float(divisor)

    def materialize_handle(self, handle, scale) -> torch.Tensor:
        lib = C.lib()
        t = handle.args[1]
        out = torch.empty(self.n, dtype=self.dtype, device=self.zbuf[0].device)
        if handle.op == "res":
            if t > self.done:
                self._run_to(t)
            if t != self.done:                           # an earlier result read after the run moved on: recompute it
                STATS["recomputed_runs"] += 1
                acc, x = self.res0_node.materialize().to(self.dtype), self._seed_keep
                for j in range(1, t + 1):
                    if j in self.coefs:
                        STATS["eager_ops"] += 1
                        acc = acc + x * self.coefs[j]
                    if j < t:
                        STATS["eager_convs"] += 1
                        x = self.g.conv(x)
                return acc * scale if scale != 1.0 else acc
            C.check(lib.pgb_unscale(self.n, self.code, C.ptr(self.ranks), None, None, float(scale), C.ptr(self.g.perm),
                                    C.ptr(out), C.stream_ptr()))
            C.count_launches(1)
            return out
        if t - 1 >= self.done:                           # pow_t lives in zbuf[(t-1) & 1] once step t-1 has run
            if t - 1 > self.done:
                self._run_to(t - 1)
            C.check(lib.pgb_unscale(self.n, self.code, C.ptr(self.zbuf[(t - 1) & 1]), C.ptr(self.sq), None, float(scale),
                                    C.ptr(self.g.perm), C.ptr(out), C.stream_ptr()))
            C.count_launches(1)
            return out
        x = self._seed_keep                              # an old power (optimization_dict keeps them): recompute it
        for _ in range(t - 1):
            STATS["eager_convs"] += 1
            x = self.g.conv(x)
        return x * scale if scale != 1.0 else x
