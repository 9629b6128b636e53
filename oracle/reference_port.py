"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by ``pygrank_b200``).

A CPU restatement, in numpy + the C scatter in ``csc_matvec.c``, of pygrank's
iterative node-ranking hot path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it,
and only as the checker or the timed CPU baseline — the product path fails
loudly when its CUDA library is missing instead of falling back to this.

Parity status: **pinned**.  ``tests/golden/make_golden.py`` ran the real
reference (``/root/reference``, pygrank 0.2.12, numpy backend, scipy 1.18.1,
numpy 2.3.5) in the build container and committed its inputs/outputs under
``tests/golden/``; ``tests/test_oracle.py`` requires this port to reproduce the
normalised CSR, the degrees, the iteration counts *and the scores* bit for bit.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  Where the arithmetic lives in scipy (an unpinned
third-party dependency, ``setup.py:28-30``) the published algorithm is restated
explicitly (``csc_matvec`` scatter, diagonal scaling of CSR data) rather than
called, and the tests cross-check the restatement against scipy itself.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Callable, Optional, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build_c(force: bool = False) -> str:
    """Compile ``csc_matvec.c`` with the Makefile next to it (gcc only)."""
    if force or not os.path.exists(_LIB_PATH) or (
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f))
                                              for f in ("csc_matvec.c", "rmat_gen.c"))):
        subprocess.run(["make", "-s", "-C", _HERE] + (["-B"] if force else []), check=True)
    return _LIB_PATH


def _c():
    global _lib
    if _lib is None:
        lib = ctypes.CDLL(build_c())
        i64, vp = ctypes.c_int64, ctypes.c_void_p
        for name in ("oracle_rowvec_times_csr_f64", "oracle_rowvec_times_csr_f64_i64",
                     "oracle_rowvec_times_csr_f32"):
            getattr(lib, name).argtypes = [i64, vp, vp, vp, vp, vp]
            getattr(lib, name).restype = None
        lib.oracle_csr_row_sums_f64.argtypes = [i64, vp, vp, vp]
        lib.oracle_csr_row_sums_f64.restype = None
        lib.oracle_rmat_edges.argtypes = [ctypes.c_int, i64, i64, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32,
                                          ctypes.c_uint32, vp, vp]
        lib.oracle_rmat_edges.restype = None
        _lib = lib
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# --------------------------------------------------------------------------------------
# L1: backend arithmetic (pygrank/core/backend/numpy.py)
# --------------------------------------------------------------------------------------

def epsilon() -> float:
    """numpy.py:84-86 — fp64 machine epsilon."""
    return float(np.finfo(float).eps)


def conv(x: np.ndarray, M: sp.csr_matrix) -> np.ndarray:
    """``signal @ M`` (numpy.py:64-65) == scipy ``csc_matvec`` on M's arrays read as CSC.

    y[indices[k]] += data[k] * x[row(k)], rows ascending, entries in storage order.
    """
    M = M.tocsr()
    n_row, n_col = M.shape
    if x.dtype == np.float32 and M.data.dtype == np.float32:
        y = np.zeros(n_col, dtype=np.float32)
        _c().oracle_rowvec_times_csr_f32(n_row, _ptr(np.ascontiguousarray(M.indptr, dtype=np.int32)),
                                         _ptr(np.ascontiguousarray(M.indices, dtype=np.int32)),
                                         _ptr(M.data), _ptr(np.ascontiguousarray(x)), _ptr(y))
        return y
    x = np.ascontiguousarray(x, dtype=np.float64)
    data = np.ascontiguousarray(M.data, dtype=np.float64)
    y = np.zeros(n_col, dtype=np.float64)
    if M.indices.dtype == np.int64 or M.indptr.dtype == np.int64:
        _c().oracle_rowvec_times_csr_f64_i64(n_row, _ptr(np.ascontiguousarray(M.indptr, dtype=np.int64)),
                                             _ptr(np.ascontiguousarray(M.indices, dtype=np.int64)),
                                             _ptr(data), _ptr(x), _ptr(y))
    else:
        _c().oracle_rowvec_times_csr_f64(n_row, _ptr(M.indptr), _ptr(M.indices), _ptr(data), _ptr(x), _ptr(y))
    return y


def conv_python(x: Sequence[float], indptr, indices, data, n_col: int):
    """Pure-Python statement of the same scatter (tiny cases only; used to pin the C)."""
    y = [0.0] * n_col
    for j in range(len(indptr) - 1):
        for k in range(indptr[j], indptr[j + 1]):
            y[indices[k]] += data[k] * x[j]
    return y


def degrees(M) -> np.ndarray:
    """numpy.py:76-77 — ``np.asarray(sum(M, axis=1)).ravel()``.

    scipy reduces each CSR row with ``np.add.reduceat(data, indptr[:-1])`` on the rows that
    are non-empty (``_minor_reduce``); restated here with the same ufunc so pairwise
    rounding matches for float weights (unweighted graphs are exact either way).
    """
    M = M.tocsr()
    out = np.zeros(M.shape[0], dtype=np.result_type(M.data.dtype, np.float64))
    nonempty = np.flatnonzero(np.diff(M.indptr))
    if len(nonempty):
        out[nonempty] = np.add.reduceat(M.data, M.indptr[nonempty])
    return out


def col_degrees(M) -> np.ndarray:
    """``left_reduction(x.T)`` (preprocessing.py:104-105): row sums of the transpose.

    scipy transposes CSR to CSC and reduces columns through ``M.T @ ones``-style scatter in
    ascending row order; restated with the scatter above.
    """
    M = M.tocsr()
    return conv(np.ones(M.shape[0]), sp.csr_matrix((M.data.astype(np.float64), M.indices, M.indptr), shape=M.shape))


# --------------------------------------------------------------------------------------
# L2: preprocessing (pygrank/core/utils/preprocessing.py:50-152)
# --------------------------------------------------------------------------------------

def _safe_reciprocal(s: np.ndarray) -> np.ndarray:
    """``S[S != 0] = 1.0 / S[S != 0]`` (preprocessing.py:111,116,...)."""
    s = np.array(s, dtype=np.float64).flatten()
    nz = s != 0
    s[nz] = 1.0 / s[nz]
    return s


def _reverse_rows(M: sp.csr_matrix) -> sp.csr_matrix:
    """Reverse the entries inside every row: the storage order scipy's ``csr_matmat`` leaves
    after ONE product with a diagonal matrix (its per-row linked list is unwound backwards),
    i.e. the order ``pg.degrees`` sees on a "col"-normalised matrix.  Two products restore
    ascending order.  Only ``degrees`` on float-weighted graphs can tell the difference."""
    M = M.tocsr()
    nnz = M.nnz
    rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    pos = M.indptr[rows] + (M.indptr[rows + 1] - 1 - np.arange(nnz))
    return sp.csr_matrix((M.data[pos], M.indices[pos], M.indptr.copy()), shape=M.shape)


def _scale_rows_then_cols(M: sp.csr_matrix, left: Optional[np.ndarray], right: Optional[np.ndarray]) -> sp.csr_matrix:
    """``diag(left) @ M @ diag(right)`` as scipy's csr_matmat evaluates it for diagonal
    factors: data[k] = (left[i] * a_ik) * right[col(k)], left product first
    (preprocessing.py:113,121,130,138).  Stored zeros are kept; indices returned sorted
    (canonical form — scipy's own output order differs between one and two products).
    """
    M = M.tocsr().astype(np.float64)
    M.sort_indices()
    rows = np.repeat(np.arange(M.shape[0]), np.diff(M.indptr))
    data = M.data.copy()
    if left is not None:
        data = left[rows] * data
    if right is not None:
        data = data * right[M.indices]
    return sp.csr_matrix((data, M.indices.copy(), M.indptr.copy()), shape=M.shape)


def to_sparse_matrix(A, normalization="auto", directed: bool = False, renormalize=False,
                     reference_storage_order: bool = False) -> sp.csr_matrix:
    """Graph → normalised CSR, canonical (sorted-index) form unless
    ``reference_storage_order`` asks for the in-row order the reference's scipy calls leave
    ("col": reversed, see ``_reverse_rows``).

    Follows ``to_sparse_matrix`` (preprocessing.py:99-143) for an adjacency that enters
    through ``pg.AdjacencyWrapper(A, directed)`` (fastgraph/wrapgraph.py:4-22): "auto" picks
    "col" for directed graphs else "symmetric" (:101-102); ``renormalize`` adds that many
    self loops (:107-108); "col" scales rows by 1/rowsum (:109-113); "symmetric" scales by
    1/sqrt(rowsum) on the left and 1/sqrt(colsum) on the right (:131-138); "laplacian" is
    I - symmetric (:114-122); "both" uses plain reciprocals on both sides (:123-130).
    """
    M = sp.csr_matrix(A, dtype=np.float64)
    M.sum_duplicates()
    if isinstance(normalization, str):
        normalization = normalization.lower()
    if normalization == "auto":
        normalization = "col" if directed else "symmetric"
    renormalize = float(renormalize)
    if renormalize != 0:
        M = (M + sp.eye(M.shape[0], format="csr") * renormalize).tocsr()
        M.sort_indices()
    if normalization == "col":
        out = _scale_rows_then_cols(M, _safe_reciprocal(degrees(M)), None)
        return _reverse_rows(out) if reference_storage_order else out
    if normalization in ("symmetric", "laplacian"):
        left = _safe_reciprocal(np.sqrt(degrees(M)))
        right = _safe_reciprocal(np.sqrt(col_degrees(M)))
        S = _scale_rows_then_cols(M, left, right)
        if normalization == "symmetric":
            return S
        L = (-S + sp.eye(M.shape[0], format="csr")).tocsr()
        L.sort_indices()
        return L
    if normalization == "both":
        return _scale_rows_then_cols(M, _safe_reciprocal(degrees(M)), _safe_reciprocal(col_degrees(M)))
    if callable(normalization):
        out = sp.csr_matrix(normalization(M))
        out.sort_indices()
        return out
    if normalization != "none":
        raise Exception("Supported normalizations: none, col, symmetric, both, laplacian, auto")
    M.sort_indices()
    return M


def canonical(M) -> sp.csr_matrix:
    """Sorted-index copy, the form in which "bit-exact CSR" is compared."""
    M = sp.csr_matrix(M).copy()
    M.sort_indices()
    return M


# --------------------------------------------------------------------------------------
# L3: convergence (pygrank/algorithms/convergence.py:9-104, measures/supervised.py:93-145)
# --------------------------------------------------------------------------------------

def _mabs(prev, cur):      # supervised.py:101-106
    return np.sum(np.abs(prev - cur)) / len(cur)


def _l1(prev, cur):        # supervised.py:125-130
    return np.sum(np.abs(prev - cur))


def _msq(prev, cur):       # supervised.py:117-122 (MSQ)
    return np.sum((prev - cur) * (prev - cur)) / len(cur)


def _maxdiff(prev, cur):   # supervised.py:93-98
    return np.max(np.abs(prev - cur))


ERRORS = {"mabs": _mabs, "l1": _l1, "msq": _msq, "max": _maxdiff}


class Convergence:
    """ConvergenceManager (convergence.py:24-101): the counter is advanced *before* the
    check (:85), ``max_iters`` raises unless ``error_type == "iters"`` (:86-90), checks are
    skipped when ``iteration % end_modulo != 0`` (:99-100), and the test is
    ``error <= max(tol, epsilon)`` or ``<= 0`` for ``tol=None`` (:101)."""

    def __init__(self, tol: Optional[float] = 1e-6, error_type="mabs", max_iters: int = 100, end_modulo: int = 1):
        self.tol, self.error_type, self.max_iters, self.end_modulo = tol, error_type, max_iters, end_modulo
        self.iteration = 0
        self.last = None
        self.errors = []

    def start(self):
        self.iteration = 0
        self.last = None
        self.errors = []

    def has_converged(self, ranks) -> bool:
        self.iteration += 1
        if self.iteration >= self.max_iters:
            if self.error_type == "iters":
                return True
            raise Exception("Could not converge within " + str(self.max_iters) + " iterations")
        done = False
        if self.last is not None and self.error_type != "iters" and self.iteration % self.end_modulo == 0:
            fn = ERRORS[self.error_type] if isinstance(self.error_type, str) else self.error_type
            err = fn(self.last, ranks)
            self.errors.append(float(err))
            done = bool(err <= (0 if self.tol is None else max(self.tol, epsilon())))
        self.last = ranks
        return done


# --------------------------------------------------------------------------------------
# L3: graph filters (pygrank/algorithms/filters/abstract_filters.py, adhoc.py, low_pass.py)
# --------------------------------------------------------------------------------------

def _rank(M, personalization, convergence: Convergence, start: Callable, step: Callable,
          warm_start=None, preserve_norm: bool = True) -> Tuple[np.ndarray, int]:
    """``GraphFilter.rank`` (abstract_filters.py:44-65)."""
    p = np.asarray(personalization, dtype=np.float64)
    norm = np.sum(np.abs(p))                       # :52
    if norm == 0:                                  # :53-54
        return p, 0
    p = p / norm                                   # :55
    ranks = np.copy(p) if warm_start is None else np.asarray(warm_start, dtype=np.float64)   # :56
    convergence.start()                            # :58
    state = start(p, ranks)                        # :59
    ranks = state.pop("ranks", ranks)
    while not convergence.has_converged(ranks):    # :60
        ranks = step(p, ranks, state)              # :61
    if preserve_norm:                              # :63-64
        ranks = ranks * norm
    return ranks, convergence.iteration


def _quotient(ranks):
    """``safe_div(ranks, sum(ranks))`` (abstract_filters.py:133-134, backend/__init__.py:14-17)."""
    s = np.sum(ranks)
    if s == 0:
        return np.zeros_like(ranks)
    return ranks / s


def pagerank(M, personalization, alpha: float = 0.85, tol=1e-6, max_iters: int = 100, use_quotient: bool = True,
             error_type="mabs", end_modulo: int = 1, warm_start=None, preserve_norm: bool = True):
    """``PageRank`` (adhoc.py:10-45): ranks <- conv(ranks, M)*alpha + p*(1-alpha), then the
    optional L1 quotient of ``RecursiveGraphFilter._step`` (abstract_filters.py:126-136)."""
    cm = Convergence(tol, error_type, max_iters, end_modulo)

    def step(p, ranks, _):
        ranks = conv(ranks, M) * alpha + p * (1 - alpha)      # adhoc.py:36
        return _quotient(ranks) if use_quotient else ranks

    ranks, iters = _rank(M, personalization, cm, lambda p, r: {}, step, warm_start, preserve_norm)
    return ranks, iters, cm.errors


def absorbing_walks(M, personalization, alpha: float = 1 - 1e-6, absorption=None, tol=1e-6, max_iters: int = 100,
                    use_quotient: bool = True, error_type="mabs", end_modulo: int = 1, preserve_norm: bool = True):
    """``AbsorbingWalks`` (adhoc.py:124-174): ``_start`` builds absorption*(1-a)/a and
    ``degrees(M)`` (:157-159); ``_formula`` is
    (conv(r, M)*deg + p*absorption) / (absorption + deg) (:166-169)."""
    cm = Convergence(tol, error_type, max_iters, end_modulo)
    n = M.shape[0]
    absorb = (np.ones(n) if absorption is None else np.asarray(absorption, dtype=np.float64)) * ((1 - alpha) / alpha)
    deg = degrees(M)

    def step(p, ranks, _):
        ranks = (conv(ranks, M) * deg + p * absorb) / (absorb + deg)
        return _quotient(ranks) if use_quotient else ranks

    ranks, iters = _rank(M, personalization, cm, lambda p, r: {}, step, None, preserve_norm)
    return ranks, iters, cm.errors


def closed_form(M, personalization, coefficient: Callable[[Optional[float], int], float], tol=1e-6,
                max_iters: int = 100, error_type="mabs", end_modulo: int = 1, preserve_norm: bool = True):
    """``ClosedFormGraphFilter`` with taylor coefficients, node space (abstract_filters.py:196-256):
    ``_start`` sets power = p, ranks = 0 (:211-213); each ``_step`` draws the next coefficient
    (:249), adds ``power*coefficient`` unless the coefficient is 0 (:225-228), then advances
    ``power = conv(power, M)`` (:255-256, :241-246)."""
    cm = Convergence(tol, error_type, max_iters, end_modulo)

    def start(p, ranks):
        return {"power": p, "coef": None, "ranks": np.repeat(0.0, len(p))}

    def step(p, ranks, st):
        st["coef"] = coefficient(st["coef"], cm.iteration)
        if st["coef"] != 0:
            ranks = ranks + st["power"] * st["coef"]
        st["power"] = conv(st["power"], M)
        return ranks

    ranks, iters = _rank(M, personalization, cm, start, step, None, preserve_norm)
    return ranks, iters, cm.errors


def heat_kernel(M, personalization, t: float = 3, **kw):
    """``HeatKernel._coefficient`` (adhoc.py:113-116): 1, then prev*t/(iteration+1)."""
    return closed_form(M, personalization, lambda prev, it: 1. if prev is None else prev * t / (it + 1), **kw)


def pagerank_closed(M, personalization, alpha: float = 0.85, **kw):
    """``PageRankClosed._coefficient`` (adhoc.py:83-84): 1, then prev*alpha."""
    return closed_form(M, personalization, lambda prev, it: 1. if prev is None else prev * alpha, **kw)


def generic_filter(M, personalization, weights=None, **kw):
    """``GenericGraphFilter._coefficient`` (low_pass.py:23-26): weights[iteration-1], 0 past the end."""
    weights = [0.9] * 10 if weights is None else list(weights)
    return closed_form(M, personalization, lambda prev, it: 0 if it > len(weights) else weights[it - 1], **kw)


def propagate(run: Callable, M, features: np.ndarray, **kw):
    """``NodeRanking.propagate`` (signals.py:225-226): one independent ``rank`` per column."""
    cols, iters = [], []
    for c in range(features.shape[1]):
        r, it, _ = run(M, features[:, c], **kw)
        cols.append(r)
        iters.append(it)
    return np.column_stack(cols), iters
