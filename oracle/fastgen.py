"""ORACLE / BENCH INFRASTRUCTURE ONLY (never imported by ``pygrank_b200``).

Fast host construction of the bench's synthetic inputs for the CPU baseline legs: the RMAT edge recipe of
``pygrank_b200.synthetic.rmat_edges_host`` restated in C (``rmat_gen.c``, called from a few host threads; bit-identical — see
``tests/test_oracle.py``) and a symmetrise/dedupe through scipy's ``coo -> csr`` (C code, no O(nnz log nnz)
``np.unique``).  RMAT-22 takes ~15 s instead of several minutes.
"""
from __future__ import annotations

import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import scipy.sparse as sp

from . import reference_port as _orc


def rmat_edges(scale: int, edge_factor: int = 16, seed: int = 1, a=0.57, b=0.19, c=0.19):
    lib = _orc._c()
    total = edge_factor << scale
    t1, t2, t3 = int(a * 4294967296.0), int((a + b) * 4294967296.0), int((a + b + c) * 4294967296.0)
    src = np.empty(total, dtype=np.int32)
    dst = np.empty(total, dtype=np.int32)
    threads = max(1, min(os.cpu_count() or 1, 16, total >> 16))
    bounds = [total * k // threads for k in range(threads + 1)]

    def part(k):   # ctypes releases the GIL: the ranges are generated in parallel
        lo, hi = bounds[k], bounds[k + 1]
        lib.oracle_rmat_edges(scale, lo, hi - lo, seed, t1, t2, t3, _orc._ptr(src[lo:hi]), _orc._ptr(dst[lo:hi]))

    with ThreadPoolExecutor(threads) as pool:
        list(pool.map(part, range(threads)))
    return src, dst


def undirected_csr(n: int, src: np.ndarray, dst: np.ndarray) -> sp.csr_matrix:
    """Symmetrise, drop self loops, collapse duplicates to weight 1 (== synthetic.undirected_csr_host)."""
    keep = src != dst
    s, d = src[keep], dst[keep]
    rows = np.concatenate([s, d])
    cols = np.concatenate([d, s])
    del s, d
    A = sp.coo_matrix((np.ones(len(rows), dtype=np.float64), (rows, cols)), shape=(n, n)).tocsr()
    del rows, cols
    A.sum_duplicates()
    A.sort_indices()
    A.data[:] = 1.0
    return A


def rmat_graph(scale: int, edge_factor: int = 16, seed: int = 1) -> sp.csr_matrix:
    src, dst = rmat_edges(scale, edge_factor, seed)
    return undirected_csr(1 << scale, src, dst)
