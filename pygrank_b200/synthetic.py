"""Synthetic graph generators named by BASELINE.json (RMAT, Barabási–Albert).

Two implementations of the *same* integer recipe:

* ``rmat_edges_host`` / ``ba_edges_host`` — numpy, for CPU tests and the CPU baseline;
* ``rmat_edges_device`` / ``ba_edges_device`` — the CUDA kernels ``pgb_rmat_edges`` /
  ``pgb_ba_edges`` in ``csrc/graphgen.cu`` (used by the bench and GPU tests).

Both draw every random bit from a counter-based splitmix64 hash of (seed, edge index,
level) and compare against integer thresholds, so host and device produce bit-identical
edge lists — the GPU tests assert it.  The reference has no generator of its own (its
graphs are downloaded, /root/reference/pygrank/benchmarks/download.py:62-72); synthetic
CSR enters it through ``pg.AdjacencyWrapper`` (/root/reference/pygrank/fastgraph/wrapgraph.py:4-22).
"""
from __future__ import annotations

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_EDGE_MUL = np.uint64(0xD1342543DE82EF95)


def _mix64(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser applied to ``x + GOLDEN`` (wrap-around uint64 arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def rmat_thresholds(a: float, b: float, c: float):
    t1 = int(a * 4294967296.0)
    t2 = int((a + b) * 4294967296.0)
    t3 = int((a + b + c) * 4294967296.0)
    return t1, t2, t3


def rmat_edges_host(scale: int, edge_factor: int = 16, seed: int = 1, a=0.57, b=0.19, c=0.19,
                    first_edge: int = 0, num_edges: int | None = None):
    """Directed RMAT edge list (src, dst) as int32 arrays; edges ``first_edge .. +num_edges``.

    Level l of edge e uses 32 bits of ``mix64(s_e + (l//2)*GOLDEN)`` with
    ``s_e = mix64(mix64(seed) ^ (e * EDGE_MUL))``: the high half for even l, the low half for
    odd l.  u < t1 → quadrant (0,0); < t2 → (0,1); < t3 → (1,0); else (1,1); the most
    significant vertex bit is decided first.
    """
    total = edge_factor << scale
    if num_edges is None:
        num_edges = total - first_edge
    t1, t2, t3 = rmat_thresholds(a, b, c)
    e = np.arange(first_edge, first_edge + num_edges, dtype=np.uint64)
    with np.errstate(over="ignore"):
        s = _mix64(_mix64(np.full(1, seed, dtype=np.uint64)) ^ (e * _EDGE_MUL))
        src = np.zeros(num_edges, dtype=np.int64)
        dst = np.zeros(num_edges, dtype=np.int64)
        for lvl in range(scale):
            if lvl % 2 == 0:
                h = _mix64(s + np.uint64(lvl // 2) * GOLDEN)
                u = (h >> np.uint64(32)).astype(np.int64)
            else:
                u = (h & np.uint64(0xFFFFFFFF)).astype(np.int64)
            sbit = (u >= t2)
            dbit = ((u >= t1) & (u < t2)) | (u >= t3)
            src = (src << 1) | sbit
            dst = (dst << 1) | dbit
    return src.astype(np.int32), dst.astype(np.int32)


def ba_edges_host(n: int, m: int, seed: int = 1):
    """Barabási–Albert-like preferential attachment (Batagelj–Brandes endpoint-copy recipe).

    Node v >= m adds m edges; edge slot k = (v-m)*m + i has source v and a target copied
    from a uniformly random earlier *endpoint* (position r = hash(k) mod 2*(v-m)*m of the
    endpoint array: even positions hold sources, odd positions hold targets); node m
    attaches to 0..m-1.  Because r is a pure function of the slot index, a reference to an
    odd position is resolved by walking to that slot and hashing again (no stored state).  Duplicate (v, t) pairs may occur
    and are collapsed by the CSR build, so the graph has slightly fewer than m*(n-m) edges.
    """
    num = (n - m) * m
    seed_h = _mix64(np.full(1, seed, dtype=np.uint64))
    cur = np.arange(num, dtype=np.int64)            # slot whose target is being resolved
    target = np.full(num, -1, dtype=np.int64)
    active = np.arange(num, dtype=np.int64)
    while len(active):
        c = cur[active]
        limit = 2 * (c // m) * m                    # endpoints owned by earlier nodes
        first = limit == 0
        with np.errstate(over="ignore"):
            h = _mix64(seed_h ^ (c.astype(np.uint64) * _EDGE_MUL))
        r = (h % np.maximum(limit, 1).astype(np.uint64)).astype(np.int64)
        even = (r % 2) == 0
        slot = r // 2
        done = first | even
        target[active[done]] = np.where(first, c % m, slot // m + m)[done]
        cur[active[~done]] = slot[~done]            # odd position: copy that slot's target
        active = active[~done]
    v = np.arange(num, dtype=np.int64) // m + m
    return v.astype(np.int32), target.astype(np.int32)


def undirected_csr_host(n: int, src: np.ndarray, dst: np.ndarray):
    """Symmetrise, drop self loops, collapse duplicates to weight 1 → scipy CSR (float64)."""
    import scipy.sparse as sp
    keep = src != dst
    s = np.concatenate([src[keep], dst[keep]]).astype(np.int64)
    d = np.concatenate([dst[keep], src[keep]]).astype(np.int64)
    key = np.unique(s * n + d)
    rows = (key // n).astype(np.int32)
    cols = (key % n).astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    return sp.csr_matrix((np.ones(len(cols), dtype=np.float64), cols, indptr), shape=(n, n))


def directed_csr_host(n: int, src: np.ndarray, dst: np.ndarray, drop_self_loops: bool = True):
    """Directed edge list → scipy CSR with duplicates collapsed to weight 1."""
    import scipy.sparse as sp
    if drop_self_loops:
        keep = src != dst
        src, dst = src[keep], dst[keep]
    key = np.unique(src.astype(np.int64) * n + dst.astype(np.int64))
    rows = (key // n).astype(np.int32)
    cols = (key % n).astype(np.int32)
    indptr = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=indptr[1:])
    return sp.csr_matrix((np.ones(len(cols), dtype=np.float64), cols, indptr), shape=(n, n))


def rmat_graph_host(scale: int, edge_factor: int = 16, seed: int = 1):
    src, dst = rmat_edges_host(scale, edge_factor, seed)
    return undirected_csr_host(1 << scale, src, dst)


def ba_graph_host(n: int, m: int, seed: int = 1):
    src, dst = ba_edges_host(n, m, seed)
    return undirected_csr_host(n, src, dst)


def seed_sets(n: int, num_sets: int, seeds_per_set: int = 10, seed: int = 0) -> np.ndarray:
    """``num_sets`` personalization seed sets of distinct uniformly random nodes."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.choice(n, size=seeds_per_set, replace=False) for _ in range(num_sets)]).astype(np.int64)
