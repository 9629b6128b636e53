#!/usr/bin/env python
"""Numpy study of the hsell unit rule under different internal node orders (no GPU).

For an RMAT graph (bench recipe) it reports, for the degree order and for the block-signature order
(nodes permuted INSIDE each hub block / inside the tail by the set of hub blocks their row touches), how
the entries split into hub units and tail entries, the padded hub slots, and a wavefront estimate of the
gather kernel (DESIGN.md §4.1 cost model: 1 per tail entry, 1 per 128-byte stream line, ~2.2 per LDS).

    python scripts/sim_relabel.py [scale] [H] [K]
"""
import sys
import time

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from pygrank_b200 import synthetic  # noqa: E402


def build(scale):
    n = 1 << scale
    src, dst = synthetic.rmat_edges_host(scale, 16, 1)
    keep = src != dst
    s = np.concatenate([src[keep], dst[keep]]).astype(np.int64)
    d = np.concatenate([dst[keep], src[keep]]).astype(np.int64)
    del src, dst
    key = np.unique(s * n + d)
    del s, d
    rows = (key // n).astype(np.int32)
    cols = (key % n).astype(np.int32)
    return n, rows, cols


def unit_stats(n, rows, cols, H, K, rule, label):
    """rows/cols: relabelled entries (any order).  rule(e, r) -> bool array: unit kept as a hub unit."""
    hub_span = K * H
    S = (n + 31) // 32
    in_hub = cols < hub_span
    nnz = len(rows)
    sl = rows[in_hub] // 32
    blk = cols[in_hub] // H
    ukey = sl.astype(np.int64) * K + blk
    # per (row, block) counts -> per unit max lane count
    rkey = rows[in_hub].astype(np.int64) * K + blk
    rk, rc = np.unique(rkey, return_counts=True)
    u_of_r = (rk // K // 32) * K + (rk % K)
    uk, inv = np.unique(u_of_r, return_inverse=True)
    ent = np.bincount(inv, weights=rc).astype(np.int64)
    mx = np.zeros(len(uk), dtype=np.int64)
    np.maximum.at(mx, inv, rc)
    rounds = (mx + 1) // 2
    use = rule(ent, rounds)
    hub_entries = int(ent[use].sum())
    hub_rounds = int(rounds[use].sum())
    sparse_tail = int(ent[~use].sum())
    col_tail = int(nnz - in_hub.sum())
    tail = sparse_tail + col_tail
    # tail rounds per slice = max lane tail length: needs per-row tail counts
    row_tail = np.bincount(rows[~in_hub], minlength=n).astype(np.int64)
    bad_units = uk[~use]
    # rows' entries in unused units
    r_unit_use = use[inv]
    np.add.at(row_tail, (rk[~r_unit_use] // K), rc[~r_unit_use])
    pad = (-n) % 32
    rt = np.concatenate([row_tail, np.zeros(pad, dtype=np.int64)]).reshape(-1, 32)
    tail_rounds = int(rt.max(1).sum())
    units = int(use.sum())
    lds = hub_rounds * 2
    wf = tail + lds * 2.2 + hub_rounds + tail_rounds + units  # + partial row stores
    print(f"[{label}] nnz={nnz/1e6:.1f}M hub={hub_entries/nnz:.3f} tail_sparse={sparse_tail/nnz:.3f} "
          f"tail_col={col_tail/nnz:.3f} | hub_rounds={hub_rounds/1e6:.2f}M slots/entry={hub_rounds*64/max(hub_entries,1):.2f} "
          f"units={units/1e6:.2f}M tail_rounds={tail_rounds/1e6:.2f}M (slots/entry {tail_rounds*32/max(tail,1):.2f}) "
          f"| est wavefronts={wf/1e6:.1f}M  (tail {tail/1e6:.1f} lds {lds*2.2/1e6:.1f} lines {(hub_rounds+tail_rounds)/1e6:.1f})")
    return wf


def signature_order(n, rows, cols, H, K, deg_rank_of):
    """new internal id for every degree-ranked node: inside each hub block (and inside the tail) nodes are
    sorted by the presence vector of hub blocks in their row (block 0 most significant, present first)."""
    hub_span = K * H
    in_hub = cols < hub_span
    blk = (cols[in_hub] // H).astype(np.int64)
    r = rows[in_hub].astype(np.int64)
    words = (K + 63) // 64
    sig = np.zeros((words, n), dtype=np.uint64)
    for w in range(words):
        m = (blk // 64) == w
        bit = np.uint64(1) << (np.uint64(63) - (blk[m] % 64).astype(np.uint64))
        # OR-reduce per row: unique (row, bit) then add
        k2 = np.unique(r[m] * 64 + (blk[m] % 64))
        np.add.at(sig[w], k2 // 64, np.uint64(1) << (np.uint64(63) - (k2 % 64).astype(np.uint64)))
    region = np.minimum(np.arange(n, dtype=np.int64) // H, K)       # block of the node, K = tail
    keys = [np.arange(n)]                                            # stable: degree rank last
    for w in range(words - 1, -1, -1):
        keys.append(~sig[w])                                         # present (1) first
    keys.append(region)
    order = np.lexsort(keys)                                         # order[k] = old id at new position k
    new_id = np.empty(n, dtype=np.int64)
    new_id[order] = np.arange(n)
    return new_id


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    H = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
    n = 1 << scale
    Ks = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [max(64, min(256, n // (8 * H)))]
    t0 = time.time()
    n, rows, cols = build(scale)
    print(f"built scale {scale}: n={n} nnz={len(rows)} in {time.time()-t0:.0f}s; H={H} Ks={Ks}")
    deg = np.bincount(rows, minlength=n)
    perm = np.argsort(-deg, kind="stable")
    iperm = np.empty(n, dtype=np.int64)
    iperm[perm] = np.arange(n)
    rows = iperm[rows].astype(np.int32)
    cols = iperm[cols].astype(np.int32)
    for K in Ks:
      K = min(K, -(-n // H))
      print("---- K =", K)
      run(n, rows, cols, H, K)


def run(n, rows, cols, H, K):
    rules = {
        "min32": lambda e, r: e >= 32,
        "cost(5.4r+2)": lambda e, r: e >= 5.4 * r + 2,
    }
    for name, rule in rules.items():
        unit_stats(n, rows, cols, H, K, rule, "degree order, " + name)
    new_id = signature_order(n, rows, cols, H, K, None)
    rows2 = new_id[rows].astype(np.int32)
    cols2 = new_id[cols].astype(np.int32)
    for name, rule in rules.items():
        unit_stats(n, rows2, cols2, H, K, rule, "signature order, " + name)


if __name__ == "__main__":
    main()
