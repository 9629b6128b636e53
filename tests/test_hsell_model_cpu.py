"""CPU model of the hub-blocked sliced-ELL path (no GPU): a numpy restatement of what
``hsell_count_kernel`` / ``hsell_fill_kernel`` write and of what the gather / reduce / update kernels
compute (csrc/hsell.cu), glued together by the REAL host code (``pygrank_b200.graph.hsell_layout`` /
``stream_layout``).  If the layout arithmetic and the documented kernel semantics fit together, the model
reproduces ``A @ z`` for any block size, unit threshold and two-level reduction setting — which is what
this test asserts; the GPU tests then check the device kernels against the same golden answers."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from pygrank_b200 import synthetic
from pygrank_b200.graph import hsell_layout

CH = 32


def count_np(indptr, indices, n, H, K, min_entries, W=1, window_of=lambda v: 0, window_min=0):
    """hsell_count_kernel: rounds of the unit of (slice, block) — 0 when the slice has fewer than
    ``min_entries`` entries there — and the tail rounds of every slice in every tail window (``window_of`` maps a
    virtual column to its window)."""
    S = (n + 31) // 32
    hub = np.zeros((K, S), dtype=np.int64)
    tail = np.zeros((W, S), dtype=np.int64)
    for s in range(S):
        lens = np.zeros((32, K + 1), dtype=np.int64)
        rows = []
        for lane in range(32):
            r = s * 32 + lane
            cols = indices[indptr[r]:indptr[r + 1]] if r < n else indices[:0]
            rows.append(cols)
            blk = np.minimum(cols // H, K)
            lens[lane] = np.bincount(blk, minlength=K + 1)
        unit = np.zeros(K + 1, dtype=bool)
        for b in range(K):
            ent, mx = lens[:, b].sum(), lens[:, b].max()
            if ent >= min_entries and ent > 0:
                hub[b, s] = (mx + 1) // 2
                unit[b] = True
        tail_len = np.zeros((32, W), dtype=np.int64)
        for lane, cols in enumerate(rows):
            for c in cols[~unit[np.minimum(cols // H, K)]]:
                tail_len[lane, window_of(c)] += 1
        if W > 1 and tail_len.sum(axis=1).max() < window_min:   # short tails stay one unit, in the last window
            total = tail_len.sum(axis=1)
            tail_len[:] = 0
            tail_len[:, W - 1] = total
        tail[:, s] = tail_len.max(axis=0)
    return (hub, tail[0]) if W == 1 else (hub, tail)


def fill_np(indptr, indices, n, H, K, hub, tail, lay, real_col=lambda v: v, W=1, window_of=lambda v: 0):
    """hsell_fill_kernel without the bank-aware order: round data [round][lane] and piece_row.  ``real_col``
    maps the (virtual) columns of the tail to positions in the gather vector (hsell_real_col); the tail of a slice
    is one unit per window."""
    S = (n + 31) // 32
    tail = np.asarray(tail).reshape(W, S)
    hub_words = np.full((lay["n_hub_chunks"] * CH * 32, 2), H, dtype=np.int64)      # (lo, hi) per word
    tail_cols = np.full(lay["n_tail_chunks"] * CH * 32, -1, dtype=np.int64)
    piece_row = np.full(max(lay["n_pieces"], 1), lay["dump_row"], dtype=np.int64)
    hub_g0 = lay["hub_g0"].numpy().reshape(K, S) if K else None
    hub_p0 = lay["hub_p0"].numpy().reshape(K, S) if K else None
    tail_g0, tail_p0 = lay["tail_g0"].numpy().reshape(W, S), lay["tail_p0"].numpy().reshape(W, S)
    slice_ptr = lay["slice_ptr"].numpy().astype(np.int64)
    for s in range(S):
        ord_ = 0
        t = np.zeros((32, W), dtype=np.int64)
        collapsed = W > 1 and not tail[:W - 1, s].any()
        win = (lambda c: W - 1) if collapsed else window_of
        for b in range(K):
            R = hub[b, s]
            for lane in range(32):
                r = s * 32 + lane
                if r >= n:
                    continue
                cols = indices[indptr[r]:indptr[r + 1]]
                mine = cols[(cols // H == b)] if b < K else cols[:0]
                if R > 0:
                    g0 = hub_g0[b, s]
                    for i, c in enumerate(mine):
                        hub_words[(g0 + i // 2) * 32 + lane, i % 2] = c - b * H
                else:
                    for c in mine:
                        w = win(c)
                        tail_cols[(tail_g0[w, s] + t[lane, w]) * 32 + lane] = real_col(c)
                        t[lane, w] += 1
            if R > 0:
                g0 = hub_g0[b, s]
                pieces = (g0 + R - 1) // CH - g0 // CH + 1
                piece_row[hub_p0[b, s]:hub_p0[b, s] + pieces] = slice_ptr[s] + ord_ + np.arange(pieces)
                ord_ += pieces
        for lane in range(32):
            r = s * 32 + lane
            if r >= n:
                continue
            cols = indices[indptr[r]:indptr[r + 1]]
            for c in cols[cols // H >= K]:
                w = win(c)
                tail_cols[(tail_g0[w, s] + t[lane, w]) * 32 + lane] = real_col(c)
                t[lane, w] += 1
        for w in range(W):
            TR = tail[w, s]
            if TR > 0:
                tg0 = tail_g0[w, s]
                pieces = (tg0 + TR - 1) // CH - tg0 // CH + 1
                piece_row[tail_p0[w, s]:tail_p0[w, s] + pieces] = slice_ptr[s] + ord_ + np.arange(pieces)
                ord_ += pieces
    return hub_words, tail_cols, piece_row


def execute_np(n, H, K, lay, hub_words, tail_cols, piece_row, z, n_segments=1, seg_len=None):
    """hsell_gather_kernel + hsell_reduce_kernel + the summation part of hsell_update_kernel.  With
    ``n_segments`` > 1 the gather vector is that many ranges of ``seg_len`` entries and hub block b is loaded as
    the pieces [b*Hs, (b+1)*Hs) of every range, one after the other (Hs = H / n_segments)."""
    partials = np.full((lay["n_partials"], 32), np.nan)
    bcb = lay["block_chunk_begin"].numpy()
    for desc, n_chunks, is_hub in ((lay["hub_chunks"], lay["n_hub_chunks"], True),
                                   (lay["tail_chunks"], lay["n_tail_chunks"], False)):
        d = desc.numpy().view(np.uint32).reshape(-1, 2).astype(np.int64)
        for c in range(n_chunks):
            p, mask = d[c, 0], d[c, 1] | (1 << (CH - 1))
            if is_hub:
                blk = int(np.searchsorted(bcb, c, side="right") - 1)
                sz = np.zeros(H + 1)                             # one hub block of z in "shared memory", + the zero slot
                if n_segments == 1:
                    seg = z[blk * H:(blk + 1) * H]
                    sz[:len(seg)] = seg
                else:
                    hs = H // n_segments
                    for sg in range(n_segments):
                        piece = z[sg * seg_len + blk * hs: sg * seg_len + min((blk + 1) * hs, seg_len)]
                        sz[sg * hs: sg * hs + len(piece)] = piece
            acc = np.zeros(32)
            for r in range(CH):
                base = (c * CH + r) * 32
                if is_hub:
                    w = hub_words[base:base + 32]
                    acc += sz[w[:, 0]] + sz[w[:, 1]]
                else:
                    cols = tail_cols[base:base + 32]
                    acc += np.where(cols >= 0, z[np.maximum(cols, 0)], 0.0)
                if (mask >> r) & 1:
                    partials[piece_row[p]] = acc
                    p += 1
                    acc = np.zeros(32)
    for start, cnt, out in lay["reduce_items"].numpy().reshape(-1, 3)[:lay["n_reduce"]]:
        partials[out] = partials[start:start + cnt].sum(axis=0)
    y = np.zeros(((n + 31) // 32) * 32)
    upd = lay["upd_rows"].numpy()
    for s, (beg, cnt) in enumerate(upd):
        y[s * 32:(s + 1) * 32] = partials[beg:beg + cnt].sum(axis=0) if cnt else 0.0
    return y[:n]


@pytest.mark.parametrize("H,K,min_entries,heavy", [(64, 3, 1, 2), (128, 6, 16, 32), (256, 2, 40, 1), (1024, 1, 32, 4),
                                                   (64, 40, 8, 3)])
@pytest.mark.parametrize("scale", [8, 10])
def test_model_reproduces_the_matvec(scale, H, K, min_entries, heavy):
    n = 1 << scale
    A = synthetic.rmat_graph_host(scale, 16, seed=7)
    deg = np.diff(A.indptr)
    perm = np.argsort(-deg, kind="stable")                       # degree-ranked, like the engine's internal order
    A = A[perm][:, perm].tocsr()
    A.sort_indices()
    indptr, indices = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    K = min(K, -(-n // H))
    hub, tail = count_np(indptr, indices, n, H, K, min_entries)
    lay = hsell_layout(torch.from_numpy(hub), torch.from_numpy(tail), heavy)
    hub_words, tail_cols, piece_row = fill_np(indptr, indices, n, H, K, hub, tail, lay)
    # every entry is stored exactly once
    stored = (hub_words != H).sum() + (tail_cols >= 0).sum()
    assert stored == A.nnz
    # every first-level row has exactly one writer; the padding pieces write the dump row
    real = piece_row[piece_row != lay["dump_row"]]
    assert np.array_equal(np.sort(real), np.arange(lay["n_rows1"]))
    rng = np.random.default_rng(scale)
    z = rng.uniform(0.5, 1.5, n)
    y = execute_np(n, H, K, lay, hub_words, tail_cols, piece_row, z)
    ref = A @ z
    assert not np.isnan(y).any()
    assert np.allclose(y, ref, rtol=1e-12, atol=0)
    if heavy <= 2:
        assert lay["n_reduce"] > 0                               # the two-level reduction really ran


@pytest.mark.parametrize("world,H,K", [(2, 64, 5), (4, 128, 3), (2, 256, 0)])
def test_model_row_partitioned_form(world, H, K):
    """The multi-GPU form: one rank's rows against the all-gathered vector (``world`` ranges of ``n_local``
    entries).  The builders see virtual columns (pygrank_b200.dist.virtual_columns), hub block b gathers
    the same positions of every range, tail columns are mapped back (real_columns / hsell_real_col)."""
    from pygrank_b200.dist import real_columns, virtual_columns
    rng = np.random.default_rng(world * 100 + H)
    n_local = 640
    n_global = world * n_local
    hs = H // world
    K = min(K, n_local // hs)
    # a power-law-ish local block: columns near the start of every range are hubs
    rows, cols = [], []
    for r in range(n_local):
        d = int(rng.integers(0, 40)) if r > 32 else int(rng.integers(200, 600))
        pos = np.minimum((rng.pareto(0.7, d) * 8).astype(np.int64), n_local - 1)
        c = np.unique(rng.integers(0, world, d) * n_local + pos)
        rows += [r] * len(c)
        cols += c.tolist()
    A = sp.coo_matrix((np.ones(len(rows)), (rows, cols)), shape=(n_local, n_global)).tocsr()
    A.sort_indices()
    v = virtual_columns(torch.from_numpy(A.indices.astype(np.int64)), n_local, world, H, K).numpy()
    Av = sp.coo_matrix((np.ones(len(rows)), (np.repeat(np.arange(n_local), np.diff(A.indptr)), v)),
                       shape=(n_local, max(int(v.max()) + 1, n_global))).tocsr()
    Av.sort_indices()
    indptr, vidx = Av.indptr.astype(np.int64), Av.indices.astype(np.int64)
    def real_col(vc):
        return int(real_columns(torch.tensor([vc]), n_local, world, H, K)[0])

    z = rng.uniform(0.5, 1.5, n_global)
    # one tail unit per slice, then one per (owner segment, slice): the tail stream is then window-major
    for W, wmin in ((1, 0), (world, 0), (world, 12)):
        window_of = (lambda v: 0) if W == 1 else (lambda v: min(real_col(v) // n_local, W - 1))
        hub, tail = count_np(indptr, vidx, n_local, H, K, 8, W, window_of, wmin)
        lay = hsell_layout(torch.from_numpy(hub), torch.from_numpy(tail), 4)
        hub_words, tail_cols, piece_row = fill_np(indptr, vidx, n_local, H, K, hub, tail, lay, real_col, W, window_of)
        assert (hub_words != H).sum() + (tail_cols >= 0).sum() == A.nnz
        y = execute_np(n_local, H, K, lay, hub_words, tail_cols, piece_row, z, n_segments=world, seg_len=n_local)
        assert np.allclose(y, A @ z, rtol=1e-12, atol=0)
        if W > 1 and wmin == 0 and lay["n_tail_chunks"] > 1:
            # window-major: the windows of the gathered positions never decrease along the tail stream
            flat = tail_cols[tail_cols >= 0]
            order = np.nonzero(tail_cols >= 0)[0]
            win = flat // n_local
            first_chunk = {}
            for pos, w_ in zip(order // (CH * 32), win):
                first_chunk.setdefault(int(w_), int(pos))
                assert pos >= first_chunk[int(w_)]
            starts = [first_chunk[k] for k in sorted(first_chunk)]
            assert starts == sorted(starts)
