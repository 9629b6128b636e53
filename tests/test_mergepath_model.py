"""CPU model of the merge-path tile decomposition used by csrc/spmv_fused.cu.

There is no GPU in the build container, so the index arithmetic of the kernel (tile partition,
per-thread item walk, rows cut by tile boundaries and their "last arrival completes the row"
protocol) is restated here statement by statement in Python with tiny tile sizes and checked
against a plain CSR row-sum.  The GPU tests then check the real kernel against the oracle.
"""
import numpy as np
import pytest


def partition(indptr, n, nnz, tile_items):
    n_tiles = -(-(n + nnz) // tile_items)
    tile_row = np.zeros(n_tiles + 1, dtype=np.int64)
    for t in range(n_tiles + 1):
        if t == n_tiles:
            tile_row[t] = n
            continue
        target = t * tile_items
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if indptr[mid + 1] + mid >= target:
                hi = mid
            else:
                lo = mid + 1
        tile_row[t] = lo
    return n_tiles, tile_row


def model_spmv(indptr, indices, z, block, ipt, tile_order=None):
    n, nnz = len(indptr) - 1, len(indices)
    tile_items = block * ipt
    n_tiles, tile_row = partition(indptr, n, nnz, tile_items)
    total_items = n + nnz
    out = np.full(n, np.nan)
    written = np.zeros(n, dtype=int)
    span_acc = np.zeros(max(n_tiles, 1))
    span_cnt = np.zeros(max(n_tiles, 1), dtype=np.int64)

    def update(row, acc, deg):
        assert deg == indptr[row + 1] - indptr[row], (row, deg)
        out[row] = acc
        written[row] += 1

    order = range(n_tiles) if tile_order is None else tile_order
    for tile in order:
        item_lo = tile * tile_items
        item_hi = min(item_lo + tile_items, total_items)
        r_lo, r_hi = tile_row[tile], tile_row[tile + 1]
        e_lo, e_hi = item_lo - r_lo, item_hi - r_hi
        nrows, nedges = r_hi - r_lo, e_hi - e_lo
        assert nrows >= 0 and nedges >= 0 and nrows + nedges == item_hi - item_lo
        row0_begin = indptr[r_lo]
        start0 = max(row0_begin, e_lo) - e_lo
        s_end = [indptr[r_lo + 1 + k] - e_lo for k in range(nrows)]
        s_rowsum = [0.0] * (nrows + 1)
        s_val = [z[indices[e_lo + i]] for i in range(nedges)]
        nitems = nrows + nedges
        for tid in range(block):
            d = tid * ipt
            if d >= nitems:
                continue
            lo, hi = 0, nrows
            while lo < hi:
                mid = (lo + hi) // 2
                if s_end[mid] + mid < d:
                    lo = mid + 1
                else:
                    hi = mid
            k, ec = lo, d - lo
            run, first = 0.0, True
            stop = ipt if d + ipt < nitems else nitems - d
            for _ in range(stop):
                if k < nrows and ec == s_end[k]:
                    if first:
                        s_rowsum[k] += run
                    else:
                        assert s_rowsum[k] == 0.0
                        s_rowsum[k] = run
                    first = False
                    run = 0.0
                    k += 1
                else:
                    run += s_val[ec]
                    ec += 1
            s_rowsum[k] += run
        lead_span = nrows > 0 and row0_begin < e_lo
        for k in range(nrows):
            if k == 0 and lead_span:
                continue
            deg = s_end[k] - (s_end[k - 1] if k else start0)
            update(r_lo + k, s_rowsum[k], deg)
        has_trail = (r_hi < n) and ((s_end[nrows - 1] < nedges) if nrows > 0 else nedges > 0)
        for who in ("lead", "trail"):
            if (who == "lead" and not lead_span) or (who == "trail" and not has_trail):
                continue
            r = r_lo if who == "lead" else r_hi
            partial = s_rowsum[0] if who == "lead" else s_rowsum[nrows]
            b, e = indptr[r], indptr[r + 1]
            t_a, t_b = (b + r) // tile_items, (e + r) // tile_items
            expected = t_b - t_a + 1
            assert expected >= 2
            if who == "lead":
                assert t_b == tile
            span_acc[t_b] += partial
            arrived = span_cnt[t_b]
            span_cnt[t_b] += 1
            if arrived == expected - 1:
                total = span_acc[t_b]
                span_acc[t_b] = 0.0
                span_cnt[t_b] = 0
                update(r, total, e - b)
    assert np.all(written == 1), np.flatnonzero(written != 1)
    assert not span_acc.any() and not span_cnt.any()       # workspace self-resets
    return out


def random_csr(rng, n, kind):
    if kind == "powerlaw":
        deg = np.minimum((rng.pareto(1.0, n) * 2).astype(int), n)
        deg[rng.uniform(size=n) < 0.35] = 0
    elif kind == "star":
        deg = np.zeros(n, dtype=int)
        deg[n // 3] = n
        deg[0] = 3
    elif kind == "empty":
        deg = np.zeros(n, dtype=int)
    elif kind == "dense_rows":
        deg = np.full(n, 7)
    else:  # long rows at the edges
        deg = rng.integers(0, 4, n)
        deg[0] = 50
        deg[-1] = 61
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    deg = np.minimum(deg, n)
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n, d, replace=False)) for d in deg] + [np.zeros(0, int)]).astype(np.int64)
    return indptr, indices


@pytest.mark.parametrize("kind", ["powerlaw", "star", "empty", "dense_rows", "edges"])
@pytest.mark.parametrize("block,ipt", [(4, 3), (8, 5), (2, 1), (32, 9)])
def test_model_matches_row_sums(kind, block, ipt):
    rng = np.random.default_rng(abs(hash((len(kind), block, ipt))) % 2 ** 32)
    for n in (1, 2, 17, 64, 301):
        indptr, indices = random_csr(rng, n, kind)
        z = rng.integers(1, 100, n).astype(np.float64)        # integers: sums are order independent
        expect = np.array([z[indices[indptr[i]:indptr[i + 1]]].sum() for i in range(n)])
        n_tiles = -(-(n + len(indices)) // (block * ipt))
        got = model_spmv(indptr, indices, z, block, ipt)
        assert np.array_equal(got, expect)
        # arrival order of the tiles must not matter (no ordering assumption between CTAs)
        order = rng.permutation(n_tiles)
        got = model_spmv(indptr, indices, z, block, ipt, tile_order=order)
        assert np.array_equal(got, expect)


# ------------------------------------------------------------------------------------------------
# v2: warp-autonomous tiles (warp_tile_kernel).  LANES plays the role of the 32 lanes; a tile is
# `subs` passes of LANES*ipt items.
def model_spmv_v2(indptr, indices, z, lanes, ipt, subs, tile_order=None):
    n, nnz = len(indptr) - 1, len(indices)
    sub_items = lanes * ipt
    tile_items = sub_items * subs
    n_tiles, tile_row = partition(indptr, n, nnz, tile_items)
    total_items = n + nnz
    out = np.full(n, np.nan)
    written = np.zeros(n, dtype=int)
    span_acc = np.zeros(max(n_tiles, 1))
    span_cnt = np.zeros(max(n_tiles, 1), dtype=np.int64)

    def update(row, acc, deg):
        assert deg == indptr[row + 1] - indptr[row]
        out[row] = acc
        written[row] += 1

    def commit(slot, partial, expected, row):
        assert expected >= 2
        span_acc[slot] += partial
        arrived = span_cnt[slot]
        span_cnt[slot] += 1
        if arrived == expected - 1:
            total = span_acc[slot]
            span_acc[slot] = 0.0
            span_cnt[slot] = 0
            update(row, total, indptr[row + 1] - indptr[row])

    order = range(n_tiles) if tile_order is None else tile_order
    for tile in order:
        item_lo = tile * tile_items
        item_hi = min(item_lo + tile_items, total_items)
        r_lo, r_hi = int(tile_row[tile]), int(tile_row[tile + 1])
        e_lo, e_hi = item_lo - r_lo, item_hi - r_hi
        lead_span = (r_hi > r_lo) and (indptr[r_lo] < e_lo)
        r_cur, carry = r_lo, 0.0
        I0 = item_lo
        while I0 < item_hi:
            nitems = min(item_hi - I0, sub_items)
            e_cur = I0 - r_cur
            # 1. marker mask, batches of `lanes` rows
            m = [0] * ipt
            nrows, b = 0, 0
            while True:
                cnt = 0
                for lane in range(lanes):
                    row = r_cur + b * lanes + lane
                    pos = nitems
                    if row < n:
                        pos = indptr[row + 1] + row - I0
                    if pos < nitems:
                        assert pos >= 0
                        m[int(pos) // lanes] |= 1 << (int(pos) % lanes)
                        cnt += 1
                nrows += cnt
                if cnt < lanes:
                    break
                b += 1
            # 2. striped gather into item space
            item = [0.0] * sub_items
            pre = 0
            s_pre = [0] * (ipt + 1)
            for s in range(ipt):
                s_pre[s] = pre
                for lane in range(lanes):
                    p = s * lanes + lane
                    is_marker = (m[s] >> lane) & 1
                    rank = pre + bin(m[s] & ((1 << lane) - 1)).count("1")
                    e = e_cur + p - rank
                    if p < nitems and not is_marker:
                        assert e_lo <= e < e_hi
                        item[p] = z[indices[e]]
                pre += bin(m[s]).count("1")
            assert pre == nrows
            # 3. blocked merge
            rowsum = [None] * (sub_items + 1)
            tails, closed, firsts, heads = [], [], [], []
            allbits = 0
            for s in range(ipt):
                allbits |= int(m[s]) << (s * lanes)
            for lane in range(lanes):
                p0 = lane * ipt
                w0, sh = p0 // lanes, p0 % lanes
                lo = m[w0]
                k = s_pre[w0] + bin(lo & ((1 << sh) - 1)).count("1")
                bits = (allbits >> p0) & ((1 << ipt) - 1)   # funnel shift of (lo, hi); ipt <= lanes assumed below
                run, head, first_k = 0.0, 0.0, -1
                for j in range(ipt):
                    if (bits >> j) & 1:
                        if first_k < 0:
                            head, first_k = run, k
                        else:
                            assert rowsum[k] is None
                            rowsum[k] = run
                        run = 0.0
                        k += 1
                    else:
                        run += item[p0 + j]
                tails.append(run)
                closed.append(first_k >= 0)
                firsts.append(first_k)
                heads.append(head)
            # segmented inclusive scan (Hillis-Steele with flags), exactly as the shuffles do it
            seg, f = list(tails), list(closed)
            d = 1
            while d < lanes:
                nseg, nf = list(seg), list(f)
                for lane in range(lanes):
                    if lane >= d and not f[lane]:
                        nseg[lane] = seg[lane] + seg[lane - d]
                        nf[lane] = f[lane - d]
                seg, f = nseg, nf
                d <<= 1
            for lane in range(lanes):
                carry_in = seg[lane - 1] if lane > 0 else 0.0
                if not any(closed[:lane]):
                    carry_in += carry
                if closed[lane]:
                    assert rowsum[firsts[lane]] is None
                    rowsum[firsts[lane]] = heads[lane] + carry_in
            carry = (carry + seg[lanes - 1]) if not any(closed) else seg[lanes - 1]
            # 4. row updates
            for k in range(nrows):
                row = r_cur + k
                assert rowsum[k] is not None
                b_, e_ = indptr[row], indptr[row + 1]
                if lead_span and row == r_lo:
                    t_a = (b_ + row) // tile_items
                    commit(tile, rowsum[k], tile - t_a + 1, row)
                else:
                    update(row, rowsum[k], e_ - b_)
            r_cur += nrows
            I0 += sub_items
        assert r_cur == r_hi
        if r_hi < n:
            b_, e_ = indptr[r_hi], indptr[r_hi + 1]
            if e_hi > max(b_, e_lo):
                t_a, t_b = (b_ + r_hi) // tile_items, (e_ + r_hi) // tile_items
                commit(t_b, carry, t_b - t_a + 1, r_hi)
    assert np.all(written == 1), np.flatnonzero(written != 1)
    assert not span_acc.any() and not span_cnt.any()
    return out


@pytest.mark.parametrize("kind", ["powerlaw", "star", "empty", "dense_rows", "edges"])
@pytest.mark.parametrize("lanes,ipt,subs", [(4, 3, 2), (8, 5, 1), (8, 3, 4), (32, 9, 8), (4, 1, 3)])
def test_warp_tile_model_matches_row_sums(kind, lanes, ipt, subs):
    rng = np.random.default_rng(abs(hash((len(kind), lanes, ipt, subs))) % 2 ** 32)
    for n in (1, 2, 17, 64, 301, 1500):
        if n == 1500 and lanes * ipt * subs < 64:
            continue
        indptr, indices = random_csr(rng, n, kind)
        z = rng.integers(1, 100, n).astype(np.float64)
        expect = np.array([z[indices[indptr[i]:indptr[i + 1]]].sum() for i in range(n)])
        got = model_spmv_v2(indptr, indices, z, lanes, ipt, subs)
        assert np.array_equal(got, expect)
        n_tiles = -(-(n + len(indices)) // (lanes * ipt * subs))
        got = model_spmv_v2(indptr, indices, z, lanes, ipt, subs, tile_order=rng.permutation(n_tiles))
        assert np.array_equal(got, expect)
