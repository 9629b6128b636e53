"""The hub-blocked sliced-ELL form (csrc/hsell.cu): the device builder must re-encode the CSR
losslessly (bit-exact structure), its schedule must cover every unit exactly once, and the two-kernel
step on it must match the oracle / golden vectors with blocks, tails and unit pieces forced to appear on
small graphs (environment knobs of pygrank_b200.graph.hsell_config)."""
import numpy as np
import pytest

from conftest import GOLDEN_GRAPHS, load_golden, rel_l1

pytestmark = pytest.mark.gpu

FP64_TOL = 1e-10
FP32_TOL = 1e-5

# (block_cols, max_blocks, min_entries, heavy_parts[, tail_windows, tail_window_min]): tiny blocks -> many blocks + a
# real tail + heavy slices; the last two cut the tail into windows of the gather vector (with / without the
# short-tail exemption)
SHAPES = [(64, 3, 1, 2), (128, 16, 16, 32), (256, 2, 40, 1), (0, 16, 16, 32), (64, 3, 1, 2, 4, 0), (128, 4, 16, 32, 3, 6)]


@pytest.fixture(scope="module")
def pgb():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import pygrank_b200
    pygrank_b200.lib()
    return pygrank_b200


def _set_shape(monkeypatch, shape):
    names = ["PGB_HSELL_BLOCK_COLS", "PGB_HSELL_BLOCKS", "PGB_HSELL_MIN_ENTRIES", "PGB_HSELL_HEAVY_PARTS",
             "PGB_HSELL_TAIL_WINDOWS", "PGB_HSELL_TAIL_WINDOW_MIN"]
    for k, v in zip(names, shape):
        monkeypatch.setenv(k, str(v))


def decode(form):
    """hsell arrays -> sorted (row, col) pairs + schedule checks, in numpy."""
    CH = 32
    H, K, S = form.block_cols, form.n_blocks, form.n_slices
    slice_ptr = form.slice_ptr.cpu().numpy().astype(np.int64)            # first-level rows, slice-major
    n1 = form.n_rows1
    assert slice_ptr[0] == 0 and slice_ptr[-1] == n1 and (np.diff(slice_ptr) >= 0).all()
    dump = n1 + form.n_reduce
    assert form.n_partials == dump + 1
    piece_row = form.piece_row.cpu().numpy().astype(np.int64)[:form.n_pieces]
    real = piece_row[piece_row != dump]
    assert np.array_equal(np.sort(real), np.arange(n1))                  # every first-level row has exactly one writer
    row2slice = np.repeat(np.arange(S, dtype=np.int64), np.diff(slice_ptr))
    part2slice = np.where(piece_row == dump, -1, row2slice[np.minimum(piece_row, max(n1 - 1, 0))]) if n1 else \
        np.full(form.n_pieces, -1, dtype=np.int64)
    upd = form.upd_rows.cpu().numpy().astype(np.int64).reshape(-1, 2)
    per_slice = np.diff(slice_ptr)
    big = np.nonzero(per_slice > form.struct.heavy_parts)[0]
    small = np.setdiff1d(np.arange(S), big)
    assert np.array_equal(upd[small, 0], slice_ptr[:-1][small]) and np.array_equal(upd[small, 1], per_slice[small])
    if form.n_reduce:
        items = form.reduce_items.cpu().numpy().astype(np.int64).reshape(-1, 3)
        assert len(items) == form.n_reduce
        assert np.array_equal(items[:, 2], n1 + np.arange(form.n_reduce)) and (items[:, 1] >= 1).all() and (items[:, 1] <= 32).all()
        covered = np.concatenate([np.arange(a, a + c) for a, c, _ in items])
        want = np.concatenate([np.arange(slice_ptr[b], slice_ptr[b + 1]) for b in big])
        assert np.array_equal(covered, want)                  # groups tile the first-level rows of the big slices
        owner = np.searchsorted(slice_ptr, items[:, 0], side="right") - 1
        for b in big:
            mine = items[owner == b, 2]
            assert upd[b, 0] == mine[0] and upd[b, 1] == len(mine) and np.array_equal(mine, mine[0] + np.arange(len(mine)))
    else:
        assert len(big) == 0
    heavy = form.heavy_slices.cpu().numpy()[:form.n_heavy]
    assert np.array_equal(heavy, np.nonzero(upd[:, 1] > form.struct.heavy_parts)[0])
    bcb = form.block_chunk_begin.cpu().numpy()
    assert len(bcb) == K + 1 and bcb[0] == 0 and bcb[-1] == form.n_hub_chunks and (np.diff(bcb) >= 0).all()
    for begin, count in ((form.cta_hub_begin, form.n_hub_chunks), (form.cta_tail_begin, form.n_tail_chunks)):
        cb = begin.cpu().numpy()
        assert len(cb) == form.n_ctas + 1 and cb[0] == 0 and cb[-1] == count and (np.diff(cb) >= 0).all()

    def pieces_of(desc, n_chunks):
        d = desc.cpu().numpy().view(np.uint32).reshape(-1, 2)[:n_chunks].astype(np.int64)
        r = np.arange(CH)
        ends = ((d[:, 1:2] >> r[None, :]) & 1) | (r[None, :] == CH - 1)
        return d[:, 0:1] + np.cumsum(ends, axis=1) - ends          # piece (partial row) of every round

    rows, cols, written = [], [], []
    if form.n_hub_chunks:
        piece = pieces_of(form.hub_chunks, form.n_hub_chunks)                       # [chunks, CH]
        written.append(np.unique(piece))
        w = form.hub_words.cpu().numpy().view(np.uint32)[:form.n_hub_chunks * CH * 32].reshape(-1, CH, 32)
        blk = (np.searchsorted(bcb, np.arange(form.n_hub_chunks), side="right") - 1).astype(np.int64)
        for part in (w & 0xffff, w >> 16):
            c_idx, r_idx, lane = np.nonzero(part != H)
            loc = part[c_idx, r_idx, lane].astype(np.int64)
            assert (loc < H).all()
            sl = part2slice[piece[c_idx, r_idx]]
            assert (sl >= 0).all()                       # real entries only in pieces some slice lists
            rows.append(sl * 32 + lane)
            cols.append(blk[c_idx] * H + loc)
    if form.n_tail_chunks:
        piece = pieces_of(form.tail_chunks, form.n_tail_chunks)
        written.append(np.unique(piece))
        c = form.tail_cols.cpu().numpy()[:form.n_tail_chunks * CH * 32].reshape(-1, CH, 32)
        c_idx, r_idx, lane = np.nonzero(c >= 0)
        sl = part2slice[piece[c_idx, r_idx]]
        assert (sl >= 0).all()
        rows.append(sl * 32 + lane)
        cols.append(c[c_idx, r_idx, lane].astype(np.int64))
    # every partial row is written by exactly one piece, and every listed part is written
    written = np.concatenate(written) if written else np.zeros(0, np.int64)
    assert np.array_equal(np.sort(written), np.arange(form.n_pieces))
    rows = np.concatenate(rows) if rows else np.zeros(0, np.int64)
    cols = np.concatenate(cols) if cols else np.zeros(0, np.int64)
    order = np.lexsort((cols, rows))
    return rows[order], cols[order]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype_name", ["float32", "float64"])
def test_builder_is_lossless(pgb, monkeypatch, shape, dtype_name):
    import torch
    from pygrank_b200 import device_synthetic
    _set_shape(monkeypatch, shape)
    scale = 13
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=2)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="symmetric")
    view = g.in_view
    form = view.hsell(getattr(torch, dtype_name))
    assert form is not None
    rows, cols = decode(form)
    indptr = view.indptr.cpu().numpy().astype(np.int64)
    indices = view.indices.cpu().numpy().astype(np.int64)
    ref_rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(indptr))
    assert rows.size == indices.size
    assert np.array_equal(rows, ref_rows) and np.array_equal(cols, indices)
    if shape[0]:
        assert form.block_cols == shape[0] and form.n_blocks == min(shape[1], -(-n // shape[0]))


@pytest.mark.parametrize("name", ["ba2000", "rmat10", "gnp600d"])
@pytest.mark.parametrize("shape", SHAPES[:3])
def test_builder_lossless_on_golden_graphs(pgb, monkeypatch, name, shape):
    import torch
    _set_shape(monkeypatch, shape)
    z, A, directed = load_golden(name)
    g = pgb.DeviceGraph.from_scipy(A, directed=directed, normalization="auto")
    for view in {id(g.in_view): g.in_view, id(g.out_view): g.out_view}.values():
        form = view.hsell(torch.float64)
        rows, cols = decode(form)
        indptr = view.indptr.cpu().numpy().astype(np.int64)
        assert np.array_equal(rows, np.repeat(np.arange(view.n, dtype=np.int64), np.diff(indptr)))
        assert np.array_equal(cols, view.indices.cpu().numpy().astype(np.int64))


def _runs(P):
    return {
        "ppr85": ("auto", lambda kw: P.PageRank(0.85, tol=1e-9, max_iters=1000, **kw)),
        "ppr85_col": ("col", lambda kw: P.PageRank(0.85, tol=1e-9, max_iters=1000, **kw)),
        "ppr90_noq": ("auto", lambda kw: P.PageRank(0.9, tol=1e-9, use_quotient=False, max_iters=1000, **kw)),
        "heat3_tol9": ("auto", lambda kw: P.HeatKernel(3, tol=1e-9, **kw)),
        "gen40": ("auto", lambda kw: P.GenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters",
                                                          max_iters=41, **kw)),
        "absorb85": ("auto", lambda kw: P.AbsorbingWalks(0.85, tol=1e-9, max_iters=1000, **kw)),
    }


@pytest.mark.parametrize("name", ["ba2000", "rmat10", "gnp600d"])
@pytest.mark.parametrize("run", ["ppr85", "ppr85_col", "ppr90_noq", "heat3_tol9", "gen40", "absorb85"])
@pytest.mark.parametrize("shape", SHAPES[:3])
def test_filters_on_forced_shapes_match_golden(pgb, monkeypatch, name, run, shape):
    """Same bars as test_gpu_parity (iteration counts equal, fp64 <= 1e-10, fp32 <= 1e-5) with the hsell
    form cut into many small blocks / unit pieces / a real tail."""
    import torch
    _set_shape(monkeypatch, shape)
    z, A, directed = load_golden(name)
    norm, make = _runs(pgb)[run]
    g = pgb.DeviceGraph.from_scipy(A, directed=directed, normalization=norm)
    P = z["P"]
    for c in range(min(P.shape[1], 2)):
        alg = make({"dtype": torch.float64})
        r = alg(g, P[:, c])
        assert g.in_view.hsell(torch.float64) is not None
        assert alg.convergence.iteration == int(z[f"run_{run}_iters"][c]), (name, run, c)
        assert rel_l1(r.numpy(), z[f"run_{run}_scores"][:, c]) <= FP64_TOL, (name, run, c)
        alg32 = make({"dtype": torch.float32})
        r32 = alg32(g, P[:, c])
        assert abs(alg32.convergence.iteration - int(z[f"run_{run}_iters"][c])) <= 1
        assert rel_l1(r32.numpy(), z[f"run_{run}_scores"][:, c]) <= FP32_TOL


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[3], (4096, 4, 16, 4)])
def test_hsell_equals_item_stream_kernel_rmat17(pgb, monkeypatch, shape):
    """The two per-iteration kernels are interchangeable: same conv, same PPR iterates (fp64 to rounding),
    on a graph with hubs, 38 % empty rows and units cut into pieces."""
    import torch
    from pygrank_b200 import _capi as C
    from pygrank_b200 import device_synthetic
    _set_shape(monkeypatch, shape)
    scale = 17
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=4)
    g = pgb.DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                   normalization="symmetric")
    gen = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    p = torch.zeros(n, dtype=torch.float64, device="cuda")
    p[torch.randint(0, n, (10,), device="cuda", generator=gen)] = 1.0
    lib = C.lib()
    try:
        y4 = g.conv(x)
        a4 = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
        r4 = a4(g, p).np
        C.check(lib.pgb_set_kernel_variant(3))
        y3 = g.conv(x)
        a3 = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
        r3 = a3(g, p).np
    finally:
        C.check(lib.pgb_set_kernel_variant(4))
    assert float((y4 - y3).abs().sum()) <= 1e-14 * float(y3.abs().sum())
    assert a4.convergence.iteration == a3.convergence.iteration
    assert float((r4 - r3).abs().sum()) <= 1e-13 * float(r3.abs().sum())
    e4, e3 = a4.convergence.errors.cpu().numpy(), a3.convergence.errors.cpu().numpy()
    assert np.allclose(e4, e3, rtol=1e-9, atol=0)


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[3]])
def test_in_kernel_graph_dropout(pgb, monkeypatch, shape):
    """K8: graph_dropout drawn inside the gather kernel (pgb_hsell_set_dropout; semantics of
    /root/reference/pygrank/core/backend/pytorch.py:34-38).  Independent masks per step make the expectation of a
    polynomial filter the undropped filter exactly; runs repeat under torch.manual_seed; hub and tail entries are both
    masked (tiny blocks force a tail)."""
    import torch
    _set_shape(monkeypatch, shape)
    z, A, directed = load_golden("ba2000")
    g = pgb.DeviceGraph.from_scipy(A, directed=directed, normalization="symmetric")
    p = z["P"][:, 0].copy()
    make = lambda: pgb.GenericGraphFilter([0.5, 0.3, 0.2, 0.1], error_type="iters", max_iters=5)
    base = make()(g, p).numpy()
    assert np.allclose(make()(g, p, graph_dropout=0).numpy(), base, rtol=1e-12, atol=0)   # RED order: not bitwise
    torch.manual_seed(3)
    first = make()(g, p, graph_dropout=0.25).numpy()
    assert np.abs(first - base).sum() > 1e-3 * np.abs(base).sum()
    runs = [first] + [make()(g, p, graph_dropout=0.25).numpy() for _ in range(39)]
    assert np.abs(runs[1] - runs[0]).sum() > 0                         # a fresh mask per call
    mean = np.mean(runs, axis=0)
    assert rel_l1(mean, base) <= 0.04, rel_l1(mean, base)              # E[filter] = undropped filter (1/sqrt(40) noise)
    # the plain conv of a dropped graph (the plugin route's graph_dropout): expectation preserved, mask fresh per call
    x = torch.ones(g.n, dtype=torch.float64, device="cuda")
    full = g.conv(x)
    d1, d2 = g.dropout(0.5).conv(x), g.dropout(0.5).conv(x)
    assert abs(float(d1.sum()) / float(full.sum()) - 1.0) < 0.05 and float((d1 - d2).abs().sum()) > 0
    # repeatable: same torch seed and the same sequence of calls -> the same masks
    from pygrank_b200 import graph as G
    G._dropout_calls[0] = 0
    torch.manual_seed(11)
    a = make()(g, p, graph_dropout=0.25).numpy()
    G._dropout_calls[0] = 0
    torch.manual_seed(11)
    b = make()(g, p, graph_dropout=0.25).numpy()
    assert np.allclose(a, b, rtol=1e-12, atol=0)
    # PageRank without the quotient runs fused with dropout; with it, it refuses
    r = pgb.PageRank(0.85, tol=1e-9, max_iters=200, use_quotient=False, error_type="iters")
    r.convergence.max_iters = 20
    out = r(g, p, graph_dropout=0.1).numpy()
    assert np.isfinite(out).all() and out.sum() > 0
    with pytest.raises(Exception, match="use_quotient"):
        pgb.PageRank(0.85)(g, p, graph_dropout=0.1)


@pytest.mark.parametrize("shape", [SHAPES[0], SHAPES[1], SHAPES[3], SHAPES[5]])
@pytest.mark.parametrize("normalization", ["symmetric", "col"])
def test_weighted_hsell_equals_item_stream_kernel(pgb, monkeypatch, shape, normalization):
    """Weighted graphs on the hub-blocked form (edge values next to the indices: pgb_hsell.hub_vals / tail_vals) against
    the item-stream kernel (the round-1 path of every weighted graph): same conv, same PPR / heat-kernel iterates, on
    forced small blocks (hub units cut into pieces, bank-ordered slots, a real tail, tail windows) and the default shape."""
    import torch
    from pygrank_b200 import _capi as C
    from pygrank_b200 import device_synthetic
    _set_shape(monkeypatch, shape)
    scale = 15
    n = 1 << scale
    src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=6)
    gen = torch.Generator(device="cuda").manual_seed(3)
    wts = torch.rand(src.numel(), dtype=torch.float64, device="cuda", generator=gen) * 4 + 0.25
    g = pgb.DeviceGraph.from_edges(n, src, dst, weights=wts, directed=False, drop_self_loops=True,
                                   normalization=normalization)
    assert g.in_view.weighted
    x = torch.rand(n, dtype=torch.float64, device="cuda", generator=gen)
    p = torch.zeros(n, dtype=torch.float64, device="cuda")
    p[torch.randint(0, n, (10,), device="cuda", generator=gen)] = 1.0
    lib = C.lib()
    out = {}
    try:
        for variant in (4, 3):
            C.check(lib.pgb_set_kernel_variant(variant))
            res = {}
            for dtype in (torch.float64, torch.float32):
                a = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=dtype)
                h = pgb.HeatKernel(3, tol=1e-9, dtype=dtype)
                res[dtype] = (g.conv(x.to(dtype)), a(g, p.to(dtype)).np, a.convergence.iteration,
                              h(g, p.to(dtype)).np, h.convergence.iteration)
            out[variant] = res
    finally:
        C.check(lib.pgb_set_kernel_variant(4))
    form = g.in_view.hsell(torch.float64)
    assert form is not None and form.hub_vals is not None and form.struct.hub_vals
    if shape[0]:
        assert form.n_tail_chunks > 0 and form.n_hub_chunks > 0
    for dtype, tol, slack in ((torch.float64, 1e-13, 0), (torch.float32, 2e-6, 1)):
        y4, r4, i4, h4, hi4 = out[4][dtype]
        y3, r3, i3, h3, hi3 = out[3][dtype]
        assert float((y4 - y3).abs().sum()) <= tol * float(y3.abs().sum())
        assert abs(i4 - i3) <= slack and abs(hi4 - hi3) <= slack
        assert float((r4 - r3).abs().sum()) <= 10 * tol * float(r3.abs().sum())
        assert float((h4 - h3).abs().sum()) <= 10 * tol * float(h3.abs().sum())


def test_weighted_forms_stay_off_transient_and_deterministic_views(pgb, monkeypatch):
    """Views whose values change per call (the plugin route's dropout masks) and the deterministic mode keep weighted
    graphs on the item stream; PGB_HSELL_WEIGHTED=0 switches the weighted forms off."""
    import torch
    from pygrank_b200 import device_synthetic
    src, dst = device_synthetic.rmat_edges_device(12, 8, seed=1)
    wts = torch.rand(src.numel(), dtype=torch.float64, device="cuda") + 0.5
    g = pgb.DeviceGraph.from_edges(1 << 12, src, dst, weights=wts, normalization="symmetric")
    view = g.in_view
    assert view.hsell(torch.float32) is not None
    assert view.with_values(view.values(torch.float64) * 2).hsell(torch.float32) is None
    monkeypatch.setenv("PGB_HSELL_WEIGHTED", "0")
    g2 = pgb.DeviceGraph.from_edges(1 << 12, src, dst, weights=wts, normalization="symmetric")
    assert g2.in_view.hsell(torch.float32) is None
    monkeypatch.delenv("PGB_HSELL_WEIGHTED")
    monkeypatch.setenv("PGB_DETERMINISTIC", "1")
    g3 = pgb.DeviceGraph.from_edges(1 << 12, src, dst, weights=wts, normalization="symmetric")
    assert g3.in_view.hsell(torch.float32) is None
    x = torch.rand(1 << 12, dtype=torch.float64, device="cuda")
    assert torch.allclose(g3.conv(x), g.conv(x), rtol=1e-12, atol=0)
