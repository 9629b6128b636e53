"""Device-resident graphs and the device preprocessor.

Mirrors ``pygrank.core.utils.preprocessing`` for the hot path
(/root/reference/pygrank/core/utils/preprocessing.py):

* ``to_sparse_matrix`` / ``preprocessor``  (:50-152, :233-287)  ->  :func:`preprocessor`,
  :meth:`DeviceGraph.from_scipy`, :meth:`DeviceGraph.from_edges`
* ``Adjacency`` (:9-28)                                          ->  :class:`DeviceGraph` (has ``.array``, ``.shape``)

HBM layout (all torch CUDA tensors, owned here, handed to libpgb200 as raw pointers):

* pull CSR of the propagation operator — ``indptr int32[n+1]``, ``indices int32[nnz]`` (sources of
  every destination row, ascending), optional raw weights in the vector dtype — in the engine's
  INTERNAL node order (optionally degree-sorted so that hub entries of the gather vector share
  cache lines), plus the merge-path partition ``tile_row int32[n_tiles+1]``;
* the normalisation kept FACTORISED: degrees (row/col sums, fp64), left/right scales
  ``L = f(rowsum)``, ``R = f(colsum)`` with the reference's exact rounding, so an unweighted
  graph streams no edge values at all — ``M[j,i] = (L[j]*a_ji)*R[i]`` is applied as a pre-scaled
  gather vector and a per-row factor inside the kernels;
* ``perm``/``iperm`` (int32[n]) between user order and internal order.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import numpy as np
import torch

from . import _capi as C

_DT = {torch.float32: C.PGB_F32, torch.float64: C.PGB_F64}


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise Exception("pygrank_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def dtype_code(dtype: torch.dtype) -> int:
    if dtype not in _DT:
        raise Exception("pygrank_b200 vectors are float32 or float64, not " + str(dtype))
    return _DT[dtype]


def build_csr(n: int, row: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], flags: int):
    """COO -> canonical CSR through ``pgb_csr_build``; returns (indptr, indices, values|None)."""
    lib = C.lib()
    nnz_in = int(row.numel())
    dev = row.device
    weighted = val is not None
    keep_vals = weighted and not (flags & C.BUILD_BINARY)
    cap = nnz_in * (2 if flags & C.BUILD_SYMMETRIZE else 1)
    ws_bytes = lib.pgb_csr_build_workspace_bytes(n, nnz_in, flags, int(weighted))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    indptr = torch.empty(n + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(max(cap, 1), dtype=torch.int32, device=dev)
    values = torch.empty(max(cap, 1), dtype=torch.float64, device=dev) if keep_vals else None
    nnz_out = ctypes.c_int64(0)
    C.check(lib.pgb_csr_build(n, nnz_in, C.ptr(row), C.ptr(col), C.ptr(val), flags, C.ptr(ws), ws_bytes,
                              C.ptr(indptr), C.ptr(indices), C.ptr(values), ctypes.byref(nnz_out), C.stream_ptr()))
    del ws
    nnz = int(nnz_out.value)
    indices = indices[:nnz].clone() if nnz < cap else indices[:nnz]
    if values is not None:
        values = values[:nnz].clone() if nnz < cap else values[:nnz]
    return indptr, indices, values


def _env_int(name: str, default: int) -> int:
    v = os.environ.get(name)
    return int(v) if v not in (None, "") else default


def _env_float(name: str, default: float) -> float:
    v = os.environ.get(name)
    return float(v) if v not in (None, "") else default


def hsell_config() -> dict:
    """Knobs of the hub-blocked sliced-ELL builder (environment overrides are for tests/experiments).
    block_cols 0 = 128 KB of the gather vector per block: measured best on B200 — the rest of the SM's
    256 KB stays L1, which the L2 gathers of the tail need for their in-flight lines (with the full
    227 KB in shared memory the same tail ran 2x slower)."""
    return {
        "enabled": _env_int("PGB_HSELL", 1) != 0,
        "block_cols": _env_int("PGB_HSELL_BLOCK_COLS", 0),
        "max_blocks": _env_int("PGB_HSELL_BLOCKS", 0),
        "min_entries": _env_int("PGB_HSELL_MIN_ENTRIES", 32),
        "round_cost": _env_float("PGB_HSELL_ROUND_COST", 0.0),
        "heavy_parts": min(max(_env_int("PGB_HSELL_HEAVY_PARTS", 32), 1), 32),
        "bank_order": _env_int("PGB_HSELL_BANK_ORDER", 1) != 0,
        # tail windows: 0 = one window per segment of a row-partitioned gather vector (single GPU: one window);
        # n > 0 = that many equal windows of the gather vector; -1 = never
        "tail_windows": _env_int("PGB_HSELL_TAIL_WINDOWS", 0),
        "tail_window_min": _env_int("PGB_HSELL_TAIL_WINDOW_MIN", 32),   # shorter tails stay one unit (4 GPUs: 8 -> .750, 16 -> .726, 32 -> .719 ms)
        # 1: pieces are RED.ADDed into one accumulator row per slice and the update pass streams it (fast path);
        # 0: partial rows added in a fixed order (bit-reproducible runs; PGB_DETERMINISTIC=1 selects it too)
        "accumulate": _env_int("PGB_HSELL_ACCUM", 1) != 0 and _env_int("PGB_DETERMINISTIC", 0) == 0,
        # weighted graphs on the hub-blocked form (edge values next to the indices); 0: item-stream kernels
        "weighted": _env_int("PGB_HSELL_WEIGHTED", 1) != 0,
    }


PANEL_ELEM_BYTES = 16        # one node of a panel: 4 x fp32 or 2 x fp64 (pgb_hsell_panel_width)


def hsell_shape(dtype: torch.dtype, n_segments: int, seg_len: int, cfg: Optional[dict] = None,
                elem_bytes: Optional[int] = None):
    """(block_cols, n_blocks) the builder will use for a gather vector of ``n_segments`` ranges of
    ``seg_len`` entries (pure arithmetic: the row-partitioned path needs it before it relabels columns).
    ``elem_bytes``: bytes per node of the gather vector when it is not one ``dtype`` scalar (panel forms: 16)."""
    cfg = dict(hsell_config(), **(cfg or {}))
    if elem_bytes is None:
        eb = 4 if dtype == torch.float32 else 8
        cap = C.lib().pgb_hsell_max_block_cols(dtype_code(dtype))
    else:
        eb = int(elem_bytes)
        cap = ((232448 - 64) // eb - 1) & ~63            # same rule as pgb_hsell_max_block_cols
    H = cfg["block_cols"] if cfg["block_cols"] > 0 else (128 * 1024) // eb
    H = min(H, cap)
    H -= H % (4 * n_segments)                        # equal 16-byte aligned parts per segment
    if H < 4 * n_segments:
        raise Exception("hsell: block_cols too small")
    Hs = H // n_segments
    max_blocks = cfg["max_blocks"]
    if max_blocks <= 0:
        # auto: hub blocks cover ~1/8 of the columns, between 64 and 256 blocks.  Measured (ms per step): fp32,
        # 16.8 M columns: 64 -> .685, 96 -> .670, 128 -> .70; fp64 (blocks half as wide): 64 -> 1.00, 128 -> .965;
        # fp32, 134 M columns on 8 ranks: 64 -> 1.40, 128 -> 1.28, 256 -> 1.19, 512 -> 1.36
        # round 2 (TEX tail + RED pieces, 16.8 M columns): 48 -> .562, 64 -> .531, 80 -> .522 ms: ~1/6.4 of the columns
        # panel forms (blocks of 8192 nodes): 192 -> 1.405, 256 -> 1.394, 384 -> 1.339 ms per 4-column step
        cap_blocks = _env_int("PGB_HSELL_BLOCKS_CAP", 256 if elem_bytes is None else 384)
        max_blocks = max(64, min(cap_blocks, (5 * n_segments * seg_len) // (32 * H)))
    K = max(min(max_blocks, -(-seg_len // Hs)), 0)
    if n_segments > 1 and K * Hs > seg_len:
        K = seg_len // Hs                            # multi-segment blocks must be full; the rest is tail
    return H, K


RELABELS = ("hub", "degree", "none")


def hub_signature_shape(n: int):
    """(G, span) of the hub-signature order: regions of G degree ranks (the narrower of the two dtypes' hub
    blocks, so the membership of BOTH forms' blocks survives the reordering) up to the wider hub span."""
    H32, K32 = hsell_shape(torch.float32, 1, n)
    H64, K64 = hsell_shape(torch.float64, 1, n)
    G = min(H32, H64)
    span = min(max(H32 * K32, H64 * K64), -(-n // G) * G)
    return G, span


def stream_layout(rounds: torch.Tensor, base: int, chunk: int = C.HSELL_CHUNK):
    """Layout of one hsell stream (plain torch, device agnostic: covered by the CPU tests).

    ``rounds`` int64 [B, S]: rounds of every unit in stream order (row b = one hub block, or the single
    tail row); every row is padded to whole chunks.  A PIECE ends at the last round of a unit and at the
    end of every chunk; pieces are numbered in stream order starting at ``base``.  Returns
    (first round of every unit [B*S], number of the first piece of every unit [B*S], pieces per unit [B*S],
    chunk descriptors int32 [n_chunks, 2] = (first piece of the chunk, mask of the rounds at which a unit
    ends), n_chunks, number of pieces, first chunk of every row [B+1])."""
    i64 = torch.int64
    dev = rounds.device
    CH = chunk
    per_row = rounds.sum(1)
    row_pad = (per_row + (CH - 1)) // CH * CH
    row_base = torch.cumsum(row_pad, 0) - row_pad
    g0 = (row_base[:, None] + torch.cumsum(rounds, 1) - rounds).reshape(-1)
    r = rounds.reshape(-1)
    g1 = g0 + r
    ex = r > 0
    n_chunks = int(row_pad.sum()) // CH
    aligned = ex & (g1 % CH == 0)                       # unit end coincides with a chunk end: one piece end
    unit_rank = torch.cumsum(ex.to(i64), 0) - ex.to(i64)
    b_before = torch.cumsum(aligned.to(i64), 0) - aligned.to(i64)
    p0 = base + unit_rank + g0 // CH - b_before         # piece ends before the unit's first round
    pieces = torch.where(ex, (g1 - 1) // CH - g0 // CH + 1, torch.zeros_like(g0))
    g1e = g1[ex]
    starts = torch.arange(n_chunks, device=dev, dtype=i64) * CH
    u = torch.searchsorted(g1e, starts, right=True)     # units that end before the chunk starts
    bcum = torch.cat([torch.zeros(1, dtype=i64, device=dev), torch.cumsum(aligned[ex].to(i64), 0)])
    p_first = base + u + torch.arange(n_chunks, device=dev, dtype=i64) - bcum[u]
    endmask = torch.zeros(max(n_chunks, 1), dtype=i64, device=dev)
    if g1e.numel():
        endmask.index_add_(0, (g1e - 1) // CH, torch.ones_like(g1e) << ((g1e - 1) % CH))
    n_parts = int(ex.sum()) + n_chunks - int(aligned.sum())
    desc = torch.stack([p_first, endmask[:n_chunks]], 1)
    desc = torch.where(desc >= 2 ** 31, desc - 2 ** 32, desc).to(torch.int32).contiguous()
    chunk_begin = torch.cat([row_base, row_pad.sum().reshape(1)]) // CH
    return g0.contiguous(), p0.contiguous(), pieces, desc, n_chunks, n_parts, chunk_begin


def hsell_layout(hr: torch.Tensor, tr: torch.Tensor, heavy_parts: int) -> dict:
    """Everything of a pgb_hsell that follows from the unit sizes alone (plain torch, device agnostic; the CPU
    tests run it against a numpy model of the kernels): ``hr`` int64 [K, S] rounds of every hub unit (0 = no
    unit), ``tr`` int64 [S] or [W, S] tail rounds of every slice (per tail window, window-major stream).  Streams and pieces as in :func:`stream_layout` (hub
    stream first); first-level partial rows slice-major (a slice's pieces in block order, then its tail pieces, windows
    ascending);
    slices with more than ``heavy_parts`` pieces reduced in groups of 32 consecutive rows into second-level
    rows stored after the first level; one dump row at the very end for the padding pieces."""
    i64 = torch.int64
    dev = tr.device
    tr = tr.reshape(1, -1) if tr.dim() == 1 else tr              # [W, S]: one tail unit per (window, slice)
    W, S = int(tr.shape[0]), int(tr.shape[1])
    K = int(hr.shape[0])
    if K > 0:
        hub_g0, hub_p0, hub_pieces, hub_chunks, n_hub_chunks, n_hub_parts, bcb = stream_layout(hr, 0)
    else:
        hub_g0 = hub_p0 = torch.zeros(1, dtype=i64, device=dev)
        hub_pieces = torch.zeros(0, dtype=i64, device=dev)
        hub_chunks = torch.zeros((1, 2), dtype=torch.int32, device=dev)
        n_hub_chunks = n_hub_parts = 0
        bcb = torch.zeros(1, dtype=i64, device=dev)
    tail_g0, tail_p0, tail_pieces, tail_chunks, n_tail_chunks, n_tail_parts, _ = stream_layout(tr, n_hub_parts)
    n_pieces = n_hub_parts + n_tail_parts                       # pieces in stream order (hub stream, then tail)
    per_slice = tail_pieces.view(W, S).sum(0) + (hub_pieces.view(K, S).sum(0) if K > 0 else 0)
    slice_ptr64 = torch.zeros(S + 1, dtype=i64, device=dev)
    torch.cumsum(per_slice, 0, out=slice_ptr64[1:])
    n_rows1 = int(slice_ptr64[-1])                              # first-level partial rows, slice-major
    # slices with many pieces (hub rows) are reduced in two levels: groups of 32 consecutive
    # first-level rows -> one second-level row each (kernel B1); the slice then reads those
    big = torch.nonzero(per_slice > heavy_parts).reshape(-1)
    upd_begin = slice_ptr64[:-1].clone()
    upd_count = per_slice.clone()
    n_reduce = 0
    reduce_items = torch.zeros(3, dtype=torch.int32, device=dev)
    if big.numel():
        cnt1 = per_slice[big]
        groups = (cnt1 + 31) // 32
        n_reduce = int(groups.sum())
        owner = torch.repeat_interleave(torch.arange(big.numel(), device=dev), groups)       # item -> big slice
        first_item = torch.cumsum(groups, 0) - groups
        k = torch.arange(n_reduce, device=dev, dtype=i64) - first_item[owner]                # group index in slice
        start = slice_ptr64[big][owner] + 32 * k
        count = torch.clamp(cnt1[owner] - 32 * k, max=32)
        out_row = n_rows1 + torch.arange(n_reduce, device=dev, dtype=i64)
        reduce_items = torch.stack([start, count, out_row], 1).to(torch.int32).contiguous()
        upd_begin[big] = n_rows1 + first_item
        upd_count[big] = groups
    dump_row = n_rows1 + n_reduce                               # written by the padding pieces of the streams
    return {
        "hub_g0": hub_g0, "hub_p0": hub_p0, "tail_g0": tail_g0, "tail_p0": tail_p0,
        "hub_chunks": hub_chunks, "tail_chunks": tail_chunks, "n_hub_chunks": n_hub_chunks,
        "n_tail_chunks": n_tail_chunks, "block_chunk_begin": bcb, "n_pieces": n_pieces, "n_rows1": n_rows1,
        "n_reduce": n_reduce, "dump_row": dump_row, "n_partials": dump_row + 1,
        "slice_ptr": slice_ptr64.to(torch.int32), "reduce_items": reduce_items, "upd_count": upd_count,
        "upd_rows": torch.stack([upd_begin, upd_count], 1).to(torch.int32).contiguous(),
    }


_dropout_calls = [0]


class in_kernel_dropout:
    """``with in_kernel_dropout(p):`` — gather launches inside draw Bernoulli edge masks in the kernel
    (pgb_hsell_set_dropout).  The seed follows torch's global seed and the number of dropout scopes opened so far, so
    ``torch.manual_seed`` makes runs repeatable."""

    def __init__(self, p: float):
        self.p = float(p)

    def __enter__(self):
        if self.p > 0:
            _dropout_calls[0] += 1
            seed = (int(torch.initial_seed()) * 0x9E3779B97F4A7C15 + _dropout_calls[0]) & 0xFFFFFFFFFFFFFFFF
            C.check(C.lib().pgb_hsell_set_dropout(self.p, seed))
        return self

    def __exit__(self, *exc):
        if self.p > 0:
            C.check(C.lib().pgb_hsell_set_dropout(0.0, 0))
        return False


class HsellForm:
    """Device arrays of one pgb_hsell (kept alive here; the C struct holds raw pointers)."""

    def __init__(self, view: "CsrView", dtype: Optional[torch.dtype], n_segments: int = 1, seg_len: Optional[int] = None,
                 cfg: Optional[dict] = None, elem_bytes: Optional[int] = None, values: Optional[torch.Tensor] = None):
        lib = C.lib()
        cfg = dict(hsell_config(), **(cfg or {}))
        st = C.stream_ptr()
        dev = view.indptr.device
        n = view.n
        n_cols = view.n_cols
        seg_len = int(seg_len) if seg_len is not None else n_cols
        i64 = torch.int64
        H, K = hsell_shape(dtype, n_segments, seg_len, cfg, elem_bytes)
        eb = int(elem_bytes) if elem_bytes is not None else (4 if dtype == torch.float32 else 8)
        S = (n + 31) // 32
        CH = C.HSELL_CHUNK
        total_cols = n_segments * seg_len
        tw = cfg["tail_windows"]
        if tw > 0:
            W = min(tw, C.HSELL_MAX_WINDOWS)
            window_len = -(-total_cols // W)
        elif tw == 0 and n_segments > 1:
            W, window_len = min(n_segments, C.HSELL_MAX_WINDOWS), -(-total_cols // min(n_segments, C.HSELL_MAX_WINDOWS))
        else:
            W, window_len = 1, max(total_cols, 1)
        hub_rounds = torch.zeros(max(K * S, 1), dtype=torch.int32, device=dev)
        tail_rounds = torch.zeros(max(W * S, 1), dtype=torch.int32, device=dev)
        C.check(lib.pgb_hsell_count(n, C.ptr(view.indptr), C.ptr(view.indices), H, K, cfg["min_entries"],
                                    float(cfg["round_cost"]), n_segments, seg_len, W, window_len,
                                    int(cfg["tail_window_min"]), C.ptr(hub_rounds), C.ptr(tail_rounds), st))

        hr = hub_rounds[:K * S].to(i64).view(K, S) if K > 0 else torch.zeros((0, S), dtype=i64, device=dev)
        lay = hsell_layout(hr, tail_rounds[:W * S].to(i64).view(W, S), cfg["heavy_parts"])
        self.n_windows, self.window_len = W, window_len
        hub_g0, hub_p0, tail_g0, tail_p0 = lay["hub_g0"], lay["hub_p0"], lay["tail_g0"], lay["tail_p0"]
        self.hub_chunks, self.tail_chunks = lay["hub_chunks"], lay["tail_chunks"]
        n_hub_chunks, n_tail_chunks, bcb = lay["n_hub_chunks"], lay["n_tail_chunks"], lay["block_chunk_begin"]
        n_pieces, n_rows1, n_reduce, n_partials = lay["n_pieces"], lay["n_rows1"], lay["n_reduce"], lay["n_partials"]
        heavy_parts, upd_count = cfg["heavy_parts"], lay["upd_count"]
        n_hub_words, n_tail_words = n_hub_chunks * CH * 32, n_tail_chunks * CH * 32
        if max(n_hub_words, n_tail_words) >= 2 ** 31 or max(n_pieces, n_rows1) >= 2 ** 31 - 2 ** 20:
            raise Exception("hsell: graph exceeds the 32-bit offsets of one device; row-partition it")
        if n_partials >= 2 ** 27:
            raise Exception("hsell: more than 2^27 partial rows; row-partition the graph")
        self.slice_ptr, self.reduce_items, self.upd_rows = lay["slice_ptr"], lay["reduce_items"], lay["upd_rows"]
        pad_word = (H | (H << 16))
        pad_word = pad_word - 2 ** 32 if pad_word >= 2 ** 31 else pad_word
        self.hub_words = torch.full((max(n_hub_words, 1),), pad_word, dtype=torch.int32, device=dev)
        self.tail_cols = torch.full((max(n_tail_words, 1),), -1, dtype=torch.int32, device=dev)
        self.piece_row = torch.full((max(n_pieces, 1),), lay["dump_row"], dtype=torch.int32, device=dev)
        scratch = torch.empty(max(view.nnz, 1), dtype=torch.int32, device=dev) if cfg["bank_order"] else None
        # weighted graphs: the edge values travel next to the indices (zero in padding slots)
        self.hub_vals = self.tail_vals = None
        if values is not None:
            if values.dtype != dtype or values.numel() != view.nnz:
                raise Exception("hsell: edge values must have the form's dtype and one entry per stored edge")
            self.hub_vals = torch.zeros(max(2 * n_hub_words, 1), dtype=dtype, device=dev)
            self.tail_vals = torch.zeros(max(n_tail_words, 1), dtype=dtype, device=dev)
        C.check(lib.pgb_hsell_fill(n, C.ptr(view.indptr), C.ptr(view.indices), H, K, n_segments, seg_len,
                                   C.ptr(hub_rounds), C.ptr(tail_rounds), C.ptr(hub_g0), C.ptr(hub_p0),
                                   C.ptr(tail_g0), C.ptr(tail_p0), C.ptr(self.slice_ptr), C.ptr(self.hub_words),
                                   C.ptr(self.tail_cols), C.ptr(self.piece_row), C.ptr(scratch),
                                   128 // eb, W, window_len,             # shared-memory banks per element
                                   dtype_code(dtype) if values is not None else C.PGB_F32, C.ptr(values),
                                   C.ptr(self.hub_vals), C.ptr(self.tail_vals), st))
        del scratch
        # slice of every piece (accumulate mode): first-level rows are slice-major, padding pieces -> row n_slices
        ps = torch.searchsorted(lay["slice_ptr"].to(i64), self.piece_row.to(i64), right=True) - 1
        self.piece_slice = torch.where(self.piece_row.to(i64) >= n_rows1, torch.full_like(ps, S), ps).to(torch.int32)
        del ps
        self.heavy_slices = torch.nonzero(upd_count > heavy_parts).reshape(-1).to(torch.int32)
        n_heavy = int(self.heavy_slices.numel())
        if n_heavy == 0:
            self.heavy_slices = torch.zeros(1, dtype=torch.int32, device=dev)
        # schedule: chunks are uniform work, so every CTA gets a contiguous, equally long range of each stream
        n_ctas = lib.pgb_device_sm_count(dev.index if dev.index is not None else torch.cuda.current_device())
        if n_ctas <= 0:
            raise Exception("pgb200: " + lib.pgb_last_error().decode())

        def cut(count):
            return ((torch.arange(n_ctas + 1, device=dev, dtype=i64) * count) // n_ctas).to(torch.int32)

        self.cta_hub_begin = cut(n_hub_chunks)
        self.cta_tail_begin = cut(n_tail_chunks)
        self.block_chunk_begin = bcb.to(torch.int32)
        self.n_partials, self.block_cols, self.n_blocks, self.n_ctas = n_partials, H, K, n_ctas
        self.n_hub_chunks, self.n_tail_chunks, self.n_heavy = n_hub_chunks, n_tail_chunks, n_heavy
        self.n_reduce, self.n_pieces, self.n_rows1 = n_reduce, n_pieces, n_rows1
        self.n_hub_words, self.n_tail_words = n_hub_words, n_tail_words
        self.n_slices, self.n_segments, self.seg_len = S, n_segments, seg_len
        self.dtype, self.elem_bytes = dtype, eb
        self.struct = C.Hsell(n, S, n_partials, seg_len, n_segments, H, K, n_ctas, n_hub_chunks, n_tail_chunks,
                              n_heavy, heavy_parts, n_reduce, 0, C.ptr(self.hub_chunks), C.ptr(self.tail_chunks),
                              C.ptr(self.hub_words), C.ptr(self.tail_cols), C.ptr(self.piece_row),
                              C.ptr(self.upd_rows), C.ptr(self.heavy_slices), C.ptr(self.reduce_items),
                              C.ptr(self.block_chunk_begin),
                              C.ptr(self.cta_hub_begin), C.ptr(self.cta_tail_begin), C.ptr(self.piece_slice),
                              C.ptr(self.hub_vals), C.ptr(self.tail_vals))

    def nbytes(self) -> int:
        return sum(int(t.numel()) * t.element_size() for t in (self.hub_chunks, self.tail_chunks, self.hub_words,
                                                               self.tail_cols, self.slice_ptr, self.piece_row,
                                                               self.upd_rows, self.hub_vals, self.tail_vals)
                   if t is not None)


class CsrView:
    """One CSR structure in HBM with its merge-path partition and cross-tile workspace."""

    def __init__(self, n: int, indptr: torch.Tensor, indices: torch.Tensor, values64: Optional[torch.Tensor]):
        lib = C.lib()
        self.n = int(n)
        self.nnz = int(indices.numel())
        self.indptr, self.indices = indptr, indices
        self._values = {torch.float64: values64} if values64 is not None else {}
        self.weighted = values64 is not None
        self.tile_items = lib.pgb_tile_items()
        self.n_tiles = (self.n + self.nnz + self.tile_items - 1) // self.tile_items
        self.tile_row = torch.empty(self.n_tiles + 1, dtype=torch.int32, device=indptr.device)
        C.check(lib.pgb_mergepath_partition(self.n, self.nnz, C.ptr(indptr), self.n_tiles, C.ptr(self.tile_row),
                                            C.stream_ptr()))
        self._ws = None
        self._istream = None
        self._vstream = {}
        self._hsell = {}
        self.transient_values = False # with_values(): values replaced per call
        self.n_cols = self.n          # length of the gather vector (larger than n for a row-partitioned slice)
        self.hsell_segments = (1, None)

    def hsell(self, dtype: torch.dtype) -> Optional[HsellForm]:
        """Hub-blocked sliced-ELL form for this dtype (None when disabled).  Weighted graphs carry their edge values
        in the form (accumulate mode, single-GPU views); views whose values change per call (``with_values``: the
        plugin route's dropout masks) stay on the item stream — a form per call would cost more than it saves."""
        cfg = hsell_config()
        if self.nnz == 0 or not cfg["enabled"]:
            return None
        if self.weighted and (self.transient_values or not cfg["accumulate"] or not cfg["weighted"]
                              or self.hsell_segments[0] != 1):
            return None
        if dtype not in self._hsell:
            self._hsell[dtype] = HsellForm(self, dtype, self.hsell_segments[0], self.hsell_segments[1],
                                           values=self.values(dtype) if self.weighted else None)
        return self._hsell[dtype]

    def hsell_panel(self) -> Optional[HsellForm]:
        """Hub-blocked form for PANELS of seed columns (16 bytes per node: 4 x fp32 or 2 x fp64, blocks of 8192 nodes);
        one form serves both dtypes.  None for weighted / row-partitioned views or when hsell is disabled."""
        if self.weighted or self.nnz == 0 or not hsell_config()["enabled"] or self.hsell_segments[0] != 1:
            return None
        if "panel" not in self._hsell:
            self._hsell["panel"] = HsellForm(self, None, 1, None, elem_bytes=PANEL_ELEM_BYTES)
        return self._hsell["panel"]

    def istream(self) -> torch.Tensor:
        """Item-space index stream (row entries + terminator -1-deg) read by the fused kernels."""
        if self._istream is None:
            self._build_streams(None)
        return self._istream

    def vstream(self, dtype: torch.dtype) -> Optional[torch.Tensor]:
        if not self.weighted:
            return None
        if dtype not in self._vstream:
            self._build_streams(dtype)
        return self._vstream[dtype]

    def _build_streams(self, dtype):
        lib = C.lib()
        dev = self.indptr.device
        # the index stream is written once and never replaced: C structs hold its raw pointer
        ist = self._istream if self._istream is not None else torch.empty(self.n + self.nnz, dtype=torch.int32,
                                                                          device=dev)
        vals = self.values(dtype) if (dtype is not None and self.weighted) else None
        vst = torch.empty(self.n + self.nnz, dtype=dtype, device=dev) if vals is not None else None
        C.check(lib.pgb_build_item_stream(self.n, self.nnz, C.ptr(self.indptr), C.ptr(self.indices),
                                          dtype_code(dtype) if vals is not None else C.PGB_F32, C.ptr(vals),
                                          C.ptr(ist), C.ptr(vst), C.stream_ptr()))
        self._istream = ist
        if vst is not None:
            self._vstream[dtype] = vst

    def values(self, dtype: torch.dtype) -> Optional[torch.Tensor]:
        if not self.weighted:
            return None
        if dtype not in self._values:
            self._values[dtype] = self._values[torch.float64].to(dtype)
        return self._values[dtype]

    def cstruct(self, dtype: torch.dtype, hsell: bool = True) -> C.Csr:
        form = self.hsell(dtype) if hsell else None
        return C.Csr(self.n, self.nnz, C.ptr(self.indptr), C.ptr(self.indices), C.ptr(self.values(dtype)),
                     C.ptr(self.tile_row), self.n_tiles, self.tile_items, C.ptr(self.istream()),
                     C.ptr(self.vstream(dtype)), ctypes.addressof(form.struct) if form is not None else None)

    def kernels_per_step(self, dtype: torch.dtype, hsell: bool = True) -> int:
        """Kernels one fused step launches (bench.py counts its own launches): gather + update on the hsell
        form, plus the second-level reduction when some slice needs it; one on the item stream."""
        form = self._hsell.get(dtype) if hsell else None
        if form is not None and hsell_config()["accumulate"]:
            return 2
        return 1 if form is None else (3 if form.n_reduce else 2)

    def new_span_ws(self, dtype: Optional[torch.dtype] = None):
        """Zeroed cross-tile workspace (one per concurrently running filter); with a dtype also the
        partial-row buffer of the hsell form."""
        dev = self.indptr.device
        acc = torch.zeros(max(self.n_tiles, 1), dtype=torch.float64, device=dev)
        cnt = torch.zeros(max(self.n_tiles, 1), dtype=torch.int32, device=dev)
        form = self.hsell(dtype) if dtype is not None else None
        partials = yacc = None
        if form is not None and hsell_config()["accumulate"]:
            yacc = torch.zeros((form.n_slices + 1) * 32, dtype=dtype, device=dev)
        elif form is not None:
            partials = torch.empty(max(form.n_partials, 1) * 32, dtype=dtype, device=dev)
        return (acc, cnt, partials, yacc)

    def span_ws(self, dtype: Optional[torch.dtype] = None):
        if self._ws is None:
            self._ws = {}
        if dtype not in self._ws:
            self._ws[dtype] = self.new_span_ws(dtype)
        return self._ws[dtype]

    def with_values(self, values64: torch.Tensor) -> "CsrView":
        other = object.__new__(CsrView)
        other.__dict__.update(self.__dict__)
        other._values = {torch.float64: values64}
        other.weighted = True
        other._ws = None
        other._vstream = {}
        other._hsell = {}
        other.transient_values = True
        return other


def span_struct(ws) -> C.SpanWs:
    return C.SpanWs(C.ptr(ws[0]), C.ptr(ws[1]), C.ptr(ws[2]) if len(ws) > 2 else None,
                    C.ptr(ws[3]) if len(ws) > 3 else None)


class IdentityNodeMap:
    """Lazy node -> index mapping for graphs whose nodes are 0..n-1 (replaces the n-entry dict
    of preprocessing.py:151, impractical at 10^8 nodes)."""

    def __init__(self, n):
        self.n = int(n)

    def __getitem__(self, k):
        k = int(k)
        if not 0 <= k < self.n:
            raise KeyError(k)
        return k

    def __iter__(self):
        return iter(range(self.n))

    def __len__(self):
        return self.n

    def __contains__(self, k):
        try:
            return 0 <= int(k) < self.n
        except (TypeError, ValueError):
            return False

    def get(self, k, default=None):
        return int(k) if k in self else default

    def keys(self):
        return range(self.n)

    def items(self):
        return ((i, i) for i in range(self.n))


_SCALE_KINDS = {
    "none": (C.SCALE_ONE, C.SCALE_ONE),
    "col": (C.SCALE_RECIP, C.SCALE_ONE),           # preprocessing.py:109-113
    "symmetric": (C.SCALE_RSQRT, C.SCALE_RSQRT),   # :131-138
    "laplacian": (C.SCALE_RSQRT, C.SCALE_RSQRT),   # :114-122 (I - symmetric)
    "both": (C.SCALE_RECIP, C.SCALE_RECIP),        # :123-130
}


class DeviceGraph:
    """A normalised propagation operator resident in HBM (the backend's graph object).

    Plays the role of ``Adjacency`` wrapping a backend matrix (preprocessing.py:9-28,145-151):
    exposes ``.array`` (itself), ``.shape``, ``_pygrank_node2id`` and the
    ``__pygrank_preprocessed`` cache entry so a second trip through ``pg.preprocessor`` short-cuts.
    """

    def __init__(self):
        self.n = 0
        self.shape = (0, 0)
        self.array = self
        self.directed = False
        self.normalization = "none"
        self.symmetric_structure = False
        self.perm = None      # internal -> user
        self.iperm = None     # user -> internal
        self.out_view: CsrView = None   # rows of A (sources), internal labels
        self.in_view: CsrView = None    # pull structure (rows = destinations)
        self.rowsum = self.colsum = None
        self.L = self.R = None          # fp64, zeros kept (S[S != 0] = 1/S)
        self.pathological = False
        self._cache = {}
        self._pygrank_node2id = None

    # ------------------------------------------------------------------ construction
    @staticmethod
    def from_edges(n: int, src: torch.Tensor, dst: torch.Tensor, weights: Optional[torch.Tensor] = None,
                   directed: bool = False, symmetrize: Optional[bool] = None, drop_self_loops: bool = False,
                   binary: bool = False, normalization: str = "auto", renormalize=False,
                   relabel: str = "hub", node2id=None) -> "DeviceGraph":
        """Edge list (device int32 tensors) -> normalised operator.  ``symmetrize`` defaults to
        ``not directed`` (an undirected edge list names every edge once)."""
        lib = C.lib()
        dev = src.device
        g = DeviceGraph()
        g.n, g.shape, g.directed = int(n), (int(n), int(n)), bool(directed)
        symmetrize = (not directed) if symmetrize is None else symmetrize
        flags = (C.BUILD_SYMMETRIZE if symmetrize else 0) | (C.BUILD_DROP_SELF_LOOPS if drop_self_loops else 0) | \
                (C.BUILD_BINARY if binary else 0)
        row = src.to(torch.int32).contiguous()
        col = dst.to(torch.int32).contiguous()
        val = None if (weights is None or binary) else weights.to(torch.float64).contiguous()
        renormalize = float(renormalize)
        indptr, indices, values = build_csr(n, row, col, val, flags)
        del row, col, val
        if renormalize != 0:                                  # preprocessing.py:107-108: M + I*renormalize
            rows = torch.empty(indices.numel(), dtype=torch.int32, device=dev)
            C.check(lib.pgb_csr_expand_rows(n, indices.numel(), C.ptr(indptr), C.ptr(rows), C.stream_ptr()))
            diag = torch.arange(n, dtype=torch.int32, device=dev)
            vals = torch.ones(indices.numel(), dtype=torch.float64, device=dev) if values is None else values
            indptr, indices, values = build_csr(
                n, torch.cat([rows, diag]), torch.cat([indices, diag]),
                torch.cat([vals, torch.full((n,), renormalize, dtype=torch.float64, device=dev)]), 0)
            del rows, diag, vals
        if values is not None and bool((values == 1.0).all()):
            values = None                                     # unweighted: stream no edge values
        g._finish(indptr, indices, values, known_symmetric=bool(symmetrize), normalization=normalization,
                  relabel=relabel)
        g._pygrank_node2id = node2id if node2id is not None else IdentityNodeMap(n)
        return g

    @staticmethod
    def from_scipy(A, directed: bool = False, normalization: str = "auto", renormalize=False,
                   relabel: str = "hub", device=None, node2id=None) -> "DeviceGraph":
        """A host scipy adjacency (what ``pg.AdjacencyWrapper`` carries, wrapgraph.py:4-22)."""
        import scipy.sparse as sp
        lib = C.lib()
        dev = _dev(device)
        A = sp.csr_matrix(A)
        n = A.shape[0]
        if A.nnz >= 2 ** 31:
            raise Exception("graphs with nnz >= 2^31 must be row-partitioned (pygrank_b200.dist)")
        indptr = torch.from_numpy(np.ascontiguousarray(A.indptr, dtype=np.int32)).to(dev)
        col = torch.from_numpy(np.ascontiguousarray(A.indices, dtype=np.int32)).to(dev)
        data = np.ascontiguousarray(A.data, dtype=np.float64)
        val = None if bool(np.all(data == 1.0)) else torch.from_numpy(data).to(dev)
        row = torch.empty(A.nnz, dtype=torch.int32, device=dev)
        C.check(lib.pgb_csr_expand_rows(n, A.nnz, C.ptr(indptr), C.ptr(row), C.stream_ptr()))
        g = DeviceGraph.from_edges(n, row, col, val, directed=directed, symmetrize=False,
                                   normalization=normalization, renormalize=renormalize, relabel=relabel,
                                   node2id=node2id)
        return g

    def _finish(self, indptr, indices, values, known_symmetric: bool, normalization, relabel: str):
        lib = C.lib()
        n, dev = self.n, indptr.device
        st = C.stream_ptr()
        nnz = int(indices.numel())
        if isinstance(normalization, str):
            normalization = normalization.lower()
        if normalization == "auto":                           # preprocessing.py:101-102
            normalization = "col" if self.directed else "symmetric"
        if normalization not in _SCALE_KINDS:
            raise Exception("Supported normalizations: none, col, symmetric, both, laplacian, auto")
        self.normalization = normalization
        if relabel not in RELABELS:
            raise Exception("relabel must be one of " + ", ".join(RELABELS))
        rows = None
        if relabel != "none" and n > 1 and nnz > 0:
            wsb = lib.pgb_degree_order_workspace_bytes(n)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            self.perm = torch.empty(n, dtype=torch.int32, device=dev)
            self.iperm = torch.empty(n, dtype=torch.int32, device=dev)
            C.check(lib.pgb_degree_order(n, C.ptr(indptr), C.ptr(ws), wsb, C.ptr(self.perm), C.ptr(self.iperm), st))
            del ws
            if relabel == "hub" and values is None:
                # refine the degree order inside every hub block / the tail by the blocks each row touches
                # (pgb_hub_order): only the hub-blocked form (unweighted graphs) gains from it
                G, span = hub_signature_shape(n)
                words = -(-(-(-span // G)) // 64)
                if 0 < words <= C.SIGNATURE_WORDS:
                    wsb = lib.pgb_hub_order_workspace_bytes(n, words)
                    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
                    C.check(lib.pgb_hub_order(n, C.ptr(indptr), C.ptr(indices), G, span, C.ptr(ws), wsb,
                                              C.ptr(self.perm), C.ptr(self.iperm), st))
                    del ws
            rows = torch.empty(nnz, dtype=torch.int32, device=dev)
            C.check(lib.pgb_csr_expand_rows(n, nnz, C.ptr(indptr), C.ptr(rows), st))
            cols = indices.clone()
            C.check(lib.pgb_relabel_coo(nnz, C.ptr(self.iperm), C.ptr(rows), C.ptr(cols), st))
            indptr, indices, values = build_csr(n, rows, cols, values, 0)
            del cols
            rows = None
        self.out_view = CsrView(n, indptr, indices, values)
        if known_symmetric:
            self.in_view = self.out_view
        else:
            rows = torch.empty(nnz, dtype=torch.int32, device=dev)
            C.check(lib.pgb_csr_expand_rows(n, nnz, C.ptr(indptr), C.ptr(rows), st))
            t_indptr, t_indices, t_values = build_csr(n, indices, rows, values, 0)
            del rows
            same = (torch.equal(t_indptr, indptr) and torch.equal(t_indices, indices) and
                    (values is None or torch.equal(t_values, values)))
            self.in_view = self.out_view if same else CsrView(n, t_indptr, t_indices, t_values)
        self.symmetric_structure = self.in_view is self.out_view
        # degrees and scales (preprocessing.py:104-138), fp64, reference rounding
        self.rowsum = torch.empty(n, dtype=torch.float64, device=dev)
        C.check(lib.pgb_csr_row_sums(n, C.ptr(self.out_view.indptr), C.ptr(self.out_view.values(torch.float64)),
                                     C.ptr(self.rowsum), st))
        if self.symmetric_structure:
            self.colsum = self.rowsum
        else:
            self.colsum = torch.empty(n, dtype=torch.float64, device=dev)
            C.check(lib.pgb_csr_row_sums(n, C.ptr(self.in_view.indptr), C.ptr(self.in_view.values(torch.float64)),
                                         C.ptr(self.colsum), st))
        lk, rk = _SCALE_KINDS[normalization]
        self.L = torch.empty(n, dtype=torch.float64, device=dev)
        C.check(lib.pgb_make_scales(n, C.ptr(self.rowsum), lk, C.ptr(self.L), st))
        if rk == lk and self.symmetric_structure:
            self.R = self.L
        else:
            self.R = torch.empty(n, dtype=torch.float64, device=dev)
            C.check(lib.pgb_make_scales(n, C.ptr(self.colsum), rk, C.ptr(self.R), st))
        # a row whose weights cancel to a zero sum keeps entries the reference multiplies by 0;
        # the scaled-domain engine cannot represent that (plain conv can)
        out_deg = self.out_view.indptr[1:] - self.out_view.indptr[:-1]
        self.pathological = bool(((self.L == 0) & (out_deg > 0)).any()) if self.out_view.weighted else False
        self.__dict__["__pygrank_preprocessed"] = {"b200": self}

    # ------------------------------------------------------------------ derived vectors
    @property
    def nnz(self) -> int:
        return self.out_view.nnz

    @property
    def symdeg(self) -> bool:
        """Scales derivable from the row pointers inside the kernel (no per-node scale streams)."""
        return (self.symmetric_structure and not self.in_view.weighted and self.normalization == "symmetric")

    def vec(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        """Cached per-dtype node vectors in internal order:
        L, R (true scales), Lp (L with 0 -> 1), w = Lp*R, sq = 1/Lp, degM = rowsum(M), c = sq*degM."""
        key = (name, dtype)
        if key in self._cache:
            return self._cache[key]
        f64 = torch.float64
        if dtype != f64:
            out = self.vec(name, f64).to(dtype)
        elif name == "L":
            out = self.L
        elif name == "R":
            out = self.R
        elif name == "Lp":
            out = torch.where(self.L == 0, torch.ones_like(self.L), self.L)
        elif name == "w":
            out = self.vec("Lp", f64) * self.R
        elif name == "sq":
            out = 1.0 / self.vec("Lp", f64)
        elif name == "degM":
            out = self._spmv_raw(self.out_view, self.R, self.L, f64)
        elif name == "c":
            out = self.vec("sq", f64) * self.vec("degM", f64)
        else:
            raise KeyError(name)
        self._cache[key] = out
        return out

    def _spmv_raw(self, view: CsrView, z: torch.Tensor, rscale: Optional[torch.Tensor], dtype, out_perm=None,
                  out=None, hsell: bool = False) -> torch.Tensor:
        """One plain gather pass.  ``hsell``: use (and if needed build) the hub-blocked form — worth it for
        repeated calls (the plugin route's conv), not for the one-off passes of the preprocessor."""
        lib = C.lib()
        if out is None:
            out = torch.empty(self.n, dtype=dtype, device=z.device)
        cs = view.cstruct(dtype, hsell=hsell)
        # (the partial-row buffer — and with it the hsell form — is only asked for when that form is used)
        C.check(lib.pgb_spmv(ctypes.byref(cs), dtype_code(dtype), C.ptr(z), C.ptr(rscale), None, C.ptr(out_perm),
                             C.ptr(out), span_struct(view.span_ws(dtype if hsell else None)), C.stream_ptr()))
        C.count_launches(view.kernels_per_step(dtype, hsell))
        return out

    # ------------------------------------------------------------------ backend operations
    def conv(self, x: torch.Tensor) -> torch.Tensor:
        """``x @ M`` (numpy.py:64-65) for a user-order vector; one pre-scale pass + one gather kernel."""
        lib = C.lib()
        dtype = x.dtype
        code = dtype_code(dtype)
        x = x.contiguous()
        if x.numel() != self.n:
            raise Exception(f"conv: vector of length {x.numel()} on a graph with {self.n} nodes")
        z = torch.empty(self.n, dtype=dtype, device=x.device)
        C.check(lib.pgb_scale(self.n, code, C.ptr(x), C.ptr(self.vec("L", dtype)), 1.0, C.ptr(self.perm), C.ptr(z),
                              C.stream_ptr()))
        rscale = None if self.normalization in ("none", "col") else self.vec("R", dtype)
        with in_kernel_dropout(getattr(self, "drop_p", 0.0)):
            y = self._spmv_raw(self.in_view, z, rscale, dtype, out_perm=self.perm, hsell=True)
        if self.normalization == "laplacian":                 # preprocessing.py:122: -M + I
            y = x - y
        return y

    def degrees(self, dtype=torch.float64) -> torch.Tensor:
        """``degrees(M)`` (numpy.py:76-77): row sums of the NORMALISED matrix, user order, summed in
        numpy's pairwise order over the stored row (bit-exact when relabel='none')."""
        lib = C.lib()
        st = C.stream_ptr()
        v = self.out_view
        dev = v.indptr.device
        data = torch.empty(max(v.nnz, 1), dtype=torch.float64, device=dev)
        C.check(lib.pgb_csr_normalized_values(self.n, C.ptr(v.indptr), C.ptr(v.indices), C.ptr(v.values(torch.float64)),
                                              C.ptr(self.L), C.ptr(self.R), C.ptr(data), st))
        if self.normalization == "laplacian":
            raise Exception("degrees() of a laplacian operator is not provided by the device preprocessor")
        sums = torch.empty(self.n, dtype=torch.float64, device=dev)
        C.check(lib.pgb_csr_row_sums_numpy(self.n, C.ptr(v.indptr), C.ptr(data), int(self.normalization == "col"),
                                           C.ptr(sums), st))
        if self.perm is not None:
            out = torch.empty_like(sums)
            out[self.perm.long()] = sums
            sums = out
        return sums.to(dtype)

    def graph_degrees(self):
        """(rowsum, colsum) of the un-normalised adjacency in user order (fp64; exact integers for
        unweighted graphs) — what preprocessing.py:104-105 reduces."""
        if self.perm is None:
            return self.rowsum, self.colsum
        p = self.iperm.long()
        return self.rowsum[p], self.colsum[p]

    def dropout(self, p: float) -> "DeviceGraph":
        """``graph_dropout`` with the torch backends' semantics (pytorch.py:34-38): every stored entry is zeroed with
        probability p and survivors are rescaled by 1/(1-p); a fresh mask per conv.  Unweighted graphs: the mask is
        drawn INSIDE the gather kernel (``pgb_hsell_set_dropout``: counter-based hash, nothing allocated or streamed);
        weighted graphs: an explicit value array on the item-stream kernels."""
        p = float(p)
        if p == 0:
            return self
        g = object.__new__(DeviceGraph)
        g.__dict__.update(self.__dict__)
        g._cache = {}
        g.array = g
        if not self.in_view.weighted and hsell_config()["enabled"] and self.nnz > 0:
            g.drop_p = p
            return g
        view = self.in_view
        base = view.values(torch.float64)
        keep = (torch.rand(view.nnz, device=view.indptr.device) >= p).to(torch.float64) / (1.0 - p)
        dropped = view.with_values(keep if base is None else base * keep)
        g.in_view = dropped                       # conv (the only consumer of a dropped graph, abstract_filters.py:59-62)
        if self.symmetric_structure:              # reads the pull structure; a directed graph keeps its push view
            g.out_view = dropped
        return g

    def to_scipy_normalized(self):
        """Normalised matrix as canonical host scipy CSR in USER labels (parity checks; not hot)."""
        import scipy.sparse as sp
        lib = C.lib()
        v = self.out_view
        data = torch.empty(max(v.nnz, 1), dtype=torch.float64, device=v.indptr.device)
        C.check(lib.pgb_csr_normalized_values(self.n, C.ptr(v.indptr), C.ptr(v.indices), C.ptr(v.values(torch.float64)),
                                              C.ptr(self.L), C.ptr(self.R), C.ptr(data), C.stream_ptr()))
        indptr = v.indptr.cpu().numpy()
        indices = v.indices.cpu().numpy()
        data = data[:v.nnz].cpu().numpy()
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(indptr))
        cols = indices.astype(np.int64)
        if self.perm is not None:
            perm = self.perm.cpu().numpy().astype(np.int64)
            rows, cols = perm[rows], perm[cols]
        order = np.lexsort((cols, rows))
        rows, cols, data = rows[order], cols[order], data[order]
        out_indptr = np.zeros(self.n + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows, minlength=self.n), out=out_indptr[1:])
        M = sp.csr_matrix((data, cols.astype(np.int32), out_indptr), shape=self.shape)
        if self.normalization == "laplacian":
            M = (-M + sp.eye(self.n, format="csr")).tocsr()
            M.sort_indices()
        return M

    def __len__(self):
        return self.n

    def __iter__(self):
        return iter(range(self.n))

    def is_directed(self):
        return self.directed


def as_device_graph(graph, normalization="auto", renormalize=False, relabel="hub", weight="weight",
                    device=None) -> DeviceGraph:
    """Anything the reference's preprocessor accepts -> DeviceGraph (preprocessing.py:88-103)."""
    if isinstance(graph, DeviceGraph):
        return graph
    if hasattr(graph, "array") and isinstance(getattr(graph, "array"), DeviceGraph):
        return graph.array
    import scipy.sparse as sp
    if sp.issparse(graph):
        return DeviceGraph.from_scipy(graph, directed=False, normalization=normalization, renormalize=renormalize,
                                      relabel=relabel, device=device)
    if hasattr(graph, "edge_row") and hasattr(graph, "edge_col") and hasattr(graph, "node_map"):
        from .ingest import from_fastgraph                   # fastgraph.Graph: its edge lists go to the device as COO
        return from_fastgraph(graph, normalization=normalization, renormalize=renormalize, relabel=relabel, device=device)
    if hasattr(graph, "to_scipy_sparse_array"):               # AdjacencyWrapper
        return DeviceGraph.from_scipy(graph.to_scipy_sparse_array(), directed=bool(graph.is_directed()),
                                      normalization=normalization, renormalize=renormalize, relabel=relabel,
                                      device=device, node2id={v: i for i, v in enumerate(graph)}
                                      if not _is_range_nodes(graph) else None)
    try:
        import networkx as nx
    except ImportError:  # pragma: no cover
        nx = None
    if nx is not None and isinstance(graph, nx.Graph):
        A = nx.to_scipy_sparse_array(graph, weight=weight, dtype=float)   # preprocessing.py:103
        return DeviceGraph.from_scipy(A, directed=graph.is_directed(), normalization=normalization,
                                      renormalize=renormalize, relabel=relabel, device=device,
                                      node2id={v: i for i, v in enumerate(graph)})
    raise Exception("cannot build a device graph from " + str(type(graph)))


def _is_range_nodes(graph) -> bool:
    it = iter(graph)
    return isinstance(it, type(iter(range(0))))


def preprocessor(normalization: str = "auto", assume_immutability: bool = False, weight: str = "weight",
                 renormalize=False, relabel: str = "hub", device=None):
    """Device twin of ``pg.preprocessor`` (preprocessing.py:233-287): returns a callable named
    ``preprocess`` (so ``filter + preprocess`` works, abstract_filters.py:91-92) mapping a graph
    to a :class:`DeviceGraph`; ``assume_immutability`` memoises per input object like MethodHasher."""
    memo = {}

    def preprocess(G):
        if isinstance(G, DeviceGraph):
            return G
        key = id(G)
        if assume_immutability and key in memo and memo[key][0] is G:
            return memo[key][1]
        out = as_device_graph(G, normalization=normalization, renormalize=renormalize, relabel=relabel,
                              weight=weight, device=device)
        if assume_immutability:
            memo[key] = (G, out)
        return out

    return preprocess
