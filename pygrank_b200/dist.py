"""Row-partitioned propagation across the GPUs of one box (one process per GPU, torch.distributed).

The reference has no multi-device path at all; this is the new capability the
north_star asks for: graphs larger than one GPU are split by destination rows, every rank runs the
same fused kernel (csrc/spmv_fused.cu) on its rows against the FULL pre-scaled rank vector, and the
only exchange per iteration is an NCCL all-gather of the rank-vector slices over NVLink plus a
16-byte all-reduce of the two convergence accumulators, after which every rank plays the same
device-side ConvergenceManager (``pgb_state_finalize``).

Partition.  Nodes are ranked by (approximate) degree and dealt boustrophedon to the ranks, and the
engine's internal node id is ``rank * n_local + position``: every rank owns a CONTIGUOUS id range
(so the all-gather output is the full vector in id order, no re-indexing), the ranges have equal
length (``all_gather_into_tensor``), and because neighbours in degree order have near-equal degree
the nnz per rank is balanced to a fraction of a percent even on power-law graphs.  Each range still
starts with its hubs, which keeps the hot end of the gather vector in a few cache lines.

The helpers that build the partition are plain torch code (device agnostic) so the N>1 host logic
is covered by world_size-2 gloo tests on CPU; the compute step is the CUDA library only.
"""
from __future__ import annotations

import ctypes
import math
import os
import time
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _capi as C
from .synthetic import rmat_thresholds


# ------------------------------------------------------------------------------------------------
# partition arithmetic (device agnostic)
def padded_size(n: int, world: int) -> int:
    return ((n + world - 1) // world) * world


def interleaved_ids(degree: torch.Tensor, world: int) -> torch.Tensor:
    """new_id[v] for every node v: the nodes in descending-degree order (stable) are dealt to the
    ranks boustrophedon (0..P-1, P-1..0, ...), the k-th one landing at position ``k // world`` of its
    rank, which balances the degree sums; ``len(degree)`` must be a multiple of ``world``."""
    n = degree.numel()
    assert n % world == 0
    n_local = n // world
    order = torch.sort(degree, descending=True, stable=True).indices          # order[k] = node
    k = torch.arange(n, device=degree.device, dtype=torch.int64)
    rr = k % (2 * world)
    owner = torch.where(rr < world, rr, 2 * world - 1 - rr)
    new_of_k = owner * n_local + (k // world)
    new_id = torch.empty(n, dtype=torch.int32, device=degree.device)
    new_id[order] = new_of_k.to(torch.int32)
    return new_id


def local_entries(src: torch.Tensor, dst: torch.Tensor, new_id: torch.Tensor, rank: int, n_local: int):
    """Symmetrised pull entries owned by ``rank`` from a chunk of raw edges: for an undirected edge
    {u, v} the entry (row u, col v) lives with u's owner and (row v, col u) with v's owner; self loops
    are dropped.  Returns (local_row int32, global_col int32)."""
    u = new_id[src.long()]
    v = new_id[dst.long()]
    lo, hi = rank * n_local, (rank + 1) * n_local
    keep = u != v
    a = keep & (u >= lo) & (u < hi)
    b = keep & (v >= lo) & (v < hi)
    rows = torch.cat([u[a] - lo, v[b] - lo])
    cols = torch.cat([v[a], u[b]])
    return rows.to(torch.int32), cols.to(torch.int32)


def virtual_columns(cols: torch.Tensor, n_local: int, world: int, block_cols: int, n_blocks: int) -> torch.Tensor:
    """Global column id (``rank * n_local + position``) -> the virtual column the hsell builders expect
    (include/pgb200.h, pgb_hsell with n_segments > 1): hub block b gathers positions
    ``[b*Hs, (b+1)*Hs)`` of EVERY rank's range (Hs = block_cols / world), laid out rank-major inside the
    block; positions past the hub blocks form the tail, rank-major.  A bijection of ``[0, world*n_local)``;
    ``real_columns`` (and hsell_real_col in csrc/hsell.cu) is its inverse."""
    cols = cols.long()
    hs = block_cols // world
    rnk = cols // n_local
    pos = cols - rnk * n_local
    blk = pos // hs
    hub = blk < n_blocks
    v_hub = blk * block_cols + rnk * hs + (pos - blk * hs)
    tl = n_local - n_blocks * hs
    v_tail = n_blocks * block_cols + rnk * tl + (pos - n_blocks * hs)
    return torch.where(hub, v_hub, v_tail)


def real_columns(vcols: torch.Tensor, n_local: int, world: int, block_cols: int, n_blocks: int) -> torch.Tensor:
    v = vcols.long()
    hs = block_cols // world
    span = n_blocks * block_cols
    blk = v // block_cols
    local = v - blk * block_cols
    rnk_h = local // hs
    hub_col = rnk_h * n_local + blk * hs + (local - rnk_h * hs)
    tl = max(n_local - n_blocks * hs, 1)
    vt = v - span
    rnk_t = vt // tl
    tail_col = rnk_t * n_local + n_blocks * hs + (vt - rnk_t * tl)
    return torch.where(v < span, hub_col, tail_col)


def reader_mask_from_need(need: torch.Tensor, rank: int, world: int, n_local: int, group=None) -> torch.Tensor:
    """``need`` uint8[world*n_local]: 1 where THIS rank reads the entry of the gather vector (its rows reference
    the column, it lies in one of its hub blocks, or it owns it).  One all-gather of these byte maps gives
    every rank, for each of its own rows, the set of ranks that read the row's value: bit r of the returned
    int32[n_local].  Collective; device agnostic (covered by the gloo CPU test)."""
    gathered = torch.empty(world * need.numel(), dtype=torch.uint8, device=need.device)
    dist.all_gather_into_tensor(gathered, need.contiguous(), group=group)
    mine = gathered.view(world, need.numel())[:, rank * n_local:(rank + 1) * n_local]
    mask = torch.zeros(n_local, dtype=torch.int32, device=need.device)
    for r in range(world):
        mask |= mine[r].to(torch.int32) << r
    return mask


# ------------------------------------------------------------------------------------------------
class DistGraph:
    """This rank's rows of a symmetric-normalised, unweighted, undirected graph."""

    def __init__(self):
        self.rank = self.world = 0
        self.n_global = self.n_local = self.n_nodes = 0
        self.nnz_local = self.nnz_global = 0
        self.view = None            # CsrView: local rows x global columns
        self.new_id = None          # int32[n_nodes]: user node -> internal id
        self.group = None
        self._cache = {}

    @property
    def offset(self) -> int:
        return self.rank * self.n_local

    @staticmethod
    def rmat(scale: int, edge_factor: int = 16, seed: int = 1, a=0.57, b=0.19, c=0.19, group=None,
             chunk_edges: int = 1 << 26, device=None) -> "DistGraph":
        """RMAT graph of the bench recipe, generated on the fly: every rank regenerates the
        (counter-based, deterministic) edge stream, first to count degrees, then to keep its rows."""
        from .graph import CsrView, build_csr
        lib = C.lib()
        g = DistGraph()
        g.group = group
        g.rank, g.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        n = 1 << scale
        g.n_nodes = n
        g.n_global = padded_size(n, g.world)
        g.n_local = g.n_global // g.world
        total = edge_factor << scale
        t1, t2, t3 = rmat_thresholds(a, b, c)
        st = C.stream_ptr()

        def edges(first, count):
            src = torch.empty(count, dtype=torch.int32, device=dev)
            dst = torch.empty(count, dtype=torch.int32, device=dev)
            C.check(lib.pgb_rmat_edges(scale, first, count, seed, t1, t2, t3, C.ptr(src), C.ptr(dst), st))
            return src, dst

        # pass 1: raw degree counts (duplicates included — only a ranking heuristic), each rank a slice
        deg = torch.zeros(g.n_global, dtype=torch.int32, device=dev)
        ones = None
        per_rank = (total + g.world - 1) // g.world
        lo, hi = g.rank * per_rank, min((g.rank + 1) * per_rank, total)
        for first in range(lo, hi, chunk_edges):
            count = min(chunk_edges, hi - first)
            src, dst = edges(first, count)
            if ones is None or ones.numel() != count:
                ones = torch.ones(count, dtype=torch.int32, device=dev)
            deg.index_add_(0, src.long(), ones)
            deg.index_add_(0, dst.long(), ones)
        del ones
        dist.all_reduce(deg, group=group)
        g.new_id = interleaved_ids(deg, g.world)
        del deg
        # pass 2: keep the entries whose destination row lives here
        rows, cols = [], []
        for first in range(0, total, chunk_edges):
            count = min(chunk_edges, total - first)
            src, dst = edges(first, count)
            r, cc = local_entries(src, dst, g.new_id, g.rank, g.n_local)
            rows.append(r)
            cols.append(cc)
            del src, dst
        rows, cols = torch.cat(rows), torch.cat(cols)
        # (row, col) pairs -> canonical CSR; the column space is the global id range, so build with
        # n = n_global and keep the first n_local+1 row pointers (rows beyond are empty)
        indptr, indices, _ = build_csr(g.n_global, rows, cols, None, C.BUILD_BINARY)
        del rows, cols
        indptr = indptr[: g.n_local + 1].clone()
        g.view = CsrView(g.n_local, indptr, indices, None)
        g.nnz_local = g.view.nnz
        tot = torch.tensor([g.nnz_local], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=group)
        g.nnz_global = int(tot.item())
        return g

    def hsell(self, dtype: torch.dtype):
        """Hub-blocked sliced-ELL form of this rank's rows (csrc/hsell.cu) against the all-gathered
        vector: built once per dtype from a copy of the CSR whose columns are relabelled so that every hub
        block is one contiguous column range (``virtual_columns``); the copy is dropped afterwards."""
        from .graph import CsrView, HsellForm, build_csr, hsell_config, hsell_shape
        if dtype in self.view._hsell:
            return self.view._hsell[dtype]
        if not hsell_config()["enabled"] or self.view.nnz == 0:
            return None
        lib = C.lib()
        H, K = hsell_shape(dtype, self.world, self.n_local)
        dev = self.view.indptr.device
        rows = torch.empty(self.view.nnz, dtype=torch.int32, device=dev)
        C.check(lib.pgb_csr_expand_rows(self.n_local, self.view.nnz, C.ptr(self.view.indptr), C.ptr(rows), C.stream_ptr()))
        vcols = virtual_columns(self.view.indices, self.n_local, self.world, H, K).to(torch.int32)
        indptr, indices, _ = build_csr(self.n_global, rows, vcols, None, 0)
        del rows, vcols
        tmp = CsrView(self.n_local, indptr[: self.n_local + 1].clone(), indices, None)
        tmp.n_cols = self.n_global
        form = HsellForm(tmp, dtype, self.world, self.n_local, cfg={"block_cols": H, "max_blocks": K})
        del tmp
        self.view._hsell[dtype] = form
        self.view.n_cols = self.n_global
        return form

    def reader_mask(self, dtype: torch.dtype) -> torch.Tensor:
        """int32[n_local]: bit r set when rank r reads the value of this local row — r's rows reference the
        column, or the column lies in one of r's hub blocks (loaded wholesale), or r owns it.  The update
        kernel stores a new value only into those ranks' buffers (``pgb_peers.row_mask``)."""
        key = ("reader_mask", dtype)
        if key in self._cache:
            return self._cache[key]
        dev = self.view.indptr.device
        need = torch.zeros(self.n_global, dtype=torch.uint8, device=dev)
        chunk = 1 << 27
        for a in range(0, self.view.nnz, chunk):                      # columns my rows gather
            need[self.view.indices[a:a + chunk].long()] = 1
        form = self.hsell(dtype)
        if form is not None and form.n_blocks > 0:                    # hub blocks are copied to shared memory whole
            hs = form.block_cols // self.world
            need.view(self.world, self.n_local)[:, : min(form.n_blocks * hs, self.n_local)] = 1
        need[self.offset:self.offset + self.n_local] = 1              # the owner reads its own rows
        mask = reader_mask_from_need(need, self.rank, self.world, self.n_local, self.group)
        del need
        self._cache[key] = mask
        return mask

    def peer_buffers(self, dtype: torch.dtype):
        """Peer-mapped buffers for the exchange fused into the step (``pgb_affine_step_peer``): both
        gather-vector buffers and the convergence-sum slots in torch symmetric memory, so that every rank's
        update kernel writes its slice straight into all ranks (NVSwitch multicast when available).
        Returns None when symmetric memory cannot be set up (then the NCCL all-gather path runs);
        ``PGB_PEER=0`` disables it.  Per-peer NVLink stores are the default: measured on 2 and 8 B200 they are
        as fast as or faster than the multicast path (1221 vs 1153 GTEPS at N=2, 1865 vs 1835 at N=8), which
        ``PGB_PEER_MULTICAST=1`` selects."""
        import os
        key = ("peer", dtype)
        if key in self._cache:
            return self._cache[key]
        out = None
        if os.environ.get("PGB_PEER", "1") != "0" and self.world > 1 and self.world <= C.MAX_PEERS:
            try:
                import torch.distributed._symmetric_memory as symm
                dev = self.view.indptr.device
                group = self.group if self.group is not None else dist.group.WORLD
                w = 4 if dtype == torch.float32 else 8
                zsym = symm.empty(2 * self.n_global, dtype=dtype, device=dev)
                hz = symm.rendezvous(zsym, group)
                asym = symm.empty(2 * 2 * C.MAX_PEERS, dtype=torch.float64, device=dev)
                ha = symm.rendezvous(asym, group)
                asym.zero_()
                mc = int(hz.multicast_ptr) if os.environ.get("PGB_PEER_MULTICAST", "0") == "1" else 0
                mask = None
                # measured: at 8 ranks the masks cut the stores to a fraction and lift 2131 -> 2470 GTEPS; at 2 ranks
                # almost every row is read by the other rank and the mask only costs (1231 -> 1158)
                # PGB_PEER_MASK: unset = masks from 4 ranks up, 1 = always (parity tests exercise the mask path
                # on 2 GPUs), 0 = never
                want_mask = os.environ.get("PGB_PEER_MASK", "")
                if not mc and (want_mask == "1" or (want_mask != "0" and self.world >= 4)):
                    mask = self.reader_mask(dtype)
                peers = []
                for parity in (0, 1):
                    ps = C.Peers()
                    ps.n, ps.rank = self.world, self.rank
                    for r in range(self.world):
                        ps.zbuf0[r] = int(hz.buffer_ptrs[r])
                        ps.zbuf1[r] = int(hz.buffer_ptrs[r]) + self.n_global * w
                        ps.acc[r] = int(ha.buffer_ptrs[r]) + parity * 2 * C.MAX_PEERS * 8
                    ps.mc_zbuf0 = mc if mc else None
                    ps.mc_zbuf1 = (mc + self.n_global * w) if mc else None
                    ps.row_mask = C.ptr(mask)
                    peers.append(ps)
                sent = None
                if mask is not None:
                    bits = sum(((mask >> r) & 1).sum() for r in range(self.world))
                    sent = float(bits) / float(self.world * self.n_local)     # fraction of (row, rank) pairs stored
                out = {"z": zsym, "hz": hz, "acc": asym, "ha": ha, "peers": peers, "multicast": bool(mc),
                       "mask": mask, "sent_fraction": sent}
            except Exception as exc:   # no symmetric memory on this system: the NCCL path is the product there
                self._peer_error = repr(exc)
                out = None
        self._cache[key] = out
        return out

    # per-dtype node vectors for the local rows ------------------------------------------------
    def vec(self, name: str, dtype: torch.dtype) -> torch.Tensor:
        from .graph import dtype_code, span_struct
        key = (name, dtype)
        if key in self._cache:
            return self._cache[key]
        f64 = torch.float64
        if dtype != f64:
            out = self.vec(name, f64).to(dtype)
        elif name == "deg":
            out = (self.view.indptr[1:] - self.view.indptr[:-1]).to(f64)
        elif name == "sq":                                   # 1/L with L = 1/sqrt(deg), deg 0 -> 1
            d = self.vec("deg", f64)
            out = torch.where(d > 0, torch.sqrt(d), torch.ones_like(d))
        elif name == "R":                                    # 1/sqrt(deg), zeros kept (preprocessing.py:133-136)
            d = self.vec("deg", f64)
            out = torch.where(d > 0, 1.0 / torch.sqrt(d), torch.zeros_like(d))
        elif name == "degM":                                 # row sums of the normalised matrix
            lib = C.lib()
            r_full = torch.empty(self.n_global, dtype=f64, device=self.view.indptr.device)
            dist.all_gather_into_tensor(r_full, self.vec("R", f64).contiguous(), group=self.group)
            out = torch.empty(self.n_local, dtype=f64, device=r_full.device)
            cs = self.view.cstruct(f64, hsell=False)
            C.check(lib.pgb_spmv(ctypes.byref(cs), dtype_code(f64), C.ptr(r_full), C.ptr(self.vec("R", f64)), None,
                                 None, C.ptr(out), span_struct(self.view.span_ws()), C.stream_ptr()))
        elif name == "c":
            out = self.vec("sq", f64) * self.vec("degM", f64)
        else:
            raise KeyError(name)
        self._cache[key] = out
        return out


class DistPageRank:
    """``pg.PageRank`` semantics (adhoc.py:34-36 + abstract_filters.py:44-65,126-136 +
    convergence.py:77-101) on a :class:`DistGraph`.  Every rank calls ``rank`` collectively."""

    def __init__(self, alpha: float = 0.85, tol: Optional[float] = 1e-6, max_iters: int = 100, end_modulo: int = 1,
                 error_type: str = "mabs", use_quotient: bool = True, dtype: torch.dtype = torch.float32,
                 chunk: int = 4):
        self.alpha, self.tol, self.max_iters, self.end_modulo = alpha, tol, int(max_iters), int(end_modulo)
        self.error_type, self.use_quotient, self.dtype, self.chunk = error_type, use_quotient, dtype, int(chunk)
        self.poison = os.environ.get("PGB_PEER_POISON", "0") == "1"   # NaN-fill the exchanged buffers before a solve
        self.iteration = 0
        self.elapsed_time = None

    def local_personalization(self, g: DistGraph, seeds, values=None):
        """Host seed list -> (this rank's slice of the personalization vector, its L1 norm)."""
        dev = g.view.indptr.device
        seeds = np.asarray(seeds, dtype=np.int64)
        vals = np.ones(len(seeds)) if values is None else np.asarray(values, dtype=np.float64)
        norm = float(np.abs(vals).sum())
        p = torch.zeros(g.n_local, dtype=self.dtype, device=dev)
        if len(seeds):
            ids = g.new_id[torch.from_numpy(seeds).to(dev)].long()
            mine = (ids >= g.offset) & (ids < g.offset + g.n_local)
            p[ids[mine] - g.offset] = torch.from_numpy(vals).to(device=dev, dtype=self.dtype)[mine]
        return p, norm

    def rank(self, g: DistGraph, seeds=None, values=None, p_local=None, norm=None) -> torch.Tensor:
        """``seeds``: user node ids (host array) with optional ``values`` (default 1), or a prebuilt
        ``p_local``/``norm`` pair from :meth:`local_personalization`.  Returns this rank's slice of
        the scores (internal id order, ``g.offset`` onwards)."""
        from .filters import _error_code
        from .graph import dtype_code, span_struct
        lib = C.lib()
        dtype, code = self.dtype, dtype_code(self.dtype)
        dev = g.view.indptr.device
        st = C.stream_ptr()
        t0 = time.perf_counter()
        n_loc, off = g.n_local, g.offset
        timing = os.environ.get("PGB_DIST_TIMING", "0") == "1"     # phase breakdown of one solve (CUDA events)
        marks = []

        def mark(name):
            if timing:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append((name, ev, time.perf_counter()))

        mark("start")
        if p_local is None:
            p, norm = self.local_personalization(g, seeds, values)
        else:
            p = p_local
        if norm == 0:
            self.iteration = 0
            return p

        err_code = _error_code(self.error_type)
        if err_code == C.ERR_MAX and g.peer_buffers(dtype) is None:
            raise Exception("MaxDifference on the row-partitioned path needs the symmetric-memory exchange")
        sf = [0.0] * C.STATE_LEN
        si = [0] * C.STATE_LEN
        sf[C.SF_ALPHA], sf[C.SF_INVS] = float(self.alpha), 1.0
        sf[C.SF_TOL] = 0.0 if self.tol is None else max(float(self.tol), float(np.finfo(float).eps))
        sf[C.SF_MEAN] = 1.0 if err_code in (C.ERR_L1, C.ERR_MAX) else float(g.n_nodes)   # Mabs divides by the node count
        sf[C.SF_NORM] = norm
        si[C.SI_MAX_ITERS], si[C.SI_END_MODULO] = self.max_iters, max(self.end_modulo, 1)
        si[C.SI_ERR_MODE], si[C.SI_QUOTIENT] = err_code, int(bool(self.use_quotient))
        state_f64 = torch.tensor(sf, dtype=torch.float64, device=dev)
        state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
        err_hist = torch.zeros(self.max_iters + 2, dtype=torch.float64, device=dev)

        sq, cvec = g.vec("sq", dtype), g.vec("c", dtype)
        peer = g.peer_buffers(dtype)
        if peer is not None:
            zfull = [peer["z"][:g.n_global], peer["z"][g.n_global:]]
        else:
            zfull = [torch.empty(g.n_global, dtype=dtype, device=dev), torch.empty(g.n_global, dtype=dtype, device=dev)]
        q = torch.empty(n_loc, dtype=dtype, device=dev)
        if peer is not None and self.poison:
            # parity runs: every entry a peer fails to deliver (reader masks leave unread entries alone) is a NaN
            peer["z"].fill_(float("nan"))
            peer["hz"].barrier(channel=1)
        if peer is not None:
            # start vector straight into every rank's buffer 0 (no all-gather); the barrier orders it before step 1
            C.check(lib.pgb_affine_init_peer(n_loc, code, C.ptr(p), None, C.ptr(sq), C.ptr(cvec), 1 - self.alpha, None,
                                             None, off, C.ptr(q), C.ptr(state_f64), ctypes.byref(peer["peers"][0]), st))
            peer["hz"].barrier(channel=2)
        else:
            C.check(lib.pgb_affine_init(n_loc, code, C.ptr(p), None, C.ptr(sq), C.ptr(cvec), 1 - self.alpha, None, None,
                                        off, C.ptr(zfull[0]), C.ptr(q), C.ptr(state_f64), st))
            dist.all_gather_into_tensor(zfull[0], zfull[0][off:off + n_loc], group=g.group)
        # BIAS (slot 1) and TACC (slot 3) in one call; INVS (slot 2) between them is rewritten by init_finish
        dist.all_reduce(state_f64[C.SF_BIAS:C.SF_TACC + 1], group=g.group)
        C.check(lib.pgb_affine_init_finish(C.ptr(state_f64), C.ptr(state_i32), st))
        C.count_launches(2)
        form = g.hsell(dtype)
        cs = g.view.cstruct(dtype, hsell=form is not None)
        ws = g.view.new_span_ws(dtype if form is not None else None)
        kernels_per_step = g.view.kernels_per_step(dtype, form is not None)
        acc = state_f64[C.SF_TACC:C.SF_EACC + 1]

        mark("init")
        budget, done = self.max_iters - 1, 0
        stop, steps, iteration = C.RUNNING, 0, 1
        # run-ahead chunk: a previous solve on this graph is the best guess of how many steps this one needs,
        # so repeated solves read the device state back once
        chunk = max(self.chunk, min(getattr(self, "_steps_hint", 0) + 1, 64), 1)
        while done < budget:
            count = min(chunk, budget - done)
            for j in range(count):
                k = done + 1 + j
                if peer is not None:
                    # exchange fused into the step: the update kernel writes z' and the convergence sums into
                    # every rank; a symmetric-memory barrier (one small kernel) replaces all-gather + all-reduce
                    C.check(lib.pgb_affine_step_peer(ctypes.byref(cs), code, float(self.alpha), None, None,
                                                     C.ptr(cvec), C.ptr(q), C.ptr(zfull[0]), C.ptr(zfull[1]), off,
                                                     C.ptr(state_f64), C.ptr(state_i32), C.ptr(err_hist),
                                                     span_struct(ws), k, ctypes.byref(peer["peers"][k & 1]), st))
                    peer["hz"].barrier(channel=0)
                    C.check(lib.pgb_state_finalize_peer(C.ptr(state_f64), C.ptr(state_i32), C.ptr(err_hist),
                                                        peer["acc"].data_ptr() + (k & 1) * 2 * C.MAX_PEERS * 8,
                                                        g.world, st))
                    continue
                C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, float(self.alpha), None, None, C.ptr(cvec),
                                             C.ptr(q), C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64),
                                             C.ptr(state_i32), C.ptr(err_hist), span_struct(ws), k, 1, 0, st))
                out = zfull[k & 1]
                dist.all_gather_into_tensor(out, out[off:off + n_loc], group=g.group)
                dist.all_reduce(acc, group=g.group)
                C.check(lib.pgb_state_finalize(C.ptr(state_f64), C.ptr(state_i32), C.ptr(err_hist), st))
            C.count_launches((kernels_per_step + 1) * count)
            done += count
            host = state_i32.cpu()
            stop, steps, iteration = int(host[C.SI_STOP]), int(host[C.SI_STEPS]), int(host[C.SI_ITERATION])
            if stop != C.RUNNING:
                break
            chunk = min(chunk * 2, 32)
        if stop == C.RUNNING:
            iteration, stop = 1, C.MAX_ITERS
        self.iteration = iteration
        self._steps_hint = steps
        self.errors = err_hist[1:steps + 1]
        if stop == C.MAX_ITERS and err_code != C.ERR_ITERS:
            raise Exception("Could not converge within " + str(self.max_iters) + " iterations")
        mark("loop")
        if peer is not None:
            peer["hz"].barrier(channel=1)   # nobody starts the next solve (overwriting buffer 0) before all have read
        result = torch.empty(n_loc, dtype=dtype, device=dev)
        zl = zfull[steps & 1][off:off + n_loc]
        C.check(lib.pgb_unscale(n_loc, code, C.ptr(zl.contiguous()), C.ptr(sq), None, norm, None, C.ptr(result), st))
        C.count_launches(1)
        mark("end")
        if timing:
            torch.cuda.synchronize()
            self.timing = {"steps": steps, "launched": done,
                           "gpu_ms": {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])},
                           "host_ms": {b[0]: (b[2] - a[2]) * 1e3 for a, b in zip(marks, marks[1:])}}
        self.elapsed_time = time.perf_counter() - t0
        return result

    def gather_user_order(self, g: DistGraph, local_scores: torch.Tensor) -> torch.Tensor:
        """All ranks: the full score vector in USER node order (testing / small graphs)."""
        full = torch.empty(g.n_global, dtype=local_scores.dtype, device=local_scores.device)
        dist.all_gather_into_tensor(full, local_scores.contiguous(), group=g.group)
        return full[g.new_id.long()][: g.n_nodes]


# ------------------------------------------------------------------------------------------------
# independent-unit sharding: seed sets / alpha values / tuner candidates are independent,
# so with the graph replicated on every GPU the columns of a propagate() call are dealt to the ranks and
# no collective touches the data path; only the results are gathered, if asked for.
def column_shard(n_columns: int, rank: int, world: int) -> range:
    """Columns owned by ``rank``: contiguous, sizes differing by at most one."""
    base, extra = divmod(n_columns, world)
    lo = rank * base + min(rank, extra)
    return range(lo, lo + base + (1 if rank < extra else 0))


def propagate_sharded(alg, graph, features: torch.Tensor, gather: bool = True, group=None):
    """``NodeRanking.propagate`` (core/signals.py:225-226) over the GPUs of one box: every rank runs
    ``alg.propagate`` (a pygrank_b200 filter) on its own columns of ``features`` against its own replica of
    ``graph``.  Returns ``(scores, iterations)``: with ``gather`` the full ``[n, B]`` matrix and the
    per-column iteration counts on every rank, else this rank's ``[n, len(shard)]`` block.  Collective."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = int(features.shape[1])
    mine = column_shard(B, rank, world)
    if len(mine):
        local = alg.propagate(graph, features[:, mine.start:mine.stop].contiguous())
    else:   # an empty shard still takes part in the gather: same dtype and device as every other rank's block
        view = getattr(graph, "out_view", None)
        local = torch.zeros((features.shape[0], 0), dtype=getattr(alg, "dtype", features.dtype),
                            device=view.indptr.device if view is not None else features.device)
    its = list(getattr(alg.convergence, "iterations", [])) if len(mine) else []
    if not gather:
        return local, its
    width = -(-B // world)                                           # pad every block to the widest shard
    block = local.new_zeros((local.shape[0], width))
    block[:, :local.shape[1]] = local
    blocks = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(blocks, block, group=group)
    all_its = [None] * world
    dist.all_gather_object(all_its, its, group=group)
    out = torch.cat([blocks[r][:, :len(column_shard(B, r, world))] for r in range(world)], dim=1)
    return out, [i for part in all_its for i in part]
