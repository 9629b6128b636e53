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
        self.view = None            # CsrView: local rows x global columns (pull structure: sources of every local node)
        self.out_view = None        # directed graphs: CsrView of the targets of every local node (None: same as view)
        self.new_id = None          # int32[n_nodes]: user node -> internal id
        self.group = None
        self.directed = False
        self.normalization = "symmetric"
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

    @staticmethod
    def from_edges(n: int, src: torch.Tensor, dst: torch.Tensor, directed: bool = False, normalization: str = "auto",
                   group=None) -> "DistGraph":
        """An arbitrary unweighted graph, row-partitioned: every rank passes the SAME edge list (device tensors,
        user node ids; an undirected edge named once) and keeps the pull rows of the nodes it owns — for a directed
        graph also their push rows, which the row sums of the normalised operator need.  Normalisations as in
        preprocessing.py:101-138 ("auto": col when directed, else symmetric), kept factorised like DeviceGraph."""
        from .graph import CsrView, build_csr, _SCALE_KINDS
        g = DistGraph()
        g.group = group
        g.rank, g.world = dist.get_rank(group), dist.get_world_size(group)
        dev = src.device
        g.n_nodes = int(n)
        g.n_global = padded_size(g.n_nodes, g.world)
        g.n_local = g.n_global // g.world
        g.directed = bool(directed)
        normalization = normalization.lower()
        if normalization == "auto":
            normalization = "col" if directed else "symmetric"
        if normalization not in _SCALE_KINDS or normalization == "laplacian":
            raise Exception("row-partitioned graphs support the normalizations none, col, symmetric, both, auto")
        g.normalization = normalization
        s64, d64 = src.long(), dst.long()
        if not directed:                                  # both directions, self loops once
            loops = s64 == d64
            s64, d64 = torch.cat([s64, d64[~loops]]), torch.cat([d64, s64[~loops]])
        # canonical entry set (duplicates collapse: unweighted)
        key = torch.unique(s64 * g.n_nodes + d64)
        s64, d64 = key // g.n_nodes, key % g.n_nodes
        del key
        rowsum = torch.bincount(s64, minlength=g.n_global)          # out-degree (row sums of the adjacency)
        colsum = torch.bincount(d64, minlength=g.n_global)          # in-degree
        full_id = interleaved_ids((rowsum + colsum).to(torch.int32), g.world)
        g.new_id = full_id[: g.n_nodes].contiguous()
        lo = g.rank * g.n_local
        u, v = full_id[s64].long(), full_id[d64].long()              # entry a_uv: u -> v
        mine = (v >= lo) & (v < lo + g.n_local)                      # pull row of v lists its sources u
        indptr, indices, _ = build_csr(g.n_global, (v[mine] - lo).to(torch.int32), u[mine].to(torch.int32), None,
                                       C.BUILD_BINARY)
        g.view = CsrView(g.n_local, indptr[: g.n_local + 1].clone(), indices, None)
        if directed:
            own = (u >= lo) & (u < lo + g.n_local)                   # push row of u lists its targets v
            ip, ix, _ = build_csr(g.n_global, (u[own] - lo).to(torch.int32), v[own].to(torch.int32), None, C.BUILD_BINARY)
            g.out_view = CsrView(g.n_local, ip[: g.n_local + 1].clone(), ix, None)
        inv = torch.empty(g.n_global, dtype=torch.int64, device=dev)
        inv[full_id.long()] = torch.arange(g.n_global, device=dev)
        mine_nodes = inv[lo: lo + g.n_local]                          # padded-user id of every local row
        g._cache[("rowsum", torch.float64)] = rowsum[mine_nodes].to(torch.float64)
        g._cache[("colsum", torch.float64)] = colsum[mine_nodes].to(torch.float64)
        g.user_of_local = mine_nodes
        g.nnz_local = g.view.nnz
        tot = torch.tensor([g.nnz_local], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=group)
        g.nnz_global = int(tot.item())
        return g

    @staticmethod
    def from_scipy(A, directed: bool = False, normalization: str = "auto", group=None, device=None) -> "DistGraph":
        """A host scipy adjacency every rank holds (what pg.AdjacencyWrapper carries); must be unweighted."""
        import scipy.sparse as sp
        A = sp.coo_matrix(A)
        if not bool(np.all(A.data == 1.0)):
            raise Exception("row-partitioned graphs are unweighted (weighted graphs: single-GPU DeviceGraph)")
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        src = torch.from_numpy(A.row.astype(np.int64)).to(dev)
        dst = torch.from_numpy(A.col.astype(np.int64)).to(dev)
        if not directed:                                   # stored symmetric: name every edge once
            once = src <= dst
            src, dst = src[once], dst[once]
        return DistGraph.from_edges(A.shape[0], src, dst, directed=directed, normalization=normalization, group=group)

    @property
    def symdeg(self) -> bool:
        """Scales derivable from the local row pointers inside the kernels (undirected, symmetric normalisation)."""
        return (not self.directed) and self.normalization == "symmetric"

    def local_vector(self, dense, dtype: torch.dtype) -> torch.Tensor:
        """This rank's slice (internal order) of a dense vector in USER node order (host array or device tensor)."""
        dev = self.view.indptr.device
        x = torch.as_tensor(dense).to(device=dev, dtype=dtype).reshape(-1)
        if x.numel() != self.n_nodes:
            raise Exception("Graph signal array dimensions " + str(x.numel()) + " should be equal to graph nodes " + str(self.n_nodes))
        if self.n_global != self.n_nodes:
            x = torch.cat([x, torch.zeros(self.n_global - self.n_nodes, dtype=dtype, device=dev)])
        return x[self._user_of_local()].contiguous()

    def _user_of_local(self) -> torch.Tensor:
        if getattr(self, "user_of_local", None) is None:
            dev = self.view.indptr.device
            full = torch.full((self.n_global,), -1, dtype=torch.int64, device=dev)
            full[self.new_id.long()] = torch.arange(self.n_nodes, device=dev)
            pad = torch.nonzero(full < 0).reshape(-1)
            full[pad] = torch.arange(self.n_nodes, self.n_nodes + pad.numel(), device=dev)
            self.user_of_local = full[self.offset: self.offset + self.n_local]
        return self.user_of_local

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

    def push_sum(self, right_local: torch.Tensor) -> torch.Tensor:
        """L_i * sum_{k: i -> k} right_k for the local nodes i (fp64): one all-gather of the local factors and one plain
        gather pass over the push structure.  Row sums of M (right = R) and of M diag(d) (right = R o d)."""
        from .graph import dtype_code, span_struct
        lib = C.lib()
        f64 = torch.float64
        full = torch.empty(self.n_global, dtype=f64, device=self.view.indptr.device)
        dist.all_gather_into_tensor(full, right_local.to(f64).contiguous(), group=self.group)
        out = torch.empty(self.n_local, dtype=f64, device=full.device)
        view = self.out_view if self.out_view is not None else self.view
        cs = view.cstruct(f64, hsell=False)
        C.check(lib.pgb_spmv(ctypes.byref(cs), dtype_code(f64), C.ptr(full), C.ptr(self.vec("L", f64)), None, None,
                             C.ptr(out), span_struct(view.span_ws()), C.stream_ptr()))
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
        elif name == "rowsum":                               # out-degree of the local nodes (undirected: the degree)
            out = (self.view.indptr[1:] - self.view.indptr[:-1]).to(f64)
        elif name == "colsum":
            out = self.vec("rowsum", f64)
        elif name in ("L", "R"):                             # preprocessing.py:109-138, zeros kept (S[S != 0] = 1/S)
            from .graph import _SCALE_KINDS
            kind = _SCALE_KINDS[self.normalization][0 if name == "L" else 1]
            d = self.vec("rowsum" if name == "L" else "colsum", f64)
            if kind == C.SCALE_ONE:
                out = torch.ones_like(d)
            else:
                base = torch.sqrt(d) if kind == C.SCALE_RSQRT else d
                out = torch.where(base != 0, 1.0 / torch.where(base != 0, base, torch.ones_like(base)), torch.zeros_like(base))
        elif name == "Lp":
            L = self.vec("L", f64)
            out = torch.where(L == 0, torch.ones_like(L), L)
        elif name == "w":
            out = self.vec("Lp", f64) * self.vec("R", f64)
        elif name == "sq":
            out = 1.0 / self.vec("Lp", f64)
        elif name == "degM":                                 # row sums of the normalised matrix: L_i * sum_k a_ik R_k
            out = self.push_sum(self.vec("R", f64))
        elif name == "c":
            out = self.vec("sq", f64) * self.vec("degM", f64)
        else:
            raise KeyError(name)
        self._cache[key] = out
        return out


class DistFilter:
    """Driver shared by the row-partitioned filters: the reference's rank() prologue (abstract_filters.py:44-65) on this
    rank's slice, then one fused step per iteration with the exchange inside the update kernel (symmetric memory) or
    an NCCL all-gather after it, the convergence test of convergence.py:77-101 decided identically on every rank from
    the rank-ordered sums, and the run-ahead / read-back loop of the single-GPU filters.  Every rank calls ``rank``
    collectively and gets its slice of the scores (internal id order, ``g.offset`` onwards)."""

    def __init__(self, tol: Optional[float] = 1e-6, max_iters: int = 100, end_modulo: int = 1, error_type: str = "mabs",
                 dtype: torch.dtype = torch.float32, chunk: int = 4, preserve_norm: bool = True):
        self.tol, self.max_iters, self.end_modulo = tol, int(max_iters), int(end_modulo)
        self.error_type, self.dtype, self.chunk, self.preserve_norm = error_type, dtype, int(chunk), preserve_norm
        self.poison = os.environ.get("PGB_PEER_POISON", "0") == "1"   # NaN-fill the exchanged buffers before a solve
        self.iteration = 0
        self.elapsed_time = None

    # -- personalization -------------------------------------------------------------------------------------------
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

    def _personalization(self, g, seeds, values, p_local, norm, dense):
        if dense is not None:
            p = g.local_vector(dense, self.dtype)
            nrm = torch.tensor([float(p.abs().sum(dtype=torch.float64))], dtype=torch.float64, device=p.device)
            dist.all_reduce(nrm, group=g.group)
            return p, float(nrm.item())
        if p_local is None:
            return self.local_personalization(g, seeds, values)
        return p_local, norm

    # -- the loop ---------------------------------------------------------------------------------------------------
    def _solve(self, g: DistGraph, p, norm, kind, alpha, w, sq, c, coef, coefvec, coef_table, quotient):
        """kind "affine": z' = (alpha*w*acc + q)/S with q = (coefvec or coef)*p/sq;  kind "poly": ranks += coef_table[k]*pow."""
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
        err_code = _error_code(self.error_type)
        if err_code == C.ERR_MAX and g.peer_buffers(dtype) is None:
            raise Exception("MaxDifference on the row-partitioned path needs the symmetric-memory exchange")
        sf = [0.0] * C.STATE_LEN
        si = [0] * C.STATE_LEN
        sf[C.SF_ALPHA], sf[C.SF_INVS] = float(alpha), 1.0
        sf[C.SF_TOL] = 0.0 if self.tol is None else max(float(self.tol), float(np.finfo(float).eps))
        sf[C.SF_MEAN] = 1.0 if err_code in (C.ERR_L1, C.ERR_MAX) else float(g.n_nodes)   # Mabs divides by the node count
        sf[C.SF_NORM] = norm
        si[C.SI_MAX_ITERS], si[C.SI_END_MODULO] = self.max_iters, max(self.end_modulo, 1)
        si[C.SI_ERR_MODE], si[C.SI_QUOTIENT] = err_code, int(bool(quotient))
        state_f64 = torch.tensor(sf, dtype=torch.float64, device=dev)
        state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
        err_hist = torch.zeros(self.max_iters + 2, dtype=torch.float64, device=dev)

        sq_full = g.vec("sq", dtype)
        peer = g.peer_buffers(dtype)
        if peer is not None:
            zfull = [peer["z"][:g.n_global], peer["z"][g.n_global:]]
        else:
            zfull = [torch.empty(g.n_global, dtype=dtype, device=dev), torch.empty(g.n_global, dtype=dtype, device=dev)]
        affine = kind == "affine"
        q = torch.empty(n_loc, dtype=dtype, device=dev) if affine else None
        ranks = torch.zeros(n_loc, dtype=dtype, device=dev) if not affine else None
        coef_dev = torch.tensor(coef_table, dtype=torch.float64, device=dev) if not affine else None
        if peer is not None and self.poison:
            # parity runs: every entry a peer fails to deliver (reader masks leave unread entries alone) is a NaN
            peer["z"].fill_(float("nan"))
            peer["hz"].barrier(channel=1)
        if peer is not None:
            # start vector straight into every rank's buffer 0 (no all-gather); the barrier orders it before step 1
            C.check(lib.pgb_affine_init_peer(n_loc, code, C.ptr(p), None, C.ptr(sq_full), C.ptr(c), float(coef),
                                             C.ptr(coefvec), None, off, C.ptr(q), C.ptr(state_f64),
                                             ctypes.byref(peer["peers"][0]), st))
            peer["hz"].barrier(channel=2)
        else:
            C.check(lib.pgb_affine_init(n_loc, code, C.ptr(p), None, C.ptr(sq_full), C.ptr(c), float(coef), C.ptr(coefvec),
                                        None, off, C.ptr(zfull[0]), C.ptr(q), C.ptr(state_f64), st))
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

        def step(k):
            if peer is not None:
                # exchange fused into the step: the update kernel writes z' and the convergence sums into every
                # rank; a symmetric-memory barrier (one small kernel) replaces all-gather + all-reduce
                pk = ctypes.byref(peer["peers"][k & 1])
                if affine:
                    C.check(lib.pgb_affine_step_peer(ctypes.byref(cs), code, float(alpha), C.ptr(w), C.ptr(sq), C.ptr(c),
                                                     C.ptr(q), C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64),
                                                     C.ptr(state_i32), C.ptr(err_hist), span_struct(ws), k, pk, st))
                else:
                    C.check(lib.pgb_poly_step_peer(ctypes.byref(cs), code, C.ptr(w), C.ptr(sq), C.ptr(coef_dev),
                                                   C.ptr(ranks), C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64),
                                                   C.ptr(state_i32), C.ptr(err_hist), span_struct(ws), k, pk, st))
                peer["hz"].barrier(channel=0)
                C.check(lib.pgb_state_finalize_peer(C.ptr(state_f64), C.ptr(state_i32), C.ptr(err_hist),
                                                    peer["acc"].data_ptr() + (k & 1) * 2 * C.MAX_PEERS * 8, g.world, st))
                return
            if affine:
                C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, float(alpha), C.ptr(w), C.ptr(sq), C.ptr(c), C.ptr(q),
                                             C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64), C.ptr(state_i32),
                                             C.ptr(err_hist), span_struct(ws), k, 1, 0, st))
            else:
                C.check(lib.pgb_poly_steps(ctypes.byref(cs), code, C.ptr(w), C.ptr(sq), C.ptr(coef_dev), C.ptr(ranks),
                                           C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64), C.ptr(state_i32),
                                           C.ptr(err_hist), span_struct(ws), k, 1, 0, st))
            out = zfull[k & 1]
            dist.all_gather_into_tensor(out, out[off:off + n_loc], group=g.group)
            dist.all_reduce(acc, group=g.group)
            C.check(lib.pgb_state_finalize(C.ptr(state_f64), C.ptr(state_i32), C.ptr(err_hist), st))

        budget, done = self.max_iters - 1, 0
        stop, steps, iteration = C.RUNNING, 0, 1
        # run-ahead chunk: a previous solve on this graph is the best guess of how many steps this one needs,
        # so repeated solves read the device state back once
        chunk = max(self.chunk, min(getattr(self, "_steps_hint", 0) + 1, 64), 1)
        while done < budget:
            count = min(chunk, budget - done)
            for j in range(count):
                step(done + 1 + j)
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
        scale = norm if self.preserve_norm else 1.0
        if affine:
            zl = zfull[steps & 1][off:off + n_loc]
            C.check(lib.pgb_unscale(n_loc, code, C.ptr(zl.contiguous()), C.ptr(sq_full), None, scale, None, C.ptr(result), st))
        else:
            C.check(lib.pgb_unscale(n_loc, code, C.ptr(ranks), None, None, scale, None, C.ptr(result), st))
        C.count_launches(1)
        mark("end")
        if timing:
            torch.cuda.synchronize()
            self.timing = {"steps": steps, "launched": done,
                           "gpu_ms": {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])},
                           "host_ms": {b[0]: (b[2] - a[2]) * 1e3 for a, b in zip(marks, marks[1:])}}
        self.elapsed_time = time.perf_counter() - t0
        return result

    def _scales(self, g: DistGraph):
        """(w, sq) the kernels read, or (None, None) when they derive them from the row pointers."""
        if g.symdeg:
            return None, None
        return g.vec("w", self.dtype), g.vec("sq", self.dtype)

    def gather_user_order(self, g: DistGraph, local_scores: torch.Tensor) -> torch.Tensor:
        """All ranks: the full score vector in USER node order (testing / small graphs)."""
        full = torch.empty(g.n_global, dtype=local_scores.dtype, device=local_scores.device)
        dist.all_gather_into_tensor(full, local_scores.contiguous(), group=g.group)
        return full[g.new_id.long()][: g.n_nodes]


class DistPageRank(DistFilter):
    """``pg.PageRank`` (adhoc.py:34-36 with RecursiveGraphFilter's quotient, abstract_filters.py:126-136)."""

    def __init__(self, alpha: float = 0.85, *args, use_quotient: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha, self.use_quotient = alpha, use_quotient

    def rank(self, g: DistGraph, seeds=None, values=None, p_local=None, norm=None, personalization=None) -> torch.Tensor:
        """``seeds``: user node ids (host array) with optional ``values`` (default 1); or a prebuilt ``p_local``/``norm``
        pair from :meth:`local_personalization`; or ``personalization``: a dense vector in user order."""
        p, norm = self._personalization(g, seeds, values, p_local, norm, personalization)
        if norm == 0:
            self.iteration = 0
            return p
        w, sq = self._scales(g)
        return self._solve(g, p, norm, "affine", self.alpha, w, sq, g.vec("c", self.dtype), 1 - self.alpha, None, None,
                           self.use_quotient)


class DistAbsorbingWalks(DistFilter):
    """``pg.AbsorbingWalks`` (adhoc.py:157-169): (conv(r, M) o deg + p o absorb) / (absorb + deg) with deg the row sums of
    the normalised operator; ``absorption`` is a dense user-order vector (default: ones)."""

    def __init__(self, alpha: float = 1 - 1.E-6, *args, use_quotient: bool = True, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha, self.use_quotient = alpha, use_quotient

    def rank(self, g: DistGraph, seeds=None, values=None, p_local=None, norm=None, personalization=None,
             absorption=None) -> torch.Tensor:
        p, norm = self._personalization(g, seeds, values, p_local, norm, personalization)
        if norm == 0:
            self.iteration = 0
            return p
        f64 = torch.float64
        dev = g.view.indptr.device
        rate = (1 - self.alpha) / self.alpha                  # adhoc.py:158
        ab = (torch.ones(g.n_local, dtype=f64, device=dev) if absorption is None else g.local_vector(absorption, f64)) * rate
        degM = g.vec("degM", f64)
        denom = ab + degM
        d1 = degM / denom
        w_run = (g.vec("w", f64) * d1).to(self.dtype)         # coefficient of the gathered sum
        coefvec = (ab / denom).to(self.dtype)                 # coefficient of the personalization
        gsum = g.push_sum(g.vec("R", f64) * d1)               # row sums of M diag(d1): the next normaliser stays linear
        c_run = (g.vec("sq", f64) * gsum).to(self.dtype)
        return self._solve(g, p, norm, "affine", 1.0, w_run, g.vec("sq", self.dtype), c_run, 0.0, coefvec, None,
                           self.use_quotient)


class DistClosedFormGraphFilter(DistFilter):
    """Taylor-coefficient polynomial filters in the node space (abstract_filters.py:196-256) on pgb_poly_step_peer."""

    def _coefficient(self, previous_coefficient, iteration: int) -> float:
        raise Exception("Use a derived class of DistClosedFormGraphFilter that implements the _coefficient method")

    def rank(self, g: DistGraph, seeds=None, values=None, p_local=None, norm=None, personalization=None) -> torch.Tensor:
        p, norm = self._personalization(g, seeds, values, p_local, norm, personalization)
        if norm == 0:
            self.iteration = 0
            return p
        coefs, prev = [0.0], None
        for k in range(1, max(self.max_iters, 1) + 1):        # step k runs with convergence.iteration == k
            prev = self._coefficient(prev, k)
            coefs.append(float(prev))
        w, sq = self._scales(g)
        return self._solve(g, p, norm, "poly", 1.0, w, sq, None, 0.0, None, coefs, False)


class DistHeatKernel(DistClosedFormGraphFilter):
    def __init__(self, t: float = 3, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.t = t

    def _coefficient(self, previous_coefficient, iteration):   # adhoc.py:113-116
        return 1. if previous_coefficient is None else previous_coefficient * self.t / (iteration + 1)


class DistPageRankClosed(DistClosedFormGraphFilter):
    def __init__(self, alpha: float = 0.85, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = alpha

    def _coefficient(self, previous_coefficient, iteration):   # adhoc.py:83-84
        return 1. if previous_coefficient is None else previous_coefficient * self.alpha


class DistGenericGraphFilter(DistClosedFormGraphFilter):
    def __init__(self, weights=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.weights = list(weights) if weights is not None else [0.9] * 10

    def _coefficient(self, _, iteration):                      # low_pass.py:23-26
        return 0 if iteration > len(self.weights) else self.weights[iteration - 1]


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
