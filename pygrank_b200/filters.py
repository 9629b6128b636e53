"""Fused on-device graph filters — the host-side mirror of pygrank's filter interface.

Same class names, constructor arguments, iteration semantics and error behaviour as the
reference (paths under /root/reference/pygrank/algorithms):

* ``GraphFilter.rank``                 filters/abstract_filters.py:44-65
* ``RecursiveGraphFilter._step``       filters/abstract_filters.py:126-136  (``use_quotient``)
* ``ClosedFormGraphFilter._step``      filters/abstract_filters.py:196-256  (taylor, node space)
* ``PageRank`` / ``PageRankClosed`` / ``HeatKernel`` / ``AbsorbingWalks``   filters/adhoc.py:10-174
* ``GenericGraphFilter``               filters/low_pass.py:5-26
* ``ConvergenceManager``               convergence.py:9-104

but the whole loop runs on the device: every iteration is ONE launch of the fused merge-path
kernel (csrc/spmv_fused.cu), convergence is decided by the last CTA of that launch, and the host
enqueues iterations ahead of the device (run-ahead launches after convergence exit immediately),
reading the 64-byte state back only once per chunk.  The unchanged reference drivers also run on
this engine through the backend plugin (pygrank_b200/backend.py); these classes are the fast path.
"""
from __future__ import annotations

import ctypes
import os
import time
from typing import Optional, Sequence

import numpy as np
import torch

from . import _capi as C
from .graph import (DeviceGraph, as_device_graph, dtype_code, in_kernel_dropout, preprocessor as device_preprocessor,
                    span_struct)

_ERROR_NAMES = {"mabs": C.ERR_MABS, "l1": C.ERR_L1, "msq": C.ERR_MSQ, "iters": C.ERR_ITERS,
                "max": C.ERR_MAX, "maxdifference": C.ERR_MAX}


def _error_code(error_type) -> int:
    name = error_type if isinstance(error_type, str) else getattr(error_type, "__name__", str(error_type))
    name = name.lower()
    if name not in _ERROR_NAMES:
        raise Exception("the device engine fuses the Mabs, L1, MSQ, MaxDifference and 'iters' criteria; got "
                        + str(error_type))
    return _ERROR_NAMES[name]


class ConvergenceManager:
    """Argument-compatible with convergence.py:24-60; the checks themselves run on the device
    (``finalize_state`` in csrc/spmv_fused.cu) and this object reports their outcome."""

    def __init__(self, tol: Optional[float] = 1.E-6, error_type="mabs", max_iters: int = 100, end_modulo: int = 1,
                 iter_exception=Exception):
        self.tol = tol
        self.error_type = error_type
        self.max_iters = int(max_iters)
        self.end_modulo = int(end_modulo)
        self.iter_exception = iter_exception
        self.iteration = 0
        self.elapsed_time = None
        self.errors = None

    def __str__(self):
        return str(self.iteration) + " iterations (" + str(self.elapsed_time) + " sec)"


class RankResult:
    """What ``rank`` returns: the scores as a device tensor in the user's node order
    (``.np``, like ``GraphSignal.np``, signals.py:81-83) with dict-style access by node."""

    def __init__(self, graph: DeviceGraph, values: torch.Tensor):
        self.graph = graph
        self.np = values

    def numpy(self) -> np.ndarray:
        return self.np.detach().cpu().numpy()

    def __getitem__(self, node) -> float:
        return float(self.np[self.graph._pygrank_node2id[node]])

    def __len__(self):
        return self.graph.n


def _personalization(g: DeviceGraph, data, dtype: torch.dtype):
    """``to_signal`` + the L1 norm of abstract_filters.py:52 — returns (device vector, norm)."""
    dev = g.out_view.indptr.device
    n = g.n
    if isinstance(data, RankResult):
        data = data.np
    if data is None:                                         # signals.py:59-60: a signal of ones
        return torch.ones(n, dtype=dtype, device=dev), float(n)
    if isinstance(data, torch.Tensor):
        p = data.to(device=dev, dtype=dtype).contiguous().reshape(-1)
        if p.numel() != n:
            raise Exception("Graph signal array dimensions " + str(p.numel()) + " should be equal to graph nodes " + str(n))
        return p, float(p.abs().sum(dtype=torch.float64))
    if isinstance(data, dict) or (isinstance(data, (list, tuple)) and len(data) != n):
        if not isinstance(data, dict):
            data = {v: 1 for v in data}                      # signals.py:314-315
        node2id = g._pygrank_node2id
        idx = np.fromiter((node2id[k] for k in data.keys()), dtype=np.int64, count=len(data))
        vals = np.fromiter((float(v) for v in data.values()), dtype=np.float64, count=len(data))
        p = torch.zeros(n, dtype=dtype, device=dev)
        if len(idx):
            p[torch.from_numpy(idx).to(dev)] = torch.from_numpy(vals).to(device=dev, dtype=dtype)
        return p, float(np.abs(vals).sum())
    arr = np.asarray(data, dtype=np.float64).reshape(-1)
    if arr.shape[0] != n:
        raise Exception("Graph signal array dimensions " + str(arr.shape[0]) + " should be equal to graph nodes " + str(n))
    return torch.from_numpy(arr).to(device=dev, dtype=dtype), float(np.abs(arr).sum())


class GraphFilter:
    """Base of the device filters (constructor arguments of abstract_filters.py:14-39 that concern
    the hot path).  ``dtype`` selects fp64 (parity mode, default) or fp32."""

    def __init__(self, preprocessor=None, convergence=None, preserve_norm: bool = True,
                 normalization: str = "auto", renormalize=False, assume_immutability: bool = False,
                 tol: Optional[float] = 1.E-6, error_type="mabs", max_iters: int = 100, end_modulo: int = 1,
                 dtype: torch.dtype = torch.float64, relabel: str = "hub", chunk: int = 8):
        self.preprocessor = preprocessor if preprocessor is not None else device_preprocessor(
            normalization=normalization, renormalize=renormalize, assume_immutability=assume_immutability,
            relabel=relabel)
        self.convergence = convergence if convergence is not None else ConvergenceManager(
            tol=tol, error_type=error_type, max_iters=max_iters, end_modulo=end_modulo)
        self.preserve_norm = preserve_norm
        self.dtype = dtype
        self.relabel = relabel
        self.chunk = int(chunk)

    # -- public API -------------------------------------------------------------------------
    def __call__(self, graph=None, personalization=None, *args, **kwargs) -> RankResult:
        return self.rank(graph, personalization, *args, **kwargs)

    def __add__(self, other):                                 # abstract_filters.py:88-98
        if isinstance(other, ConvergenceManager):
            self.convergence = other
        elif hasattr(other, "__name__") and other.__name__ == "preprocess":
            self.preprocessor = other
        else:
            raise Exception("Can only add convergence managers and preprocessors to graph filters")
        return self

    def rank(self, graph=None, personalization=None, warm_start=None, graph_dropout: float = 0, **kwargs) -> RankResult:
        if graph is None and isinstance(personalization, RankResult):
            graph = personalization.graph
        g = self._device_graph(graph)
        if graph_dropout != 0:
            self._check_dropout(g)
        p, norm = _personalization(g, personalization, self.dtype)
        cm = self.convergence
        cm.iteration, cm.errors = 0, None
        t0 = time.perf_counter()
        if norm == 0:                                         # abstract_filters.py:53-54
            cm.elapsed_time = time.perf_counter() - t0
            return RankResult(g, p)
        if g.pathological:
            raise Exception("a row of the graph has entries but a zero weight sum; the fused filters cannot represent it")
        warm = None
        if warm_start is not None:
            warm, _ = _personalization(g, warm_start, self.dtype)
        with in_kernel_dropout(graph_dropout):
            out = self._run(g, p, norm, warm, **kwargs)
        cm.elapsed_time = time.perf_counter() - t0
        return RankResult(g, out)

    def propagate(self, graph, features, *args, **kwargs) -> torch.Tensor:
        """``NodeRanking.propagate`` (signals.py:225-226): one rank per feature column.  Filters
        with a batched kernel (``_run_batched``) advance a panel of columns per pass over the CSR;
        the others run column by column like the reference."""
        g = self._device_graph(graph)
        cols = features if isinstance(features, torch.Tensor) else torch.as_tensor(np.asarray(features))
        if cols.dim() != 2 or cols.shape[0] != g.n:
            raise Exception("propagate expects a features matrix with one row per node")
        if self._can_batch(g, *args, n_columns=int(cols.shape[1]), **kwargs):
            return self._propagate_batched(g, cols, *args, **kwargs)
        self.convergence.iterations = []
        out = torch.empty((g.n, int(cols.shape[1])), dtype=self.dtype, device=g.out_view.indptr.device)
        for c in range(cols.shape[1]):
            out[:, c] = self.rank(g, cols[:, c].contiguous(), *args, **kwargs).np
            self.convergence.iterations.append(self.convergence.iteration)
        return out

    def _can_batch(self, g, *args, n_columns: int = 2, **kwargs) -> bool:
        return False

    def _panel_family(self, g) -> Optional[str]:
        """Which batched kernel propagate() uses: "hsell" = panels of 4 fp32 / 2 fp64 columns through the hub-blocked
        form (pgb_affine_steps_panel: 16-byte shared-memory / texture gathers, the index streams read once per panel);
        "csr" = panels of 8 / 4 columns through the item-stream kernel (pgb_affine_steps_batched); None = column by
        column.  PGB_PANEL: 0 = never, csr = the item-stream panel, anything else = batch with the best family."""
        forced = os.environ.get("PGB_PANEL")
        if forced == "0":
            return None
        if forced == "csr":
            return "csr"
        return "hsell" if g.in_view.hsell_panel() is not None else "csr"

    def _propagate_panels(self, g: DeviceGraph, cols: torch.Tensor, poly_coefs=None, **kwargs) -> torch.Tensor:
        """All feature columns through ``pgb_affine_steps_panel``: the panel's ``pgb_hsell_panel_width`` columns are
        SLOTS scheduled on the device — a column enters a free slot (normalisation, scaled start vector, affine term,
        state: abstract_filters.py:52-56), iterates with its own alpha, normaliser, error and stop decision, and is
        written out when it stops while the others keep going, so no slot waits for the slowest column of a fixed group
        and the host only polls the number of finished columns.  Results and iteration counts equal the reference's
        column-by-column loop (signals.py:225-226).  ``poly_coefs`` (closed-form filters): the coefficient of step k at
        index k — every slot accumulates ``ranks += coef[k] * power_k`` at its own step while its power advances
        (abstract_filters.py:225-256) and leaves with the accumulated result."""
        lib = C.lib()
        dtype, code = self.dtype, dtype_code(self.dtype)
        f64, i32 = torch.float64, torch.int32
        dev, n = g.out_view.indptr.device, g.n
        st = C.stream_ptr()
        cm = self.convergence
        B = int(cols.shape[1])
        PB = lib.pgb_hsell_panel_width(code)
        poly = poly_coefs is not None
        if poly:
            alpha, alpha_s, w_run, c_run, coef, coefvec, col_alphas, quotient = 1.0, 1.0, None, None, 0.0, None, None, False
        else:
            a = self._affine_args(g, **kwargs)
            alpha, alpha_s, w_run, c_run, coef, coefvec = (a["alpha"], a["alpha_s"], a["w_run"], a["c_run"], a["coef"],
                                                           a["coefvec"])
            col_alphas = self._column_alphas(B)               # per column (multiplier, alpha_s, coef) or None
            quotient = bool(self.use_quotient)
        sq = g.vec("sq", dtype)
        c = None if poly else (c_run if c_run is not None else g.vec("c", dtype))
        symdeg = g.symdeg and w_run is None
        w = None if symdeg else (w_run if w_run is not None else g.vec("w", dtype))
        view = g.in_view
        form = view.hsell_panel()
        err_code = _error_code(cm.error_type)
        tol = 0.0 if cm.tol is None else max(float(cm.tol), float(np.finfo(float).eps))
        cols = cols.to(device=dev, dtype=dtype)               # [n, B] in user order, any strides
        out = torch.empty((n, B), dtype=dtype, device=dev)
        iterations, errors = [], []
        if B == 0:
            cm.iterations, cm.column_errors = [], []
            return out
        t0 = time.perf_counter()
        esz = cols.element_size()
        hist = cm.max_iters + 2
        L = C.STATE_LEN
        # working set of the panel (allocated once, reused by every group of columns)
        yacc = torch.zeros((form.n_slices + 1) * 32 * PB, dtype=dtype, device=dev)
        tail_queue = torch.zeros(1, dtype=i32, device=dev)
        zbuf = [torch.zeros((n, PB), dtype=dtype, device=dev), torch.zeros((n, PB), dtype=dtype, device=dev)]
        q = torch.zeros((n, PB), dtype=dtype, device=dev)     # affine: the constant term; polynomial: the accumulated results
        coef_dev = torch.tensor([float(v) for v in poly_coefs], dtype=f64, device=dev) if poly else None
        err_hist = torch.zeros((PB, hist), dtype=f64, device=dev)
        si_host = np.zeros(PB * L + 4, dtype=np.int32)        # + ticket, panel stop word, executed steps, spare
        si_host[:PB * L].reshape(PB, L)[:, C.SI_STOP] = C.CONVERGED          # empty slots never run
        si_host[PB * L + 1] = C.CONVERGED
        # Columns are staged in groups: column-major blocks in the engine's row order (pgb_panel_stage), so that a slot
        # loads / stores its column as one contiguous stream instead of one sector per value at random rows.  A group is
        # as large as memory comfortably allows (input + output stage).
        free_bytes = torch.cuda.mem_get_info(dev)[0]
        group = int(max(4 * PB, min(B, (free_bytes * 2 // 5) // max(2 * n * esz, 1))))
        group = int(getattr(self, "panel_group", group))
        stage_in = torch.empty((min(group, B), n), dtype=dtype, device=dev)
        stage_out = torch.empty((min(group, B), n), dtype=dtype, device=dev)
        chunk = max(int(getattr(self, "panel_chunk", 8)), 1)  # steps enqueued between polls of the finished count
        polls = [torch.empty(4, dtype=i32).pin_memory() for _ in range(2)]
        events = [torch.cuda.Event(), torch.cuda.Event()]
        C.count_launches(1)
        marks = [] if os.environ.get("PGB_PANEL_TIMING") else None

        def mark(label):
            if marks is not None:
                torch.cuda.synchronize()
                marks.append((label, time.perf_counter()))

        mark("buffers")
        for j0 in range(0, B, group):
            G = min(group, B - j0)
            C.check(lib.pgb_panel_stage(n, code, 0, cols.data_ptr(), int(cols.stride(0)), int(cols.stride(1)),
                                        C.ptr(g.perm), j0, G, C.ptr(stage_in), st))
            for t in (zbuf[0], zbuf[1], q, yacc, tail_queue):
                t.zero_()
            sf = torch.zeros((PB, L), dtype=f64, device=dev)
            si = torch.from_numpy(si_host).to(dev)
            sched = torch.zeros(4, dtype=i32, device=dev)
            slot_col = torch.full((PB,), -1, dtype=i32, device=dev)
            slot_plan = torch.full((2 * PB,), -1, dtype=i32, device=dev)
            plan_norm = torch.zeros(PB, dtype=f64, device=dev)
            col_result = torch.zeros((G, 4), dtype=i32, device=dev)
            keep_hist = G * hist <= (1 << 24)                 # error histories of every column (128 MB at most)
            col_err = torch.zeros((G, hist), dtype=f64, device=dev) if keep_hist else None
            params = None
            if col_alphas is not None:
                params = torch.tensor([[float(v) for v in t] for t in col_alphas[j0:j0 + G]], dtype=f64,
                                      device=dev).contiguous()
            job = C.PanelJob(G, hist, stage_in.data_ptr(), 1, n, stage_out.data_ptr(), 1, n, None, C.ptr(sq),
                             None if params is not None else C.ptr(coefvec), C.ptr(params), float(alpha),
                             float(alpha_s), float(coef), tol, 1.0 if err_code in (C.ERR_L1, C.ERR_MAX) else float(n),
                             int(cm.max_iters), max(int(cm.end_modulo), 1), err_code, int(quotient),
                             int(self.preserve_norm), C.ptr(sched), C.ptr(slot_col), C.ptr(slot_plan),
                             C.ptr(plan_norm), C.ptr(col_result), C.ptr(col_err), int(poly), 0,
                             C.ptr(q) if poly else None, C.ptr(coef_dev))
            budget = (cm.max_iters + 2) * (G // PB + 2) + chunk   # more steps than any schedule needs: guards a driver bug

            # chunk i+1 is enqueued before the finished count after chunk i is looked at (pinned copy + event): the
            # device never waits for the host; the launches enqueued past the end are no-ops
            def enqueue(i):
                C.check(lib.pgb_affine_steps_panel(ctypes.byref(form.struct), C.ptr(view.indptr), code,
                                                   ctypes.byref(job), C.ptr(w), C.ptr(c), None if poly else C.ptr(q),
                                                   C.ptr(zbuf[0]),
                                                   C.ptr(zbuf[1]), C.ptr(sf), C.ptr(si), C.ptr(err_hist), C.ptr(yacc),
                                                   C.ptr(tail_queue), i * chunk + 1, chunk, st))
                C.count_launches(6 * chunk)
                polls[i & 1].copy_(sched, non_blocking=True)
                events[i & 1].record()

            mark("staged in")
            enqueue(0)
            i = 0
            while True:
                enqueue(i + 1)
                events[i & 1].synchronize()
                if int(polls[i & 1][1]) >= G:
                    break
                i += 1
                if i * chunk > budget:
                    raise Exception("pygrank_b200: panel scheduling did not terminate")
            mark("job")
            C.check(lib.pgb_panel_stage(n, code, 1, out.data_ptr(), B, 1, C.ptr(g.perm), j0, G, C.ptr(stage_out), st))
            C.count_launches(2)
            res = col_result.cpu().numpy()
            mark("staged out")
            for j in range(G):
                it, stop, steps, live = (int(v) for v in res[j])
                if not live:                                                    # abstract_filters.py:53-54
                    iterations.append(0)
                    errors.append(None)
                    continue
                iterations.append(it)
                errors.append(col_err[j, 1:steps + 1] if col_err is not None else None)
                if stop == C.MAX_ITERS and err_code != C.ERR_ITERS and cm.iter_exception is not None:
                    raise cm.iter_exception("Could not converge within " + str(cm.max_iters) + " iterations")
        if marks is not None:
            prev = t0
            for label, t in marks:
                print(f"[panel] {label}: {(t - prev) * 1e3:.2f} ms", flush=True)
                prev = t
        cm.iterations = iterations
        cm.iteration = iterations[-1] if iterations else 0
        cm.errors = errors[-1] if errors else None
        cm.column_errors = errors
        cm.elapsed_time = time.perf_counter() - t0
        return out

    def _check_dropout(self, g: DeviceGraph):
        """graph_dropout (abstract_filters.py:59-62) is drawn inside the gather kernel of the hub-blocked form."""
        if g.in_view.hsell(self.dtype) is None:
            raise Exception("in-kernel graph_dropout needs an unweighted graph (hub-blocked form); weighted graphs: "
                            "use the backend plugin path")
        if getattr(self, "use_quotient", False):
            raise Exception("graph_dropout with use_quotient=True is not fused (the next normaliser is no longer linear "
                            "in the iterate); pass use_quotient=False or use the backend plugin path")

    def _device_graph(self, graph) -> DeviceGraph:
        """The preprocessor's output as a DeviceGraph.  A custom preprocessor that returns a host matrix (the
        reference's ``pg.preprocessor`` under numpy, a callable normalisation, an ``Adjacency`` around a scipy
        matrix) has ALREADY normalised it (preprocessing.py:104-145): it is uploaded as is, like
        ``scipy_sparse_to_backend`` does, never normalised a second time."""
        g = self.preprocessor(graph)
        if isinstance(g, DeviceGraph):
            pass
        elif isinstance(getattr(g, "array", None), DeviceGraph):
            g = g.array
        else:
            import scipy.sparse as sp
            M = getattr(g, "array", g)
            if not sp.issparse(M):
                raise Exception("the preprocessor returned " + str(type(g)) + ": the fused filters need a DeviceGraph "
                                "(pygrank_b200.preprocessor) or an already normalised scipy matrix")
            g = DeviceGraph.from_scipy(M, directed=True, normalization="none", relabel=self.relabel,
                                       node2id=getattr(g, "_pygrank_node2id", None))
        if g.normalization == "laplacian":
            # the fused steps iterate on D^-1/2 A D^-1/2; the I - M of preprocessing.py:122 is only applied by
            # DeviceGraph.conv, i.e. on the backend plugin route
            raise Exception("normalization='laplacian' is not fused; run the filter through the b200 backend plugin "
                            "(pg.PageRank under pg.Backend('b200')), whose conv applies I - M")
        return g

    # -- machinery shared by the subclasses ---------------------------------------------------
    def _new_state(self, g: DeviceGraph, norm: float, alpha_s: float, quotient: bool):
        cm = self.convergence
        dev = g.out_view.indptr.device
        code = _error_code(cm.error_type)
        sf = [0.0] * C.STATE_LEN
        si = [0] * C.STATE_LEN
        sf[C.SF_ALPHA] = float(alpha_s)
        sf[C.SF_INVS] = 1.0
        sf[C.SF_TOL] = 0.0 if cm.tol is None else max(float(cm.tol), float(np.finfo(float).eps))  # convergence.py:101
        sf[C.SF_MEAN] = 1.0 if code in (C.ERR_L1, C.ERR_MAX) else float(g.n)
        sf[C.SF_NORM] = float(norm)
        si[C.SI_MAX_ITERS] = cm.max_iters
        si[C.SI_END_MODULO] = max(cm.end_modulo, 1)
        si[C.SI_ERR_MODE] = code
        si[C.SI_QUOTIENT] = int(bool(quotient))
        state_f64 = torch.tensor(sf, dtype=torch.float64, device=dev)
        state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
        err_hist = torch.zeros(cm.max_iters + 2, dtype=torch.float64, device=dev)
        return state_f64, state_i32, err_hist

    def _drive(self, launch, state_i32: torch.Tensor, err_hist: torch.Tensor):
        """Run-ahead loop: enqueue a chunk of steps, read the state, repeat.  Returns #steps done."""
        cm = self.convergence
        budget = cm.max_iters - 1                             # convergence.py:85-90: at most max_iters-1 steps
        done = 0
        # the previous solve of this filter is the best guess of how many steps the next one needs
        chunk = max(self.chunk, min(getattr(self, "_steps_hint", 0) + 1, 64), 1)
        stop, steps, iteration = C.RUNNING, 0, 1
        while done < budget:
            k = min(chunk, budget - done)
            launch(done + 1, k)
            done += k
            host = state_i32.cpu()
            stop, steps, iteration = int(host[C.SI_STOP]), int(host[C.SI_STEPS]), int(host[C.SI_ITERATION])
            if stop != C.RUNNING:
                break
            chunk = min(chunk * 2, 64)
        if stop == C.RUNNING:                                 # max_iters <= 1: stopped before any step
            iteration, stop = 1, C.MAX_ITERS
        cm.iteration = iteration
        self._steps_hint = steps
        cm.errors = err_hist[1:steps + 1]
        if stop == C.MAX_ITERS and _error_code(cm.error_type) != C.ERR_ITERS and cm.iter_exception is not None:
            raise cm.iter_exception("Could not converge within " + str(cm.max_iters) + " iterations")
        return steps

    def _run(self, g, p, norm, warm, **kwargs):
        raise Exception("Use a derived class of GraphFilter")

    def _work_buffers(self, g: DeviceGraph, dtype, dev):
        """Both iterate buffers, the affine term and the step workspace of this filter on this graph, kept between
        solves (propagate() and tuner sweeps issue hundreds of solves on one graph: 4 allocations and a 64 MB memset
        each otherwise).  Results are never views of these: every solve returns a fresh vector."""
        key = (id(g.in_view), dtype)
        cache = self.__dict__.setdefault("_buffers", {})
        if key not in cache:
            cache.clear()                                     # one graph at a time: do not pin the memory of old ones
            n = g.n
            cache[key] = ([torch.empty(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)],
                          torch.empty(n, dtype=dtype, device=dev), g.in_view.new_span_ws(dtype), g.in_view)
        zbuf, q, ws, _ = cache[key]
        return zbuf, q, ws


class RecursiveGraphFilter(GraphFilter):
    """ranks <- (coefficient_vector * conv(ranks, M) + affine_term) [/ sum], the shape shared by
    PageRank and AbsorbingWalks (abstract_filters.py:126-136)."""

    def __init__(self, use_quotient: bool = True, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(use_quotient, bool) and use_quotient is not None:
            raise Exception("the fused filters take use_quotient=True/False (postprocessor quotients run on the plugin path)")
        self.use_quotient = bool(use_quotient)

    def _affine_args(self, g: DeviceGraph, **kwargs) -> dict:
        """alpha, alpha_s, w_run, c_run, coef, coefvec of the affine recursion (per derived filter)."""
        raise Exception("Use a derived class of RecursiveGraphFilter")

    def _run(self, g, p, norm, warm, **kwargs):
        return self._affine(g, p, norm, warm, **self._affine_args(g, **kwargs))

    def _column_alphas(self, n_columns: int):
        """Per-column (alpha, alpha_s, coef) of a panel, or None when every column runs this filter's own parameters
        (set by sweep())."""
        return getattr(self, "_sweep", None)

    def _can_batch(self, g, warm_start=None, graph_dropout: float = 0, n_columns: int = 2, **kwargs) -> bool:
        # Both panel kernels stream no edge values (every BASELINE config is unweighted).
        if g.in_view.weighted or warm_start is not None or graph_dropout != 0 or g.pathological:
            return False
        family = self._panel_family(g)
        if family is None or (family == "hsell" and self.convergence.max_iters <= 1):
            return False
        if family == "csr" and _error_code(self.convergence.error_type) == C.ERR_MAX:
            return False                                      # the item-stream panel kernel has no max reduction
        if os.environ.get("PGB_PANEL") not in (None, ""):
            return True
        # default: a single column is cheaper through the single-vector kernel (a panel would carry padding columns);
        # without the hub-blocked form the item-stream panel is the only fast path
        return n_columns >= 2 or family == "csr"

    def _propagate_batched(self, g: DeviceGraph, cols: torch.Tensor, **kwargs) -> torch.Tensor:
        kwargs.pop("n_columns", None)
        if self._panel_family(g) == "hsell":
            return self._propagate_panels(g, cols, **kwargs)
        if self._column_alphas(int(cols.shape[1])) is not None:
            raise Exception("per-column alpha needs the hub-blocked panel kernel (unweighted graph, PGB_HSELL=1)")
        return self._propagate_batched_csr(g, cols, **kwargs)

    def _propagate_batched_csr(self, g: DeviceGraph, cols: torch.Tensor, **kwargs) -> torch.Tensor:
        """All feature columns through ``pgb_affine_steps_batched``: panels of ``pgb_panel_width``
        columns share one pass over the index stream per iteration; every column keeps its own
        normaliser, error and stop decision, so results and iteration counts equal the reference's
        column-by-column loop (signals.py:225-226)."""
        lib = C.lib()
        dtype, code = self.dtype, dtype_code(self.dtype)
        f64 = torch.float64
        dev, n = g.out_view.indptr.device, g.n
        st = C.stream_ptr()
        cm = self.convergence
        B = int(cols.shape[1])
        PB = lib.pgb_panel_width(code)
        kwargs.pop("n_columns", None)
        a = self._affine_args(g, **kwargs)
        alpha, alpha_s, w_run, c_run, coef, coefvec = (a["alpha"], a["alpha_s"], a["w_run"], a["c_run"], a["coef"],
                                                       a["coefvec"])
        sq = g.vec("sq", dtype)
        c = c_run if c_run is not None else g.vec("c", dtype)
        symdeg = g.symdeg and w_run is None
        w = None if symdeg else (w_run if w_run is not None else g.vec("w", dtype))
        sq_arg = None if symdeg else sq
        view = g.in_view
        cs = view.cstruct(dtype, hsell=False)                 # the panel kernel reads the item stream
        err_code = _error_code(cm.error_type)
        tol = 0.0 if cm.tol is None else max(float(cm.tol), float(np.finfo(float).eps))
        perm = None if g.perm is None else g.perm.long()
        out = torch.empty((n, B), dtype=dtype, device=dev)
        acc = torch.zeros(max(view.n_tiles, 1) * PB, dtype=f64, device=dev)
        cnt = torch.zeros(max(view.n_tiles, 1), dtype=torch.int32, device=dev)
        ws = (acc, cnt)
        hist = cm.max_iters + 2
        budget = cm.max_iters - 1
        iterations, errors = [], []
        t0 = time.perf_counter()
        for c0 in range(0, B, PB):
            nb = min(PB, B - c0)
            pp = cols[:, c0:c0 + nb].to(device=dev, dtype=dtype)
            if perm is not None:
                pp = pp[perm]
            norms = pp.abs().sum(dim=0, dtype=f64)                              # abstract_filters.py:52
            live = norms > 0
            pn = (pp.to(f64) / torch.where(live, norms, torch.ones_like(norms))).to(dtype)   # :55
            del pp
            zbuf = [torch.zeros((n, PB), dtype=dtype, device=dev), torch.zeros((n, PB), dtype=dtype, device=dev)]
            q = torch.zeros((n, PB), dtype=dtype, device=dev)
            zbuf[0][:, :nb] = pn / sq[:, None]
            qc = coefvec.to(f64)[:, None] if coefvec is not None else float(coef)
            q[:, :nb] = (qc * pn.to(f64)).to(dtype) / sq[:, None]
            del pn
            tacc = (zbuf[0].to(f64) * c.to(f64)[:, None]).sum(dim=0)
            bias = (q.to(f64) * sq.to(f64)[:, None]).sum(dim=0)
            sf = torch.zeros((PB, C.STATE_LEN), dtype=f64, device=dev)
            sf[:, C.SF_ALPHA] = float(alpha_s)
            sf[:, C.SF_BIAS] = bias
            sf[:, C.SF_INVS] = 1.0 / (float(alpha_s) * tacc + bias) if self.use_quotient else 1.0
            sf[:, C.SF_TOL] = tol
            sf[:, C.SF_MEAN] = 1.0 if err_code == C.ERR_L1 else float(n)
            sf[:nb, C.SF_NORM] = norms
            si_host = np.zeros((PB, C.STATE_LEN), dtype=np.int32)
            si_host[:, C.SI_MAX_ITERS] = cm.max_iters
            si_host[:, C.SI_END_MODULO] = max(cm.end_modulo, 1)
            si_host[:, C.SI_ERR_MODE] = err_code
            si_host[:, C.SI_QUOTIENT] = int(self.use_quotient)
            live_host = live.cpu().numpy()
            si_host[:, C.SI_STOP] = C.CONVERGED                                 # padding / zero columns never run
            si_host[:nb, C.SI_STOP] = np.where(live_host, C.RUNNING, C.CONVERGED)
            si = torch.from_numpy(np.concatenate([si_host.reshape(-1), np.zeros(1, np.int32)])).to(dev)
            err_hist = torch.zeros((PB, hist), dtype=f64, device=dev)
            C.count_launches(1)
            done, chunk = 0, max(self.chunk, 1)
            host = si_host
            while done < budget and (host[:, C.SI_STOP] == C.RUNNING).any():
                k = min(chunk, budget - done)
                C.check(lib.pgb_affine_steps_batched(ctypes.byref(cs), code, float(alpha), C.ptr(w), C.ptr(sq_arg),
                                                     C.ptr(c), C.ptr(q), C.ptr(zbuf[0]), C.ptr(zbuf[1]), 0, C.ptr(sf),
                                                     C.ptr(si), C.ptr(err_hist), hist, span_struct(ws), done + 1, k, st))
                C.count_launches(k)
                done += k
                host = si.cpu().numpy()[:PB * C.STATE_LEN].reshape(PB, C.STATE_LEN)
                chunk = min(chunk * 2, 64)
            steps = host[:, C.SI_STEPS]
            final = zbuf[int(steps.max()) & 1]
            scale = torch.where(live, norms, torch.ones_like(norms)) if self.preserve_norm else torch.ones_like(norms)
            res = final[:, :nb] * sq[:, None] * scale.to(dtype)[None, :]
            if perm is not None:
                out[perm, c0:c0 + nb] = res
            else:
                out[:, c0:c0 + nb] = res
            hist_host = None
            for j in range(nb):
                if not live_host[j]:                                            # abstract_filters.py:53-54
                    iterations.append(0)
                    errors.append(None)
                    continue
                stop, it = int(host[j, C.SI_STOP]), int(host[j, C.SI_ITERATION])
                if stop == C.RUNNING:                                           # max_iters <= 1
                    stop, it = C.MAX_ITERS, 1
                iterations.append(it)
                errors.append(err_hist[j, 1:int(steps[j]) + 1])
                if stop == C.MAX_ITERS and err_code != C.ERR_ITERS and cm.iter_exception is not None:
                    raise cm.iter_exception("Could not converge within " + str(cm.max_iters) + " iterations")
        cm.iterations = iterations
        cm.iteration = iterations[-1] if iterations else 0
        cm.errors = errors[-1] if errors else None
        cm.column_errors = errors
        cm.elapsed_time = time.perf_counter() - t0
        return out

    def _affine(self, g: DeviceGraph, p, norm, warm, alpha: float, alpha_s: float, w_run, c_run, coef, coefvec):
        lib = C.lib()
        dtype, code = self.dtype, dtype_code(self.dtype)
        dev, n = p.device, g.n
        st = C.stream_ptr()
        state_f64, state_i32, err_hist = self._new_state(g, norm, alpha_s, self.use_quotient)
        sq = g.vec("sq", dtype)
        c = c_run if c_run is not None else g.vec("c", dtype)
        view = g.in_view
        zbuf, q, ws = self._work_buffers(g, dtype, dev)
        C.check(lib.pgb_affine_init(n, code, C.ptr(p), C.ptr(warm), C.ptr(sq), C.ptr(c), float(coef), C.ptr(coefvec),
                                    C.ptr(g.perm), 0, C.ptr(zbuf[0]), C.ptr(q), C.ptr(state_f64), st))
        C.check(lib.pgb_affine_init_finish(C.ptr(state_f64), C.ptr(state_i32), st))
        C.count_launches(3)                                   # init, init_finish, final unscale
        cs = view.cstruct(dtype)
        symdeg = g.symdeg and w_run is None
        w = None if symdeg else (w_run if w_run is not None else g.vec("w", dtype))
        sq_arg = None if symdeg else sq

        def launch(first, count):
            C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, float(alpha), C.ptr(w), C.ptr(sq_arg), C.ptr(c),
                                         C.ptr(q), C.ptr(zbuf[0]), C.ptr(zbuf[1]), 0, C.ptr(state_f64),
                                         C.ptr(state_i32), C.ptr(err_hist), span_struct(ws), first, count, 1, st))
            C.count_launches(count * view.kernels_per_step(dtype))

        steps = self._drive(launch, state_i32, err_hist)
        out = torch.empty(n, dtype=dtype, device=dev)
        C.check(lib.pgb_unscale(n, code, C.ptr(zbuf[steps & 1]), C.ptr(sq), None,
                                float(norm) if self.preserve_norm else 1.0, C.ptr(g.perm), C.ptr(out), st))
        return out


class PageRank(RecursiveGraphFilter):
    """Personalized PageRank power method (adhoc.py:10-45):
    ranks <- conv(ranks, M)*alpha + personalization*(1-alpha)."""

    def __init__(self, alpha: float = 0.85, *args, **kwargs):
        self.alpha = alpha
        super().__init__(*args, **kwargs)

    def _affine_args(self, g, **kwargs):
        return dict(alpha=self.alpha, alpha_s=self.alpha, w_run=None, c_run=None, coef=1 - self.alpha, coefvec=None)

    def sweep(self, graph, personalization, alphas: Sequence[float], **kwargs) -> torch.Tensor:
        """One solve per restart parameter in ``alphas`` for the same personalization — the candidates that
        ParameterTuner / optimize evaluate one after the other (autotune/parameterized.py:117-167,
        optimization.py:160-180) — as panels of the hub-blocked panel kernel: every column has its own alpha, normaliser,
        error and stop decision, the graph is streamed once per panel iteration instead of once per candidate.
        Returns scores [n, len(alphas)]; ``convergence.iterations`` holds each candidate's iteration count."""
        g = self._device_graph(graph)
        alphas = [float(a) for a in alphas]
        p, _ = _personalization(g, personalization, self.dtype)
        cols = p.unsqueeze(1).expand(g.n, len(alphas))        # no copy: the panels slice it
        if self._can_batch(g, n_columns=len(alphas), **kwargs) and self._panel_family(g) == "hsell":
            self._sweep = [(a, a, 1.0 - a) for a in alphas]
            try:
                return self._propagate_batched(g, cols, **kwargs)
            finally:
                self._sweep = None
        out = torch.empty((g.n, len(alphas)), dtype=self.dtype, device=p.device)
        iterations, keep = [], self.alpha
        try:
            for j, a in enumerate(alphas):
                self.alpha = a
                out[:, j] = self.rank(g, p, **kwargs).np
                iterations.append(self.convergence.iteration)
        finally:
            self.alpha = keep
        self.convergence.iterations = iterations
        return out


class AbsorbingWalks(RecursiveGraphFilter):
    """Partially absorbing random walks (adhoc.py:124-174):
    ranks <- (conv(ranks, M)*degrees + personalization*absorption) / (absorption + degrees)."""

    def __init__(self, alpha: float = 1 - 1.E-6, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.alpha = alpha

    def _affine_args(self, g, absorption=None, **kwargs):
        dtype = self.dtype
        f64 = torch.float64
        dev = g.out_view.indptr.device
        rate = (1 - self.alpha) / self.alpha                  # adhoc.py:158
        if absorption is None:
            ab = torch.full((g.n,), rate, dtype=f64, device=dev)
        else:
            ab_user, _ = _personalization(g, absorption, f64)
            ab = (ab_user if g.perm is None else ab_user[g.perm.long()]) * rate
        degM = g.vec("degM", f64)
        denom = ab + degM
        d1 = degM / denom
        w_run = (g.vec("w", f64) * d1).to(dtype)              # coefficient of the gathered sum
        coefvec = (ab / denom).to(dtype)                      # coefficient of the personalization
        gsum = g._spmv_raw(g.out_view, g.R * d1, g.L, f64)    # rowsum of M*diag(d1): next normaliser is linear
        c_run = (g.vec("sq", f64) * gsum).to(dtype)
        return dict(alpha=1.0, alpha_s=1.0, w_run=w_run, c_run=c_run, coef=0.0, coefvec=coefvec)


class _HostConvergence:
    """ConvergenceManager.has_converged (convergence.py:77-101) for the filters that run op by op on the device
    (Chebyshev / Krylov / cached powers): the counter, the max_iters rule, end_modulo and the error test, with the
    error itself one reduction on the device."""

    def __init__(self, cm: ConvergenceManager, n: int):
        self.cm, self.n = cm, n
        self.code = _error_code(cm.error_type)
        self.last = None
        cm.iteration = 0

    def has_converged(self, ranks: torch.Tensor) -> bool:
        cm = self.cm
        cm.iteration += 1
        if cm.iteration >= cm.max_iters:
            if self.code == C.ERR_ITERS or cm.iter_exception is None:
                return True
            raise cm.iter_exception("Could not converge within " + str(cm.max_iters) + " iterations")
        converged = False
        if self.last is not None and self.code != C.ERR_ITERS and cm.iteration % max(cm.end_modulo, 1) == 0:
            d = (self.last - ranks).to(torch.float64)
            if self.code == C.ERR_MABS:
                err = float(d.abs().sum()) / self.n
            elif self.code == C.ERR_L1:
                err = float(d.abs().sum())
            elif self.code == C.ERR_MSQ:
                err = float((d * d).sum()) / self.n
            else:
                err = float(d.abs().max())
            converged = err <= (0 if cm.tol is None else max(float(cm.tol), float(np.finfo(float).eps)))
        self.last = ranks
        return converged


def _obj_key(obj) -> str:
    """obj2id of the reference (core/utils/preprocessing.py:166-171): a uuid attached to the personalization object
    when it accepts attributes, else its identity."""
    import uuid
    if isinstance(obj, RankResult):
        obj = obj.np
    try:
        if not hasattr(obj, "uuid"):
            obj.uuid = uuid.uuid1()
        return str(obj.uuid)
    except (AttributeError, TypeError):
        return "id" + str(id(obj))


class ClosedFormGraphFilter(GraphFilter):
    """Polynomial filters (abstract_filters.py:139-267).  Taylor coefficients in the node space without a power cache —
    HeatKernel / PageRankClosed / GenericGraphFilter as the BASELINE configs use them — run fused on pgb_poly_steps.
    ``coefficient_type="chebyshev"``, ``krylov_dims`` (Lanczos, filters/krylov_space.py:16-50) and
    ``optimization_dict`` (the power cache tuners rely on, abstract_filters.py:241-246: O(n) per candidate once the
    powers of a personalization are known) run on the device op by op: one gather kernel per ``conv`` plus torch
    elementwise passes, with the reference's recursion, iteration counting and exceptions."""

    def __init__(self, krylov_dims=None, coefficient_type: str = "taylor", optimization_dict=None, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.krylov_dims = krylov_dims
        self.coefficient_type = coefficient_type.lower()
        self.optimization_dict = optimization_dict
        if self.coefficient_type not in ("taylor", "chebyshev"):
            raise Exception("Invalid coefficient type")
        self._active_dict = None

    def rank(self, graph=None, personalization=None, *args, **kwargs):
        if self.optimization_dict is not None:                 # abstract_filters.py:230-236
            key = _obj_key(personalization)
            self._active_dict = self.optimization_dict.setdefault(key, dict())
        else:
            self._active_dict = None
        return super().rank(graph, personalization, *args, **kwargs)

    def _fusable(self) -> bool:
        return self.krylov_dims is None and self.coefficient_type == "taylor" and self.optimization_dict is None

    # -- op-by-op evaluation on the device ------------------------------------------------------------------------
    def _retrieve_power(self, power, g, iteration):
        nxt = lambda: g.conv(power) if self.krylov_dims is None else power @ self._krylov_H
        if self._active_dict is not None:
            if iteration not in self._active_dict:
                self._active_dict[iteration] = nxt()
            return self._active_dict[iteration]
        return nxt()

    def _recursion(self, result, next_term, coef, iteration):
        if self.coefficient_type == "chebyshev":               # abstract_filters.py:206-217
            if iteration == 2:
                self._prev_term = next_term
            if iteration > 2:
                next_term = 2 * next_term - self._prev_term
                self._prev_term = next_term
                if coef == 0:
                    return result, next_term
            return result + next_term * coef, next_term
        if coef == 0:
            return result, next_term
        return result + next_term * coef, next_term

    def _run_eager(self, g, p, norm, warm):
        f64 = torch.float64
        dtype = self.dtype
        pn = (p.to(f64) / norm).to(dtype)
        conv_host = _HostConvergence(self.convergence, g.n)
        coef = None
        self._prev_term = 0
        ranks = torch.zeros(g.n, dtype=dtype, device=p.device)      # abstract_filters.py:213
        if self.krylov_dims is not None:
            ranks = pn.clone()                                      # the Krylov branch of _start keeps the start vector
            K = int(self.krylov_dims)
            V, H = self._krylov_base(g, pn, K)
            self._krylov_H = H
            result = H * 0
            power = torch.eye(K, dtype=dtype, device=p.device)
            bound = self._krylov_error_bound(V, H, g, pn)
            if bound > 0.01:
                raise Exception("Krylov approximation with estimated relative error " + str(bound)
                                + " > 0.01 is too rough to be meaningful (try on lager graphs)")
        else:
            power = pn
        while not conv_host.has_converged(ranks):
            it = self.convergence.iteration
            coef = self._coefficient(coef, it)
            if self.krylov_dims is not None:
                result, power = self._recursion(result, power, coef, it)
                ranks = (V @ result)[:, 0].contiguous()        # krylov2original (krylov_space.py:47-50)
            else:
                ranks, power = self._recursion(ranks, power, coef, it)
            power = self._retrieve_power(power, g, it)
        C.count_launches(self.convergence.iteration)
        return ranks * norm if self.preserve_norm else ranks

    def _krylov_base(self, g, pn, K):
        """Lanczos basis of the Krylov space of the (symmetric) operator (krylov_space.py:16-41)."""
        base = [pn / torch.dot(pn, pn) ** 0.5]
        norms, alphas = [], []
        for j in range(K):
            v = base[j]
            w = g.conv(v)
            a = torch.dot(v, w)
            alphas.append(a)
            nw = w - a * v
            if j > 0:
                nw = nw - base[j - 1] * norms[j - 1]
            nrm = (nw ** 2).sum() ** 0.5
            norms.append(nrm)
            if j != K - 1:
                base.append(nw / nrm)
        H = torch.diag(torch.stack(alphas))
        if K > 1:
            off = torch.stack(norms[1:])
            # the reference places base_norms[1:] on both off-diagonals (krylov_space.py:39)
            H = H + torch.diag(off, -1) + torch.diag(off, 1)
        return torch.stack(base, dim=1), H

    def _krylov_error_bound(self, V, H, g, pn, max_powers: int = 1) -> float:
        x = pn / torch.dot(pn, pn) ** 0.5
        res = torch.eye(V.shape[1], dtype=V.dtype, device=V.device)
        errors = []
        for power in range(max_powers + 1):
            approx = (V @ res)[:, 0]
            errors.append(float((x - approx).abs().sum()) / g.n)      # Mabs (supervised.py:101-106)
            if power < max_powers:
                res = res @ H
                x = g.conv(x)
        return max(errors)

    def _coefficient_table(self):
        """coef[k] = the coefficient step k applies (step k runs with convergence.iteration == k); coef[0] unused."""
        coefs, prev = [0.0], None
        for k in range(1, max(self.convergence.max_iters, 1) + 1):
            prev = self._coefficient(prev, k)
            coefs.append(float(prev))
        return coefs

    def _can_batch(self, g, warm_start=None, graph_dropout: float = 0, n_columns: int = 2, **kwargs) -> bool:
        # taylor filters in the node space: columns as slots of the hub-blocked panel kernel (polynomial mode)
        if not self._fusable() or g.in_view.weighted or warm_start is not None or graph_dropout != 0 or g.pathological:
            return False
        if self._panel_family(g) != "hsell" or self.convergence.max_iters <= 1:
            return False
        return n_columns >= 2 or os.environ.get("PGB_PANEL") not in (None, "")

    def _propagate_batched(self, g: DeviceGraph, cols: torch.Tensor, **kwargs) -> torch.Tensor:
        kwargs.pop("n_columns", None)
        return self._propagate_panels(g, cols, poly_coefs=self._coefficient_table(), **kwargs)

    def _coefficient(self, previous_coefficient, iteration: int) -> float:
        raise Exception("Use a derived class of ClosedFormGraphFilter that implements the _coefficient method")

    def _run(self, g, p, norm, warm, **kwargs):
        if not self._fusable():
            return self._run_eager(g, p, norm, warm)
        lib = C.lib()
        dtype, code = self.dtype, dtype_code(self.dtype)
        dev, n = p.device, g.n
        st = C.stream_ptr()
        cm = self.convergence
        coef_dev = torch.tensor(self._coefficient_table(), dtype=torch.float64, device=dev)
        state_f64, state_i32, err_hist = self._new_state(g, norm, 1.0, False)
        sq = g.vec("sq", dtype)
        zbuf = [torch.empty(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)]
        ranks = torch.zeros(n, dtype=dtype, device=dev)       # abstract_filters.py:213
        C.check(lib.pgb_affine_init(n, code, C.ptr(p), None, C.ptr(sq), None, 0.0, None, C.ptr(g.perm), 0,
                                    C.ptr(zbuf[0]), None, C.ptr(state_f64), st))   # power = personalization (:212)
        C.check(lib.pgb_affine_init_finish(C.ptr(state_f64), C.ptr(state_i32), st))
        view = g.in_view
        cs = view.cstruct(dtype)
        ws = view.new_span_ws(dtype)
        symdeg = g.symdeg
        w = None if symdeg else g.vec("w", dtype)
        sq_arg = None if symdeg else sq

        def launch(first, count):
            C.check(lib.pgb_poly_steps(ctypes.byref(cs), code, C.ptr(w), C.ptr(sq_arg), C.ptr(coef_dev), C.ptr(ranks),
                                       C.ptr(zbuf[0]), C.ptr(zbuf[1]), 0, C.ptr(state_f64), C.ptr(state_i32),
                                       C.ptr(err_hist), span_struct(ws), first, count, 1, st))
            C.count_launches(count * view.kernels_per_step(dtype))

        C.count_launches(3)
        self._drive(launch, state_i32, err_hist)
        out = torch.empty(n, dtype=dtype, device=dev)
        C.check(lib.pgb_unscale(n, code, C.ptr(ranks), None, None, float(norm) if self.preserve_norm else 1.0,
                                C.ptr(g.perm), C.ptr(out), st))
        return out


class HeatKernel(ClosedFormGraphFilter):
    """adhoc.py:95-121 — coefficient 1, then previous*t/(iteration+1)."""

    def __init__(self, t: float = 3, *args, **kwargs):
        self.t = t
        super().__init__(*args, **kwargs)

    def _coefficient(self, previous_coefficient, iteration):
        return 1. if previous_coefficient is None else previous_coefficient * self.t / (iteration + 1)


class PageRankClosed(ClosedFormGraphFilter):
    """adhoc.py:62-92 — coefficient 1, then previous*alpha."""

    def __init__(self, alpha: float = 0.85, *args, **kwargs):
        self.alpha = alpha
        super().__init__(*args, **kwargs)

    def _coefficient(self, previous_coefficient, iteration):
        return 1. if previous_coefficient is None else previous_coefficient * self.alpha


class GenericGraphFilter(ClosedFormGraphFilter):
    """low_pass.py:5-26 — coefficient weights[iteration-1], 0 past the end."""

    def __init__(self, weights: Optional[Sequence[float]] = None, **kwargs):
        super().__init__(**kwargs)
        self.weights = list(weights) if weights is not None else [0.9] * 10

    def _coefficient(self, _, iteration):
        if iteration > len(self.weights):
            return 0
        return self.weights[iteration - 1]
