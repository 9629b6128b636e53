#!/usr/bin/env python
"""bench.py — PPR GTEPS on synthetic RMAT graphs (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale S]

* N=1 workload (BASELINE.json north_star): personalized PageRank, alpha 0.85, tol 1e-9, symmetric
  normalisation, fp32 vectors, RMAT scale 24 (edge factor 16, a/b/c/d=.57/.19/.19/.05, symmetrised, self
  loops dropped, duplicates collapsed).  N>1 (BASELINE config 4): weak scaling, PageRank alpha 0.9, RMAT scale
  24+log2(N) row-partitioned across the N GPUs (pygrank_b200/dist.py).
* one "step" = one full PPR solve from a 10-seed personalization, run to convergence with the reference's
  ConvergenceManager semantics.  TEPS = nnz x conv calls / seconds.
* `value`    : personalization vectors already resident in HBM, result left on the device (fp32).
  `value_f64`/`roofline_f64`: the same solves in fp64 (the parity mode).
* `e2e`      : the same solves through the public API from HOST seed lists, the full score vector copied back to
               pinned host memory inside the timed region.  `e2e_plugin`: the same solves issued by the UNMODIFIED
               reference (`pg.PageRank` under `pg.Backend("b200")`, baseline/_ref) when it is installed.
* `roofline` : the fused per-iteration step timed alone with CUDA events on the launch stream; algorithmic bytes =
               nnz*4 + (n+1)*4 + 5*n*w over MEASURED_PEAKS.json hbm_gbs.
* `cpu_baseline` / `--impl reference`: the REAL reference (pygrank from baseline/_ref, numpy backend,
               pg.AdjacencyWrapper + pre-hashed preprocessor) on a bounded sample of the workload
               (same recipe at RMAT scale 22; the oracle port only if the reference cannot be imported).  Its path
               is serial (scipy csc_matvec + numpy passes): cores = 1.
* `parity`   : the GPU engine against that CPU run on the SAME graph and seeds, in the same process: iteration
               counts and relative L1 of the scores, fp64 and fp32.  A failed check exits non-zero.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOL, MAX_ITERS = 1e-9, 1000
CPU_SAMPLE_SCALE = 22


def alpha_for(n_gpus):
    return 0.85 if n_gpus == 1 else 0.9   # BASELINE configs: PPR 0.85 (north_star target) / C4 PageRank 0.9 on scale 27


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks/throttle reasons while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax = float(parts[1])
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference itself on the box's host cores
# ------------------------------------------------------------------------------------------------------
def import_reference():
    """The unmodified pygrank installed in baseline/_ref (never /root/reference at run time), or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "pygrank")):
        return None
    os.environ.setdefault("pygrankBackend", "numpy")
    import tempfile
    if "PYGRANK_KEEP_HOME" not in os.environ:
        os.environ["HOME"] = tempfile.mkdtemp()
    for p in (os.path.join(ROOT, "baseline", "stubs"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import pygrank
        return pygrank
    except Exception:
        return None


class CpuReference:
    """PPR solves by the reference's own CPU path on an RMAT sample of the bench recipe."""

    def __init__(self, scale, alpha):
        from oracle import fastgen
        self.scale, self.alpha = scale, alpha
        t0 = time.perf_counter()
        self.A = fastgen.rmat_graph(scale, 16, seed=1)
        self.n, self.nnz = self.A.shape[0], self.A.nnz
        self.pg = import_reference()
        if self.pg is not None:
            pg = self.pg
            pg.load_backend("numpy")
            self.kind = "reference"
            self.G = pg.AdjacencyWrapper(self.A, directed=False)
            self.pre = pg.preprocessor(normalization="symmetric", assume_immutability=True)
            self.pre(self.G)                                     # pre-hashed, documentation.md:382-385
            self.alg = pg.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, preprocessor=self.pre)
            import scipy
            self.versions = f"pygrank {getattr(pg, '__version__', '0.2.12')}, scipy {scipy.__version__}, numpy {np.__version__}"
        else:
            from oracle import reference_port as orc
            self.kind = "port"
            self.orc = orc
            self.M = orc.to_sparse_matrix(self.A, "symmetric", False)
            self.versions = "oracle/reference_port.py (baseline/_ref not importable)"
        self.setup_s = time.perf_counter() - t0

    def solve(self, seeds):
        """(scores fp64[n], ConvergenceManager.iteration, seconds)."""
        p = np.zeros(self.n)
        p[seeds] = 1.0
        t0 = time.perf_counter()
        if self.kind == "reference":
            r = self.alg(self.pg.to_signal(self.G, p))
            out, iters = np.asarray(r.np, dtype=np.float64), int(self.alg.convergence.iteration)
        else:
            out, iters, _ = self.orc.pagerank(self.M, p, self.alpha, tol=TOL, max_iters=MAX_ITERS)
        return out, iters, time.perf_counter() - t0

    def run(self, steps, warmup):
        from pygrank_b200 import synthetic
        seeds = synthetic.seed_sets(self.n, warmup + steps, 10, seed=0)
        calls, elapsed, results = 0, 0.0, []
        for i, s in enumerate(seeds):
            out, iters, dt = self.solve(s)
            if i >= warmup:
                calls += iters - 1
                elapsed += dt
                results.append((s, out, iters))
        gteps = self.nnz * calls / elapsed / 1e9
        sample = (f"{self.kind}: {self.versions}; RMAT scale {self.scale} (n={self.n}, nnz={self.nnz}), {steps} PPR solves "
                  f"alpha={self.alpha} tol={TOL} fp64, {calls} conv calls in {elapsed:.2f}s after {warmup} warm-up solves "
                  f"(graph + preprocessing {self.setup_s:.1f}s untimed)")
        return gteps, elapsed / max(steps, 1) * 1e3, sample, results


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    alpha = alpha_for(args.gpus)
    cpu = CpuReference(CPU_SAMPLE_SCALE, alpha)
    gteps, ms, sample, _ = cpu.run(args.steps, args.warmup)
    cfg = workload_config(args.gpus, args.scale, alpha)
    cfg.update(sample_rmat_scale=cpu.scale, sample_n=cpu.n, sample_nnz=cpu.nnz,
               sample_note="bounded sample of the workload: same recipe at a scale the serial CPU path finishes in minutes")
    line = {
        "impl": "reference", "metric": "PPR GTEPS", "value": gteps, "unit": "GTEPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": gteps, "unit": "GTEPS", "cores": 1, "kind": cpu.kind, "sample": sample},
        "e2e": {"value": gteps, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus, scale, alpha):
    return {"workload": f"PPR alpha={alpha} tol={TOL} symmetric-normalised, RMAT scale {scale} ef16 "
                        f"(a,b,c,d=.57,.19,.19,.05) symmetrised/deduped, 10-seed personalization per solve",
            "rmat_scale": scale, "partition": "single" if n_gpus == 1 else f"rows x{n_gpus}, exchange fused into the step",
            "l2": "inputs larger than L2 (CSR indices >> 126 MB)"}


# ------------------------------------------------------------------------------------------------------
# GPU arm, one GPU
# ------------------------------------------------------------------------------------------------------
def parity_check(cpu, results, relabel):
    """GPU engine vs the CPU reference run on the same graph and seeds (this process, this box)."""
    import torch

    import pygrank_b200 as pgb
    from pygrank_b200 import device_synthetic
    g = device_synthetic.rmat_graph_device(cpu.scale, 16, seed=1, normalization="symmetric", relabel=relabel)
    assert g.n == cpu.n and g.nnz == cpu.nnz, "device and host generators disagree"
    out = {"against": cpu.kind, "rmat_scale": cpu.scale, "n": cpu.n, "nnz": cpu.nnz, "solves": len(results), "ok": True}
    for name, dtype, tol in (("f64", torch.float64, 1e-10), ("f32", torch.float32, 1e-5)):
        alg = pgb.PageRank(cpu.alpha, tol=TOL, max_iters=MAX_ITERS, dtype=dtype)
        worst, its_gpu, its_cpu = 0.0, [], []
        for seeds, ref, iters in results:
            got = alg(g, [int(v) for v in seeds]).numpy().astype(np.float64)
            worst = max(worst, float(np.abs(got - ref).sum() / np.abs(ref).sum()))
            its_gpu.append(int(alg.convergence.iteration))
            its_cpu.append(int(iters))
        same = its_gpu == its_cpu
        ok = bool(np.isfinite(worst)) and worst <= tol and (same if name == "f64" else
                                                             all(abs(a - b) <= 1 for a, b in zip(its_gpu, its_cpu)))
        out[name] = {"rel_l1": worst, "tolerance": tol, "iterations": its_gpu, "iterations_cpu": its_cpu,
                     "iterations_equal": same, "ok": ok}
        out["ok"] = out["ok"] and ok
    hs = g.in_view.hsell(torch.float32)
    out["hsell"] = None if hs is None else {"block_cols": hs.block_cols, "n_blocks": hs.n_blocks}
    del g
    torch.cuda.empty_cache()
    return out


def device_leg(args, g, pers, alg, n_warm, total):
    import torch

    from pygrank_b200 import _capi as C
    for i in range(n_warm):
        alg(g, pers[i])
    torch.cuda.synchronize()
    launches0 = C.LAUNCHES[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    calls = 0
    ev0.record()
    for i in range(n_warm, total):
        alg(g, pers[i])
        calls += alg.convergence.iteration - 1
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    return ms, calls, C.LAUNCHES[0] - launches0


def plugin_leg(args, g, seeds, alpha, dtype, n_warm, total):
    """The same solves issued by the UNMODIFIED reference through the backend plugin: pg.PageRank under
    pg.Backend("b200") with the device preprocessor injected, host seed dicts in, scores to pinned host memory."""
    import torch

    import pygrank_b200 as pgb
    pg = import_reference()
    if pg is None:
        return None
    try:
        pgb.install(pg)
        from pygrank_b200 import backend as b200
        b200.configure(dtype=dtype)
        hosts = [torch.empty(g.n, dtype=dtype).pin_memory(), torch.empty(g.n, dtype=dtype).pin_memory()]
        copy_stream = torch.cuda.Stream()
        keep = []
        with pg.Backend("b200"):
            pre = pgb.preprocessor(normalization="symmetric", assume_immutability=True)
            alg = pg.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, preprocessor=pre)

            dev = g.out_view.indptr.device
            idx_host = [torch.from_numpy(np.asarray(s, dtype=np.int64)).pin_memory() for s in seeds]

            def solve(i):
                # host seed list -> device personalization (the reference's own dict route would fill and upload a dense
                # n-vector on the host, core/signals.py:62-66: SURVEY §8 f2, not the path measured here)
                p = torch.zeros(g.n, dtype=dtype, device=dev)
                p[idx_host[i].to(dev, non_blocking=True)] = 1.0
                r = alg(pg.to_signal(g, p))
                out = b200.to_tensor(r.np)
                ready = torch.cuda.Event()
                ready.record()
                copy_stream.wait_event(ready)
                with torch.cuda.stream(copy_stream):      # as in the e2e leg: the copy of solve k overlaps solve k+1
                    hosts[i & 1].copy_(out, non_blocking=True)
                out.record_stream(copy_stream)
                keep.append(out)
                del keep[:-2]
                return alg.convergence.iteration - 1

            for i in range(n_warm):
                solve(i)
            torch.cuda.synchronize()
            calls = 0
            t0 = time.perf_counter()
            for i in range(n_warm, total):
                calls += solve(i)
            copy_stream.synchronize()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        pg.load_backend("numpy")
        b200.configure(dtype=torch.float64)
        return {"value": g.nnz * calls / dt / 1e9, "unit": "GTEPS", "conv_calls_per_solve": calls / (total - n_warm),
                "route": "unmodified pg.PageRank (baseline/_ref) under pg.Backend('b200'), device preprocessor injected, "
                         "personalization built on the device from the host seed list",
                "h2d_bytes_per_step": 160, "d2h_bytes_per_step": g.n * (4 if dtype == torch.float32 else 8)}
    except Exception as exc:   # reported, never fatal for the bench line
        return {"value": None, "error": repr(exc)[:300]}


def propagate_leg(args, g, dtype, alpha, columns=32):
    """NodeRanking.propagate (core/signals.py:225-226) of `columns` seed sets on the bench graph through the hub-blocked
    panel path (pgb_affine_steps_panel: 4 fp32 / 2 fp64 columns per gather, columns scheduled over the panel's slots on the
    device), checked column by column against single solves of the same engine."""
    import torch

    import pygrank_b200 as pgb
    from pygrank_b200 import synthetic
    try:
        dev = g.out_view.indptr.device
        n = g.n
        P = torch.zeros((n, columns), dtype=dtype, device=dev)
        for c, sset in enumerate(synthetic.seed_sets(n, columns, 10, seed=7)):
            P[torch.from_numpy(sset).to(dev), c] = 1.0
        alg = pgb.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, dtype=dtype)
        alg.propagate(g, P[:, :8])
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        t0 = time.perf_counter()
        out = alg.propagate(g, P)
        ev1.record()
        torch.cuda.synchronize()
        secs = max(ev0.elapsed_time(ev1) * 1e-3, time.perf_counter() - t0)
        its = list(alg.convergence.iterations)
        calls = sum(i - 1 for i in its)
        one = pgb.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, dtype=dtype)
        worst, its_off = 0.0, 0
        t1 = time.perf_counter()
        for c in range(columns):
            ref = one(g, P[:, c].contiguous()).np
            worst = max(worst, float((out[:, c] - ref).abs().sum(dtype=torch.float64) / ref.abs().sum(dtype=torch.float64)))
            its_off = max(its_off, abs(one.convergence.iteration - its[c]))
        torch.cuda.synchronize()
        seq = time.perf_counter() - t1
        tol = 1e-5 if dtype == torch.float32 else 1e-10
        return {"columns": columns, "seconds": secs, "edge_column_gteps": g.nnz * calls / secs / 1e9,
                "column_iterations_min_max": [min(its), max(its)],
                "route": "pgb_affine_steps_panel (hub-blocked panel kernel, device-scheduled slots)"
                         if alg._panel_family(g) == "hsell" else "item-stream panel kernel",
                "column_by_column_seconds_incl_check": seq,
                "parity": {"vs": "single solves of the same engine, every column", "worst_rel_l1": worst, "tol": tol,
                           "iterations_max_abs_diff": its_off, "ok": bool(worst <= tol and its_off <= (1 if dtype == torch.float32 else 0))}}
    except Exception as exc:   # reported, never fatal for the bench line
        return {"edge_column_gteps": None, "error": repr(exc)[:300]}


def run_single(args):
    import torch

    import pygrank_b200 as pgb
    from pygrank_b200 import device_synthetic, synthetic

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    alpha = alpha_for(1)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if dtype == torch.float32 else 8
    scale = args.scale
    t0 = time.perf_counter()
    g = device_synthetic.rmat_graph_device(scale, 16, seed=1, normalization="symmetric", relabel=args.relabel)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    n, nnz = g.n, g.nnz
    total = args.warmup + args.steps
    seeds = synthetic.seed_sets(n, total, 10, seed=0)
    alg = pgb.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, dtype=dtype)

    pers = []
    for s in seeds:
        p = torch.zeros(n, dtype=dtype, device=dev)
        p[torch.from_numpy(s).to(dev)] = 1.0
        pers.append(p)
    if args.kernel_only:
        return kernel_leg(args, g, pers[0], dtype, alpha)
    sampler = ClockSampler(0)
    sampler.start()
    # ---- device-resident leg -------------------------------------------------------------
    dev_ms, conv_calls, launches = device_leg(args, g, pers, alg, args.warmup, total)
    value = nnz * conv_calls / (dev_ms * 1e-3) / 1e9

    # ---- end-to-end leg: host seed lists in, full score vector out to pinned host memory -----
    # Every solve's scores go to pinned host memory inside the timed region; the copy of solve k runs on a
    # side stream while solve k+1 computes (two host buffers), as a serving loop would do it.
    host_out = [torch.empty(n, dtype=dtype).pin_memory(), torch.empty(n, dtype=dtype).pin_memory()]
    copy_stream = torch.cuda.Stream()
    seed_lists = [[int(v) for v in s] for s in seeds]

    def solve_to_host(i, keep):
        r = alg(g, seed_lists[i])
        ready = torch.cuda.Event()
        ready.record()
        copy_stream.wait_event(ready)
        with torch.cuda.stream(copy_stream):
            host_out[i & 1].copy_(r.np, non_blocking=True)
        r.np.record_stream(copy_stream)
        keep.append(r)
        del keep[:-2]
        return alg.convergence.iteration - 1

    keep = []
    for i in range(args.warmup):
        solve_to_host(i, keep)
    torch.cuda.synchronize()
    e2e_calls = 0
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    t1 = time.perf_counter()
    for i in range(args.warmup, total):
        e2e_calls += solve_to_host(i, keep)
    copy_stream.synchronize()
    ev3.record()
    torch.cuda.synchronize()
    e2e_s = max(ev2.elapsed_time(ev3) * 1e-3, time.perf_counter() - t1)
    e2e_value = nnz * e2e_calls / e2e_s / 1e9
    clocks = sampler.stop()
    del keep

    kern = kernel_leg(args, g, pers[0], dtype, alpha)
    peak, peak_src = hbm_peak()

    # ---- the other precision on the same workload (fp64 is the parity mode) ---------------------------
    other = torch.float64 if dtype == torch.float32 else torch.float32
    oname = "f64" if dtype == torch.float32 else "f32"
    alg_o = pgb.PageRank(alpha, tol=TOL, max_iters=MAX_ITERS, dtype=other)
    pers_o = [p.to(other) for p in pers]
    o_ms, o_calls, _ = device_leg(args, g, pers_o, alg_o, min(args.warmup, 2), total)
    o_value = nnz * o_calls / (o_ms * 1e-3) / 1e9
    kern_o = kernel_leg(args, g, pers_o[0], other, alpha, probe=False)
    del pers_o

    e2e_plugin = plugin_leg(args, g, seeds, alpha, dtype, min(args.warmup, 2), total) if not args.no_plugin else None

    panel = propagate_leg(args, g, dtype, alpha) if not args.no_panel else None

    cpu = parity = None
    if not args.no_cpu:
        ref = CpuReference(CPU_SAMPLE_SCALE, alpha)
        gt, _, sample, results = ref.run(2, 1)
        cpu = {"value": gt, "unit": "GTEPS", "cores": 1, "kind": ref.kind, "sample": sample}
        parity = parity_check(ref, results, args.relabel)

    form = kern["form"]

    def roof(k, dt_name):
        return {"bound": "hbm", "achieved": k["achieved"], "peak": peak, "unit": "GB/s", "frac": k["achieved"] / peak,
                "traffic": measured_traffic(scale, dt_name, k["form"]),
                "kernel": ("one fused step = hsell_gather_kernel<%s> (dominant) + hsell_update_%skernel<%s,AFFINE,SYMDEG=%s>"
                           % (dt_name, "accum_" if k["accumulate"] else "", dt_name, k["symdeg"])
                           if k["form"] is not None else "item_stream_kernel<%s,unweighted,AFFINE,SYMDEG=%s>" % (dt_name, k["symdeg"])),
                "kernel_ms": k["kernel_ms"], "kernel_gteps": nnz / (k["kernel_ms"] * 1e-3) / 1e9,
                "algorithmic_bytes": k["alg_bytes"], "peak_source": peak_src}

    r_main = roof(kern, args.dtype)
    r_main.update({"gather_probe_ms": kern["probe_ms"],
                   "hsell": None if form is None else {
                       "block_cols": form.block_cols, "n_blocks": form.n_blocks, "hub_chunks": form.n_hub_chunks,
                       "tail_chunks": form.n_tail_chunks, "hub_slots": form.n_hub_words * 2,
                       "tail_slots": form.n_tail_words, "pieces": form.n_pieces, "bytes": form.nbytes(),
                       "tail_gathers": "TEX pipe",
                       "pieces_flushed_by": "RED.ADD into y" if kern["accumulate"] else "partial rows"}})
    line = {
        "metric": "PPR GTEPS", "value": value, "unit": "GTEPS", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": dict(workload_config(1, scale, alpha), n=n, nnz=nnz, conv_calls_per_solve=conv_calls / args.steps,
                       relabel=args.relabel, graph_build_s=round(build_s, 2)),
        "e2e": {"value": e2e_value, "unit": "GTEPS", "h2d_bytes_per_step": 10 * 8 + 10 * 8,
                "d2h_bytes_per_step": n * w + 64 * 2},
        "e2e_plugin": e2e_plugin,
        "propagate": panel,
        "gpu_launches": launches,
        "roofline": r_main,
        "value_" + oname: o_value, "roofline_" + oname: roof(kern_o, oname),
        "cpu_baseline": cpu,
        "parity": parity,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if parity is not None and not parity["ok"]:
        sys.exit(3)


def kernel_leg(args, g, pers0, dtype, alpha, probe=True):
    """The dominant kernels alone: fixed-iteration fused PPR steps, CUDA events on the launch stream."""
    import ctypes

    import torch

    from pygrank_b200 import _capi as C
    from pygrank_b200.graph import dtype_code, hsell_config, span_struct
    dev = pers0.device
    n, nnz = g.n, g.nnz
    w = 4 if dtype == torch.float32 else 8
    lib = C.lib()
    code = dtype_code(dtype)
    st = C.stream_ptr()
    reps = 30
    state_f64 = torch.zeros(C.STATE_LEN, dtype=torch.float64, device=dev)
    state_i32 = torch.zeros(C.STATE_LEN, dtype=torch.int32, device=dev)
    sf = [0.0] * C.STATE_LEN
    si = [0] * C.STATE_LEN
    sf[C.SF_ALPHA], sf[C.SF_INVS], sf[C.SF_MEAN], sf[C.SF_NORM] = alpha, 1.0, float(n), 10.0
    si[C.SI_MAX_ITERS], si[C.SI_END_MODULO], si[C.SI_ERR_MODE], si[C.SI_QUOTIENT] = 10 ** 6, 1, C.ERR_ITERS, 1
    state_f64.copy_(torch.tensor(sf, dtype=torch.float64))
    state_i32.copy_(torch.tensor(si, dtype=torch.int32))
    sq, cvec = g.vec("sq", dtype), g.vec("c", dtype)
    zbuf = [torch.empty(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)]
    q = torch.empty(n, dtype=dtype, device=dev)
    C.check(lib.pgb_affine_init(n, code, C.ptr(pers0), None, C.ptr(sq), C.ptr(cvec), 1 - alpha, None, C.ptr(g.perm), 0,
                                C.ptr(zbuf[0]), C.ptr(q), C.ptr(state_f64), st))
    C.check(lib.pgb_affine_init_finish(C.ptr(state_f64), C.ptr(state_i32), st))
    cs = g.in_view.cstruct(dtype)
    ws = g.in_view.new_span_ws(dtype)
    form = g.in_view.hsell(dtype)
    symdeg = g.symdeg
    wv = None if symdeg else g.vec("w", dtype)
    sqa = None if symdeg else sq
    err_hist = torch.zeros(reps * 2 + 16, dtype=torch.float64, device=dev)

    def steps(first, count):
        C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, alpha, C.ptr(wv), C.ptr(sqa), C.ptr(cvec), C.ptr(q),
                                     C.ptr(zbuf[0]), C.ptr(zbuf[1]), 0, C.ptr(state_f64), C.ptr(state_i32),
                                     C.ptr(err_hist), span_struct(ws), first, count, 1, st))

    steps(1, 5)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    steps(6, reps)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = k0.elapsed_time(k1) / reps
    probe_ms = None
    if probe:   # gather ceiling of the plain CSR form of this graph: index stream + gathers only (pgb_gather_probe)
        scratch = torch.zeros(8, dtype=dtype, device=dev)
        for _ in range(3):
            C.check(lib.pgb_gather_probe(ctypes.byref(cs), code, C.ptr(zbuf[0]), C.ptr(scratch), st))
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(10):
            C.check(lib.pgb_gather_probe(ctypes.byref(cs), code, C.ptr(zbuf[0]), C.ptr(scratch), st))
        p1.record()
        torch.cuda.synchronize()
        probe_ms = p0.elapsed_time(p1) / 10
    alg_bytes = nnz * 4 + (n + 1) * 4 + 5 * n * w
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    out = {"kernel_ms": kernel_ms, "probe_ms": probe_ms, "alg_bytes": alg_bytes, "achieved": achieved, "form": form,
           "symdeg": symdeg, "accumulate": form is not None and hsell_config()["accumulate"]}
    if args.kernel_only:
        print(json.dumps({"kernel_ms": kernel_ms, "gteps": nnz / (kernel_ms * 1e-3) / 1e9, "achieved_gbs": achieved,
                          "probe_ms": probe_ms, "n": n, "nnz": nnz,
                          "hsell": None if form is None else {"block_cols": form.block_cols, "n_blocks": form.n_blocks,
                                                              "hub_chunks": form.n_hub_chunks,
                                                              "tail_chunks": form.n_tail_chunks,
                                                              "pieces": form.n_pieces}}))
    return out


def measured_traffic(scale, dtype, form):
    """DRAM bytes per fused step from the committed ncu --set full captures of this round (profiles/traffic_r2.json:
    gather + update kernels of the default bench command), valid for the workload they were taken on."""
    path = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if scale != 24 or form is None or not os.path.exists(path):
        return None
    try:
        per = json.load(open(path)).get(dtype, {}).get("bytes_per_launch", {})
        return sum(per.values()) if per else None
    except (ValueError, KeyError):
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=None, help="RMAT scale (default 24 + log2(gpus))")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--relabel", default="hub", choices=["hub", "degree", "none"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline + parity legs")
    ap.add_argument("--no-plugin", action="store_true", help="skip the e2e_plugin leg")
    ap.add_argument("--no-panel", action="store_true", help="skip the propagate (panel kernel) leg")
    ap.add_argument("--no-parity", action="store_true", help="N>1: skip the pre-timing parity solves")
    ap.add_argument("--kernel-only", action="store_true",
                    help="experiments: time only the fused step (no solves, no e2e, no CPU leg) and print a short line")
    args = ap.parse_args()
    if args.scale is None:
        args.scale = 24 + max(int(np.log2(max(args.gpus, 1))), 0)
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 or world > 1:
        from pygrank_b200 import dist_bench
        return dist_bench.run(args)
    return run_single(args)


if __name__ == "__main__":
    main()
