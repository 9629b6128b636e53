#!/usr/bin/env python
"""bench.py — PPR GTEPS on synthetic RMAT graphs (BASELINE.json metric), one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--scale S]

* N=1 workload (BASELINE.json configs / north_star): personalized PageRank, alpha 0.85, tol 1e-9,
  symmetric normalisation, fp32 vectors, RMAT scale 24 (edge factor 16, a/b/c/d=.57/.19/.19/.05,
  symmetrised, self loops dropped, duplicates collapsed).  N>1: weak scaling, RMAT scale
  24+log2(N) row-partitioned across the N GPUs with an NVLink all-gather of the rank slice per
  iteration (pygrank_b200/dist.py).
* one "step" = one full PPR solve from a 10-seed personalization, run to convergence with the
  reference's ConvergenceManager semantics.  TEPS = nnz x conv calls / seconds.
* `value`  : personalization vectors already resident in HBM, result left on the device.
* `e2e`    : the same solves through the public API from HOST seed lists, the full score vector
             copied back to pinned host memory inside the timed region.
* `roofline`: the fused per-iteration kernel timed alone with CUDA events on the launch stream;
             algorithmic bytes = nnz*4 + (n+1)*4 + 5*n*w over MEASURED_PEAKS.json hbm_gbs.
* `cpu_baseline`: the oracle (numpy/scipy port of the reference path, 1 core — the reference's
             scipy csc_matvec is serial) on a bounded sample (RMAT scale 20, same recipe).
* `--impl reference`: only the CPU leg, printed as its own JSON line.
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

ALPHA, TOL, MAX_ITERS = 0.85, 1e-9, 1000
CPU_SAMPLE_SCALE = 20


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


def cpu_reference_leg(steps, warmup, scale=CPU_SAMPLE_SCALE):
    """The reference's CPU path (oracle port: numpy + the serial csc_matvec scatter) on the sample."""
    from oracle import reference_port as orc
    from pygrank_b200 import synthetic
    t0 = time.perf_counter()
    A = synthetic.rmat_graph_host(scale, 16, seed=1)
    M = orc.to_sparse_matrix(A, "symmetric", False)
    n, nnz = A.shape[0], A.nnz
    seeds = synthetic.seed_sets(n, warmup + steps, 10, seed=0)
    setup = time.perf_counter() - t0
    conv_calls, elapsed = 0, 0.0
    for i, s in enumerate(seeds):
        p = np.zeros(n)
        p[s] = 1.0
        t1 = time.perf_counter()
        _, iters, _ = orc.pagerank(M, p, ALPHA, tol=TOL, max_iters=MAX_ITERS)
        dt = time.perf_counter() - t1
        if i >= warmup:
            conv_calls += iters - 1
            elapsed += dt
    gteps = nnz * conv_calls / elapsed / 1e9
    sample = (f"RMAT scale {scale} (n={n}, nnz={nnz}), {steps} PPR solves alpha={ALPHA} tol={TOL}, fp64, "
              f"{conv_calls} conv calls in {elapsed:.2f}s (graph setup {setup:.1f}s untimed)")
    return gteps, elapsed / max(steps, 1) * 1e3, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    gteps, ms, sample = cpu_reference_leg(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "PPR GTEPS", "value": gteps, "unit": "GTEPS", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.scale),
        "cpu_baseline": {"value": gteps, "unit": "GTEPS", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": gteps, "unit": "GTEPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(n_gpus, scale):
    return {"workload": f"PPR alpha={ALPHA} tol={TOL} symmetric-normalised, RMAT scale {scale} ef16 "
                        f"(a,b,c,d=.57,.19,.19,.05) symmetrised/deduped, 10-seed personalization per solve",
            "rmat_scale": scale, "partition": "single" if n_gpus == 1 else f"rows x{n_gpus} + allgather",
            "l2": "inputs larger than L2 (CSR indices >> 126 MB)"}


def run_single(args):
    import torch

    import pygrank_b200 as pgb
    from pygrank_b200 import _capi as C
    from pygrank_b200 import device_synthetic, synthetic
    from pygrank_b200.graph import dtype_code, span_struct
    import ctypes

    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
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
    alg = pgb.PageRank(ALPHA, tol=TOL, max_iters=MAX_ITERS, dtype=dtype)

    # ---- device-resident leg -------------------------------------------------------------
    pers = []
    for s in seeds:
        p = torch.zeros(n, dtype=dtype, device=dev)
        p[torch.from_numpy(s).to(dev)] = 1.0
        pers.append(p)
    if args.kernel_only:
        return kernel_leg(args, g, pers[0], dtype, None)
    for i in range(args.warmup):
        alg(g, pers[i])
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = C.LAUNCHES[0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    conv_calls = 0
    torch.cuda.synchronize()
    ev0.record()
    for i in range(args.warmup, total):
        alg(g, pers[i])
        conv_calls += alg.convergence.iteration - 1
    ev1.record()
    torch.cuda.synchronize()
    dev_ms = ev0.elapsed_time(ev1)
    launches = C.LAUNCHES[0] - launches0
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

    kern = kernel_leg(args, g, pers[0], dtype, None)
    kernel_ms, probe_ms, alg_bytes, achieved, form, symdeg = (kern[k] for k in
                                                              ("kernel_ms", "probe_ms", "alg_bytes", "achieved", "form", "symdeg"))
    peak, peak_src = hbm_peak()

    cpu = None
    if not args.no_cpu:
        gt, _, sample = cpu_reference_leg(2, 1)
        cpu = {"value": gt, "unit": "GTEPS", "cores": 1, "kind": "port", "sample": sample}
    return finish_line(args, locals())


def kernel_leg(args, g, pers0, dtype, _unused):
    """The dominant kernel alone: fixed-iteration fused PPR steps, CUDA events on the launch stream."""
    import ctypes

    import torch

    from pygrank_b200 import _capi as C
    from pygrank_b200.graph import dtype_code, span_struct
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
    sf[C.SF_ALPHA], sf[C.SF_INVS], sf[C.SF_MEAN], sf[C.SF_NORM] = ALPHA, 1.0, float(n), 10.0
    si[C.SI_MAX_ITERS], si[C.SI_END_MODULO], si[C.SI_ERR_MODE], si[C.SI_QUOTIENT] = 10 ** 6, 1, C.ERR_ITERS, 1
    state_f64.copy_(torch.tensor(sf, dtype=torch.float64))
    state_i32.copy_(torch.tensor(si, dtype=torch.int32))
    sq, cvec = g.vec("sq", dtype), g.vec("c", dtype)
    zbuf = [torch.empty(n, dtype=dtype, device=dev), torch.empty(n, dtype=dtype, device=dev)]
    q = torch.empty(n, dtype=dtype, device=dev)
    C.check(lib.pgb_affine_init(n, code, C.ptr(pers0), None, C.ptr(sq), C.ptr(cvec), 1 - ALPHA, None, C.ptr(g.perm), 0,
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
        C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, ALPHA, C.ptr(wv), C.ptr(sqa), C.ptr(cvec), C.ptr(q),
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
    # gather ceiling of this graph: index stream + gathers only (pgb_gather_probe)
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
           "symdeg": symdeg}
    if args.kernel_only:
        print(json.dumps({"kernel_ms": kernel_ms, "gteps": nnz / (kernel_ms * 1e-3) / 1e9, "achieved_gbs": achieved,
                          "probe_ms": probe_ms, "n": n, "nnz": nnz,
                          "hsell": None if form is None else {"block_cols": form.block_cols, "n_blocks": form.n_blocks,
                                                              "hub_chunks": form.n_hub_chunks,
                                                              "tail_chunks": form.n_tail_chunks,
                                                              "partial_rows": form.n_partials}}))
    return out


def measured_traffic(scale, dtype, form):
    """DRAM bytes per fused step from the committed ncu captures (profiles/traffic_r1.json: gather + update
    kernels; the reduce kernel moves < 1 % of that), valid for the workload they were taken on."""
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic_r1.json")
    if scale != 24 or dtype != "f32" or form is None or not os.path.exists(path):
        return None
    try:
        per = json.load(open(path))["bytes_per_launch"]
        parts = [v for k, v in per.items() if "hsell_gather_kernel" in k or "hsell_update_kernel" in k]
        return sum(parts) if len(parts) == 2 else None
    except (ValueError, KeyError):
        return None


def finish_line(args, v):
    n, nnz, w, scale = v["n"], v["nnz"], v["w"], v["scale"]
    value, dev_ms, conv_calls, build_s = v["value"], v["dev_ms"], v["conv_calls"], v["build_s"]
    e2e_value, launches, achieved, peak, peak_src = v["e2e_value"], v["launches"], v["achieved"], v["peak"], v["peak_src"]
    form, symdeg, kernel_ms, alg_bytes, probe_ms, cpu, clocks = (v["form"], v["symdeg"], v["kernel_ms"], v["alg_bytes"],
                                                                 v["probe_ms"], v["cpu"], v["clocks"])
    line = {
        "metric": "PPR GTEPS", "value": value, "unit": "GTEPS", "n_gpus": 1, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": dict(workload_config(1, scale), n=n, nnz=nnz, conv_calls_per_solve=conv_calls / args.steps,
                       relabel=args.relabel, graph_build_s=round(build_s, 2)),
        "e2e": {"value": e2e_value, "unit": "GTEPS", "h2d_bytes_per_step": 10 * 8 + 10 * 8,
                "d2h_bytes_per_step": n * w + 64 * 2},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": measured_traffic(scale, args.dtype, form),
                     "kernel": ("one fused step = hsell_gather_kernel<%s> (dominant, ~2/3 of the step) + hsell_reduce_kernel"
                                " + hsell_update_kernel<%s,AFFINE,SYMDEG=%s>"
                                if form is not None else "item_stream_kernel<%s,unweighted,AFFINE,SYMDEG=%s>")
                     % ((args.dtype, args.dtype, symdeg) if form is not None else (args.dtype, symdeg)),
                     "kernel_ms": kernel_ms, "kernel_gteps": nnz / (kernel_ms * 1e-3) / 1e9,
                     "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                     "gather_probe_ms": probe_ms, "gather_probe_gteps": nnz / (probe_ms * 1e-3) / 1e9,
                     "frac_of_gather_probe": probe_ms / kernel_ms,
                     "hsell": None if form is None else {
                         "block_cols": form.block_cols, "n_blocks": form.n_blocks, "hub_chunks": form.n_hub_chunks,
                         "tail_chunks": form.n_tail_chunks, "hub_slots": form.n_hub_words * 2,
                         "tail_slots": form.n_tail_words, "partial_rows": form.n_partials,
                         "heavy_slices": form.n_heavy, "bytes": form.nbytes()}},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=int, default=None, help="RMAT scale (default 24 + log2(gpus))")
    ap.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    ap.add_argument("--relabel", default="hub", choices=["hub", "degree", "none"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
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
