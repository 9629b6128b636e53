"""bench.py's N>1 leg: weak-scaling PageRank on a row-partitioned RMAT graph (scale 24 + log2 N, BASELINE config 4).

Before anything is timed the run checks itself (`parity` in the JSON line, non-zero exit on failure):
* the SAME code path (same world size, same exchange, reader masks as selected, NaN-poisoned symmetric buffers) on a
  small RMAT graph against the single-GPU engine on rank 0: fp64 iteration counts equal, fp64 <= 1e-10 / fp32 <= 1e-5
  relative L1;
* on the timed graph itself a size-independent property, evaluated WITHOUT the engine's kernels (plain torch gathers
  over this rank's CSR rows): the iterates after m and m+1 steps obey r_{m+1} = (alpha*M r_m + (1-alpha)*p)/s with one
  s on every row; the highest-scored rows of every rank must agree on it to rounding.  (The converged vector itself is
  no fixed point to that accuracy: Mabs divides by n, so at 10^8 nodes the loop stops while the top rows still move.)
"""
from __future__ import annotations

import ctypes
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def run(args):
    from . import _capi as C
    from . import synthetic
    from .dist import DistGraph, DistPageRank
    from .graph import dtype_code, span_struct
    import bench as B

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=dev)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    w = 4 if dtype == torch.float32 else 8
    scale = args.scale
    t0 = time.perf_counter()
    g = DistGraph.rmat(scale, 16, seed=1)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    alpha = B.alpha_for(world)
    parity = None
    if not args.no_parity:
        parity = small_scale_parity(alpha, rank, world, dev)
    alg = DistPageRank(alpha, tol=B.TOL, max_iters=B.MAX_ITERS, dtype=dtype)
    total = args.warmup + args.steps
    seeds = synthetic.seed_sets(g.n_nodes, total, 10, seed=0)
    pers = [alg.local_personalization(g, s) for s in seeds]

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.perf_counter()
        e0.record()
        calls = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = max(e0.elapsed_time(e1), (time.perf_counter() - t) * 1e3 * 0.0)
        worst = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(worst, op=dist.ReduceOp.MAX)
        return float(worst.item()), calls

    for i in range(args.warmup):
        alg.rank(g, p_local=pers[i][0], norm=pers[i][1])
    if parity is not None:
        parity["one_step_on_timed_graph"] = one_step_check(g, alpha, dtype, seeds[0])
        parity["ok"] = bool(parity["ok"] and parity["one_step_on_timed_graph"]["ok"])
    sampler = B.ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = C.LAUNCHES[0]

    def resident():
        calls = 0
        for i in range(args.warmup, total):
            alg.rank(g, p_local=pers[i][0], norm=pers[i][1])
            calls += alg.iteration - 1
        return calls

    dev_ms, conv_calls = timed(resident)
    if os.environ.get("PGB_DIST_TIMING", "0") == "1" and rank == 0:
        print("DIST TIMING", json.dumps(getattr(alg, "timing", None)), flush=True)
    launches = C.LAUNCHES[0] - launches0
    value = g.nnz_global * conv_calls / (dev_ms * 1e-3) / 1e9

    # scores of every solve to pinned host memory inside the timed region; the copy of solve k overlaps solve k+1
    host_out = [torch.empty(g.n_local, dtype=dtype).pin_memory(), torch.empty(g.n_local, dtype=dtype).pin_memory()]
    copy_stream = torch.cuda.Stream()

    def end_to_end():
        calls = 0
        keep = []
        for i in range(args.warmup, total):
            r = alg.rank(g, seeds[i])
            ready = torch.cuda.Event()
            ready.record()
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                host_out[i & 1].copy_(r, non_blocking=True)
            r.record_stream(copy_stream)
            keep.append(r)
            del keep[:-2]
            calls += alg.iteration - 1
        copy_stream.synchronize()
        return calls

    e2e_ms, e2e_calls = timed(end_to_end)
    e2e_value = g.nnz_global * e2e_calls / (e2e_ms * 1e-3) / 1e9
    clocks = sampler.stop() if rank == 0 else None

    # the fused kernel alone on this rank's rows (no exchange), CUDA events on the launch stream
    lib = C.lib()
    code = dtype_code(dtype)
    st = C.stream_ptr()
    n_loc, off = g.n_local, g.offset
    sf = [0.0] * C.STATE_LEN
    si = [0] * C.STATE_LEN
    sf[C.SF_ALPHA], sf[C.SF_INVS], sf[C.SF_MEAN], sf[C.SF_NORM] = alpha, 1.0, float(g.n_nodes), 10.0
    si[C.SI_MAX_ITERS], si[C.SI_END_MODULO], si[C.SI_ERR_MODE], si[C.SI_QUOTIENT] = 10 ** 6, 1, C.ERR_ITERS, 0
    state_f64 = torch.tensor(sf, dtype=torch.float64, device=dev)
    state_i32 = torch.tensor(si, dtype=torch.int32, device=dev)
    zfull = [torch.rand(g.n_global, dtype=dtype, device=dev), torch.rand(g.n_global, dtype=dtype, device=dev)]
    q = torch.zeros(n_loc, dtype=dtype, device=dev)
    cvec = g.vec("c", dtype)
    form = g.hsell(dtype)
    cs = g.view.cstruct(dtype, hsell=form is not None)
    ws = g.view.new_span_ws(dtype if form is not None else None)
    err_hist = torch.zeros(128, dtype=torch.float64, device=dev)
    reps = 20

    def steps(first, count):
        C.check(lib.pgb_affine_steps(ctypes.byref(cs), code, alpha, None, None, C.ptr(cvec), C.ptr(q),
                                     C.ptr(zfull[0]), C.ptr(zfull[1]), off, C.ptr(state_f64), C.ptr(state_i32),
                                     C.ptr(err_hist), span_struct(ws), first, count, 1, st))

    steps(1, 3)
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    steps(4, reps)
    k1.record()
    torch.cuda.synchronize()
    kernel_ms = torch.tensor([k0.elapsed_time(k1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(kernel_ms, op=dist.ReduceOp.MAX)
    kernel_ms = float(kernel_ms.item())
    alg_bytes = g.nnz_local * 4 + (n_loc + 1) * 4 + (g.n_global + 4 * n_loc) * w
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = B.hbm_peak()
    nnz_all = [None] * world
    dist.all_gather_object(nnz_all, g.nnz_local)
    peer = g.peer_buffers(dtype)
    if peer is None:
        exchange = "NCCL all_gather_into_tensor of n/N rank-vector slices + 16-byte all_reduce per iteration"
    else:
        exchange = ("fused into the update kernel: z' and the convergence sums written into every rank's symmetric "
                    "memory (%s) + one barrier kernel per iteration; no NCCL in the loop"
                    % ("NVSwitch multicast, multimem.st" if peer["multicast"] else
                       "one NVLink store per peer that reads the row" +
                       ("" if peer.get("sent_fraction") is None else ": %.0f%% of (row, rank) pairs" % (100 * peer["sent_fraction"]))))
    if rank == 0:
        line = {
            "metric": "PPR GTEPS", "value": value, "unit": "GTEPS", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": dict(B.workload_config(world, scale, alpha), n=g.n_nodes, nnz=g.nnz_global,
                           nnz_per_rank=[int(x) for x in nnz_all], conv_calls_per_solve=conv_calls / args.steps,
                           exchange=exchange,
                           graph_build_s=round(build_s, 2)),
            "e2e": {"value": e2e_value, "unit": "GTEPS", "h2d_bytes_per_step": 10 * 8 + 10 * 8,
                    "d2h_bytes_per_step": g.n_local * w * world + 64 * 2},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None,
                         "kernel": ("hsell gather+update<%s> (one step, per rank, max over ranks)" if form is not None
                                    else "item_stream_kernel<%s,unweighted,AFFINE,SYMDEG> (per rank, max over ranks)") % args.dtype,
                         "kernel_ms": kernel_ms, "algorithmic_bytes": alg_bytes, "peak_source": peak_src},
            "cpu_baseline": None,
            "parity": parity,
            "clocks": clocks,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        raise SystemExit(3)


PARITY_SCALE = 20


def small_scale_parity(alpha, rank, world, dev):
    """The N-rank path (this world size, this exchange configuration, NaN-poisoned buffers) against the single-GPU
    engine on RMAT-20; rank 0 compares, every rank learns the verdict."""
    import pygrank_b200 as pgb
    from . import device_synthetic, synthetic
    from .dist import DistGraph, DistPageRank
    g = DistGraph.rmat(PARITY_SCALE, 16, seed=1)
    peer = g.peer_buffers(torch.float32)
    exchange = "nccl all-gather" if peer is None else ("multicast" if peer["multicast"] else
                                                       "unicast stores" + (" + reader mask" if peer["mask"] is not None else ""))
    seeds = synthetic.seed_sets(1 << PARITY_SCALE, 2, 10, seed=0)
    single = device_synthetic.rmat_graph_device(PARITY_SCALE, 16, seed=1) if rank == 0 else None
    out = {"ranks": world, "rmat_scale": PARITY_SCALE, "against": "single-GPU engine (fp64), same graph and seeds",
           "exchange": exchange, "poisoned_buffers": True, "ok": True}
    for name, dtype, tol in (("f64", torch.float64, 1e-10), ("f32", torch.float32, 1e-5)):
        worst, its_d, its_s = 0.0, [], []
        for s in seeds:
            alg = DistPageRank(alpha, tol=1e-9, max_iters=1000, dtype=dtype)
            alg.poison = True
            full = alg.gather_user_order(g, alg.rank(g, s))
            if rank == 0:
                ref_alg = pgb.PageRank(alpha, tol=1e-9, max_iters=1000, dtype=torch.float64)
                ref = ref_alg(single, [int(v) for v in s]).np
                worst = max(worst, float((full.double() - ref).abs().sum() / ref.abs().sum()))
                its_d.append(int(alg.iteration))
                its_s.append(int(ref_alg.convergence.iteration))
        if rank == 0:
            same = its_d == its_s
            ok = bool(np.isfinite(worst)) and worst <= tol and (same if name == "f64" else
                                                                 all(abs(a - b) <= 1 for a, b in zip(its_d, its_s)))
            out[name] = {"rel_l1": worst, "tolerance": tol, "iterations": its_d, "iterations_single_gpu": its_s,
                         "iterations_equal": same, "ok": ok}
            out["ok"] = out["ok"] and ok
    flag = torch.tensor([1 if out["ok"] else 0], device=dev)
    dist.broadcast(flag, 0)
    out["ok"] = bool(int(flag.item()))
    del g, single
    torch.cuda.empty_cache()
    return out


def one_step_check(g, alpha, dtype, seeds, m: int = 6, samples: int = 512):
    """Size-independent check on the TIMED graph (any convergence state): the engine's iterates after m and m+1
    steps (two fixed-length solves, error_type="iters") must satisfy r_{m+1} = (alpha*M r_m + (1-alpha)*p)/s with ONE
    s for all rows.  The right-hand side is evaluated on the highest-scored rows of every rank with plain torch index
    arithmetic over the rank's CSR rows (M_ij = 1/sqrt(d_i d_j)) — no engine kernel — and the ratios v_i / r_{m+1,i}
    of all ranks must agree to rounding."""
    from .dist import DistPageRank
    dev = g.view.indptr.device
    f64 = torch.float64
    iterates = []
    for k in (m, m + 1):
        alg = DistPageRank(alpha, tol=1e-9, max_iters=k + 1, dtype=dtype, error_type="iters")
        p_local, norm = alg.local_personalization(g, seeds)
        iterates.append(alg.rank(g, p_local=p_local, norm=norm))
    r_m, r_next = iterates
    r_full = torch.empty(g.n_global, dtype=dtype, device=dev)
    dist.all_gather_into_tensor(r_full, r_m.contiguous(), group=g.group)
    deg_local = (g.view.indptr[1:] - g.view.indptr[:-1]).to(torch.int32)
    deg_full = torch.empty(g.n_global, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(deg_full, deg_local.contiguous(), group=g.group)
    rows = torch.topk(r_next, min(samples, g.n_local)).indices
    rows = rows[deg_local[rows] > 0][:256]
    ratios = []
    ip = g.view.indptr
    for i in rows.tolist():
        cols = g.view.indices[int(ip[i]):int(ip[i + 1])].long()
        y = (r_full[cols].to(f64) / torch.sqrt(deg_full[cols].to(f64))).sum() / np.sqrt(float(deg_local[i]))
        v = alpha * y + (1 - alpha) * p_local[i].to(f64)
        ratios.append(v / r_next[i].to(f64))
    if ratios:
        rt = torch.stack(ratios)
        lo, hi, cnt = rt.min(), rt.max(), torch.tensor(float(len(ratios)), device=dev, dtype=f64)
    else:
        lo, hi = torch.tensor(float("inf"), device=dev, dtype=f64), torch.tensor(float("-inf"), device=dev, dtype=f64)
        cnt = torch.tensor(0.0, device=dev, dtype=f64)
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=g.group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=g.group)
    dist.all_reduce(cnt, group=g.group)
    spread = float((hi - lo) / ((hi + lo) / 2)) if float(cnt) > 0 else float("nan")
    tol = 1e-4 if dtype == torch.float32 else 1e-10
    ok = float(cnt) > 0 and np.isfinite(spread) and spread <= tol
    return {"property": "r_{m+1,i} * s = alpha*(M r_m)_i + (1-alpha)*p_i with one s on all rows of all ranks "
                        "(right-hand side by torch gathers over the CSR rows, no engine kernel)",
            "steps_m": m, "rows_checked": int(cnt), "ratio_min": float(lo), "ratio_max": float(hi),
            "relative_spread": spread, "tolerance": tol, "ok": bool(ok)}
