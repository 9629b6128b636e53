#!/usr/bin/env python
"""Measures the BASELINE.json configs that are not the bench.py headline, each with a parity check
against the oracle at a size the CPU finishes in seconds.  Writes gpurun_out/configs_<tag>.json.

    python scripts/run_configs.py [--tag r1] [--quick]

C1  PageRank(0.85, tol 1e-9) symmetric on Barabási–Albert n=100k m=5, 10 seed sets (fp64), vs oracle
C2  HeatKernel(t=3) and GenericGraphFilter K=40 on RMAT scale 22, fp64 and fp32
C3  256 seed sets through propagate() (panel SpMM kernel) on RMAT scale 24, fp32
C5  32 alphas of PageRank + AbsorbingWalks on Barabási–Albert n=10M m=8 (fp32)
(C4, multi-GPU, is measured by bench.py --gpus N.)
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel_l1(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).sum() / np.abs(b).sum())


def timed(fn, reps=1):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = None
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    import torch

    import pygrank_b200 as pgb
    from oracle import reference_port as orc
    from pygrank_b200 import device_synthetic, synthetic

    report = {"gpu": torch.cuda.get_device_name(0)}

    # ---------------------------------------------------------------- C1
    n, m = (20_000, 5) if args.quick else (100_000, 5)
    A = synthetic.ba_graph_host(n, m, seed=1)
    g = pgb.DeviceGraph.from_scipy(A, directed=False, normalization="symmetric")
    M = orc.to_sparse_matrix(A, "symmetric", False)
    seeds = synthetic.seed_sets(n, 10, 10, seed=0)
    alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
    alg(g, [int(v) for v in seeds[0]])
    gpu_s, cpu_s, calls, worst, iters_equal = 0.0, 0.0, 0, 0.0, True
    for s in seeds:
        p = np.zeros(n)
        p[s] = 1.0
        dt, r = timed(lambda: alg(g, [int(v) for v in s]).numpy())
        gpu_s += dt
        t0 = time.perf_counter()
        ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
        cpu_s += time.perf_counter() - t0
        calls += iters - 1
        worst = max(worst, rel_l1(r, ref))
        iters_equal &= (alg.convergence.iteration == iters)
    report["C1"] = {"graph": f"BA-like n={n} m={m} (nnz {A.nnz})", "solves": 10, "conv_calls": calls,
                    "gpu_e2e_s": gpu_s, "gpu_gteps": A.nnz * calls / gpu_s / 1e9,
                    "cpu_oracle_s": cpu_s, "cpu_gteps": A.nnz * calls / cpu_s / 1e9,
                    "iterations_equal": bool(iters_equal), "worst_rel_l1_fp64": worst}
    print("C1", report["C1"], flush=True)

    # ---------------------------------------------------------------- C2
    scale = 16 if args.quick else 22
    g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
    n = g.n
    seeds = synthetic.seed_sets(n, 3, 10, seed=0)
    w40 = [0.9 ** k for k in range(40)]
    c2 = {"graph": f"RMAT scale {scale} (n {n}, nnz {g.nnz})"}
    for dtype, name in ((torch.float64, "f64"), (torch.float32, "f32")):
        heat = pgb.HeatKernel(3, dtype=dtype)
        gen = pgb.GenericGraphFilter(w40, error_type="iters", max_iters=41, dtype=dtype)
        for label, alg in (("heat3", heat), ("generic40", gen)):
            alg(g, [int(v) for v in seeds[0]])
            dt, _ = timed(lambda: [alg(g, [int(v) for v in s]) for s in seeds])
            calls = (alg.convergence.iteration - 1) * len(seeds)
            c2[f"{label}_{name}"] = {"iterations": alg.convergence.iteration, "s_per_solve": dt / len(seeds),
                                     "gteps": g.nnz * calls / dt / 1e9}
    # the same filters issued by the UNMODIFIED reference through the b200 backend plugin (deferred vectors -> PolyRun)
    try:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from refutil import import_pygrank
        pg = import_pygrank()
        if pg is not None:
            from pygrank_b200 import backend as b200, lazy
            pgb.install(pg)
            b200.configure(dtype=torch.float32)
            with pg.Backend("b200"):
                pre = pgb.preprocessor(normalization="symmetric", assume_immutability=True)
                pers = []
                for s in seeds:
                    pv = torch.zeros(n, dtype=torch.float32, device="cuda")
                    pv[torch.from_numpy(s).cuda()] = 1.0
                    pers.append(pv)
                for label, mk in (("heat3", lambda: pg.HeatKernel(3, preprocessor=pre)),
                                  ("generic40", lambda: pg.GenericGraphFilter(w40, error_type="iters", max_iters=41, preprocessor=pre))):
                    alg = mk()
                    b200.to_tensor(alg(pg.to_signal(g, pers[0])).np)
                    lazy.reset_stats()
                    dt, _ = timed(lambda: [b200.to_tensor(alg(pg.to_signal(g, pv)).np) for pv in pers])
                    calls = (alg.convergence.iteration - 1) * len(seeds)
                    c2[f"{label}_f32_plugin"] = {"iterations": alg.convergence.iteration, "s_per_solve": dt / len(seeds),
                                                 "gteps": g.nnz * calls / dt / 1e9, "lazy_stats": dict(lazy.STATS)}
            pg.load_backend("numpy")
            b200.configure(dtype=torch.float64)
    except Exception as exc:   # reported, not fatal
        c2["plugin_error"] = repr(exc)[:300]
    # parity at a CPU-sized scale
    ps = 16
    gs = device_synthetic.rmat_graph_device(ps, 16, seed=1)
    Ms = orc.to_sparse_matrix(synthetic.rmat_graph_host(ps, 16, seed=1), "symmetric", False)
    p = np.zeros(1 << ps)
    p[synthetic.seed_sets(1 << ps, 1, 10, seed=0)[0]] = 1.0
    for label, mk, ref_fn in (("heat3", lambda dt: pgb.HeatKernel(3, dtype=dt), lambda: orc.heat_kernel(Ms, p, 3)),
                              ("generic40", lambda dt: pgb.GenericGraphFilter(w40, error_type="iters", max_iters=41, dtype=dt),
                               lambda: orc.generic_filter(Ms, p, w40, error_type="iters", max_iters=41))):
        ref, iters, _ = ref_fn()
        a64, a32 = mk(torch.float64), mk(torch.float32)
        r64, r32 = a64(gs, p).numpy(), a32(gs, p).numpy()
        c2[f"parity_{label}"] = {"scale": ps, "iterations_equal": a64.convergence.iteration == iters,
                                 "rel_l1_fp64": rel_l1(r64, ref), "rel_l1_fp32": rel_l1(r32, ref)}
    report["C2"] = c2
    print("C2", c2, flush=True)

    # ---------------------------------------------------------------- C3
    scale, B = (14, 16) if args.quick else (24, 256)
    g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
    n = g.n
    gen = torch.Generator(device="cuda").manual_seed(0)
    P = torch.zeros((n, B), dtype=torch.float32, device="cuda")
    idx = torch.randint(0, n, (10, B), device="cuda", generator=gen)
    P[idx, torch.arange(B, device="cuda")[None, :].expand(10, B)] = 1.0
    alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float32)
    alg.propagate(g, P[:, :8].contiguous())
    dt, out = timed(lambda: alg.propagate(g, P))            # default route: panels of 4 columns on the hub-blocked form
    its = list(alg.convergence.iterations)
    calls = sum(i - 1 for i in its)
    routes = {}
    for route in ("0", "csr"):                              # column by column (single-vector hsell) / item-stream panels
        os.environ["PGB_PANEL"] = route
        alg.propagate(g, P[:, :8].contiguous())
        dt_r, out_r = timed(lambda: alg.propagate(g, P))
        routes[route] = (dt_r, rel_l1(out_r[:, 1].cpu().numpy(), out[:, 1].cpu().numpy().astype(np.float64)),
                         list(alg.convergence.iterations))
        del out_r
    os.environ.pop("PGB_PANEL")
    # one panel, fixed step count: ms per panel iteration of the two kernels (gather + update) for 4 / 2 columns
    panel_ms = {}
    for dt_name, dtype_p, width in (("f32", torch.float32, 4), ("f64", torch.float64, 2)):
        fixed = pgb.PageRank(0.85, error_type="iters", max_iters=41, dtype=dtype_p)
        cols = P[:, :width].to(dtype_p).contiguous()
        fixed.propagate(g, cols)
        dtp = min(timed(lambda: fixed.propagate(g, cols))[0] for _ in range(3))
        panel_ms[dt_name] = {"columns": width, "steps": 40, "ms_per_panel_step": dtp / 40 * 1e3,
                             "edge_column_gteps": g.nnz * 40 * width / dtp / 1e9}
    one = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float32)
    one(g, P[:, 0].contiguous())
    dt1, r1 = timed(lambda: one(g, P[:, 0].contiguous()).np)
    c3 = {"graph": f"RMAT scale {scale} (n {n}, nnz {g.nnz})", "seed_sets": B, "propagate_s": dt,
          "edge_column_gteps": g.nnz * calls / dt / 1e9, "route": "hsell panels (pgb_affine_steps_panel)",
          "column_by_column_s": routes["0"][0], "column_by_column_edge_column_gteps": g.nnz * calls / routes["0"][0] / 1e9,
          "csr_panel_s": routes["csr"][0], "csr_panel_edge_column_gteps": g.nnz * calls / routes["csr"][0] / 1e9,
          "col1_rel_l1_vs_column_by_column": routes["0"][1], "col1_rel_l1_vs_csr_panel": routes["csr"][1],
          "iterations_off_by_more_than_one_vs_column_by_column":
              int(sum(abs(a - b) > 1 for a, b in zip(its, routes["0"][2]))),
          "panel_step": panel_ms,
          "column_iterations_min_max": [min(its), max(its)],
          "single_column_s": dt1, "single_column_gteps": g.nnz * (one.convergence.iteration - 1) / dt1 / 1e9,
          "col0_rel_l1_propagate_vs_single": rel_l1(out[:, 0].cpu().numpy(), r1.cpu().numpy().astype(np.float64))}
    del P, out
    ps = 14
    gs = device_synthetic.rmat_graph_device(ps, 16, seed=1)
    Ms = orc.to_sparse_matrix(synthetic.rmat_graph_host(ps, 16, seed=1), "symmetric", False)
    Pp = np.zeros((1 << ps, 12))
    for c, sset in enumerate(synthetic.seed_sets(1 << ps, 12, 10, seed=3)):
        Pp[sset, c] = 1.0
    a64 = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
    o64 = a64.propagate(gs, torch.from_numpy(Pp).cuda()).cpu().numpy()
    worst, same = 0.0, True
    for c in range(12):
        ref, iters, _ = orc.pagerank(Ms, Pp[:, c], 0.85, tol=1e-9, max_iters=1000)
        worst = max(worst, rel_l1(o64[:, c], ref))
        same &= a64.convergence.iterations[c] == iters
    c3["parity"] = {"scale": ps, "columns": 12, "iterations_equal": bool(same), "worst_rel_l1_fp64": worst}
    report["C3"] = c3
    print("C3", c3, flush=True)
    del g, gs

    # ---------------------------------------------------------------- C5
    n, m = (200_000, 8) if args.quick else (10_000_000, 8)
    t0 = time.perf_counter()
    g = device_synthetic.ba_graph_device(n, m, seed=1)
    torch.cuda.synchronize()
    build = time.perf_counter() - t0
    seeds = synthetic.seed_sets(n, 1, 10, seed=0)[0]
    alphas = np.linspace(0.5, 0.99, 32)
    algs = [pgb.PageRank(float(a), tol=1e-9, max_iters=2000, dtype=torch.float32) for a in alphas]
    algs[0](g, [int(v) for v in seeds])
    dt, _ = timed(lambda: [a(g, [int(v) for v in seeds]) for a in algs])
    calls = sum(a.convergence.iteration - 1 for a in algs)
    sw = pgb.PageRank(0.85, tol=1e-9, max_iters=2000, dtype=torch.float32)
    pvec = torch.zeros(n, dtype=torch.float32, device="cuda")
    pvec[torch.as_tensor(np.asarray(seeds, dtype=np.int64), device="cuda")] = 1.0
    sw.sweep(g, pvec, [float(a) for a in alphas[:4]])
    dts, swept = timed(lambda: sw.sweep(g, pvec, [float(a) for a in alphas]))
    sweep_its = list(sw.convergence.iterations)
    last = algs[-1](g, [int(v) for v in seeds]).np
    ab = pgb.AbsorbingWalks(0.85, tol=1e-9, max_iters=2000, dtype=torch.float32)
    dta, _ = timed(lambda: ab(g, [int(v) for v in seeds]))
    report["C5"] = {"graph": f"BA-like n={n} m={m} (nnz {g.nnz})", "build_s": build, "sweep_32_alphas_s": dt,
                    "sweep_conv_calls": calls, "sweep_gteps": g.nnz * calls / dt / 1e9,
                    "panel_sweep_32_alphas_s": dts, "panel_sweep_gteps": g.nnz * sum(i - 1 for i in sweep_its) / dts / 1e9,
                    "panel_sweep_iterations_off_by_more_than_one":
                        int(sum(abs(a.convergence.iteration - b) > 1 for a, b in zip(algs, sweep_its))),
                    "panel_sweep_last_alpha_rel_l1_vs_single": rel_l1(swept[:, -1].cpu().numpy(),
                                                                      last.cpu().numpy().astype(np.float64)),
                    "absorbing_iterations": ab.convergence.iteration, "absorbing_s": dta,
                    "absorbing_gteps": g.nnz * (ab.convergence.iteration - 1) / dta / 1e9}
    print("C5", report["C5"], flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"configs_{args.tag}.json"), "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
