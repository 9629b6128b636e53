#!/usr/bin/env python
"""Which golden cases does the fp32 mode finish at a different iteration count than the reference (fp64)?
Writes gpurun_out/fp32_iterations_r2.json (summarised in DESIGN.md §3)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from conftest import GOLDEN_GRAPHS, load_golden, rel_l1  # noqa: E402
from test_gpu_parity import RUN_NAMES, _runs  # noqa: E402

report = {"cases": 0, "equal": 0, "off_by_one": [], "worse": [], "worst_rel_l1": 0.0}
for name in GOLDEN_GRAPHS:
    z, A, directed = load_golden(name)
    for run in RUN_NAMES:
        norm, make, _ = _runs(pgb)[run]
        g = pgb.DeviceGraph.from_scipy(A, directed=directed, normalization=norm)
        for c in range(z["P"].shape[1]):
            alg = make({"dtype": torch.float32})
            try:
                got = alg(g, z["P"][:, c].copy()).numpy()
            except Exception as exc:
                report["worse"].append([name, run, c, repr(exc)[:80]])
                continue
            want = int(z[f"run_{run}_iters"][c])
            report["cases"] += 1
            d = alg.convergence.iteration - want
            err = rel_l1(got, z[f"run_{run}_scores"][:, c])
            report["worst_rel_l1"] = max(report["worst_rel_l1"], err)
            if d == 0:
                report["equal"] += 1
            elif abs(d) == 1:
                e = alg.convergence.errors.cpu().numpy()
                report["off_by_one"].append({"graph": name, "run": run, "column": c, "fp32": alg.convergence.iteration,
                                             "reference": want, "last_errors": [float(x) for x in e[-2:]],
                                             "tol": float(alg.convergence.tol) if alg.convergence.tol else None})
            else:
                report["worse"].append([name, run, c, alg.convergence.iteration, want])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(report, open(os.path.join(ROOT, "gpurun_out", "fp32_iterations_r2.json"), "w"), indent=1)
print(json.dumps({k: (v if not isinstance(v, list) else len(v)) for k, v in report.items()}))
for o in report["off_by_one"][:12]:
    print(o)
print(report["worse"][:5])
