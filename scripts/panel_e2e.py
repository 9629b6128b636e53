"""propagate() of B seed sets on RMAT-<scale> through the hub-blocked panel path: wall time, edge-column rate and (with
PGB_PANEL_TIMING=1) the phases of the driver."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from pygrank_b200 import device_synthetic  # noqa: E402

scale = int(os.environ.get("PROBE_SCALE", "24"))
B = int(os.environ.get("PROBE_COLUMNS", "256"))
g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
n = g.n
gen = torch.Generator(device="cuda").manual_seed(0)
P = torch.zeros((n, B), dtype=torch.float32, device="cuda")
idx = torch.randint(0, n, (10, B), device="cuda", generator=gen)
P[idx, torch.arange(B, device="cuda")[None, :].expand(10, B)] = 1.0
alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float32)
for attr in ("panel_chunk", "panel_group"):
    if os.environ.get("PROBE_" + attr.upper()):
        setattr(alg, attr, int(os.environ["PROBE_" + attr.upper()]))
alg.propagate(g, P[:, :8])
torch.cuda.synchronize()
for rep in range(2):
    t0 = time.perf_counter()
    out = alg.propagate(g, P)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    calls = sum(i - 1 for i in alg.convergence.iterations)
    print(json.dumps({"rep": rep, "seconds": dt, "column_steps": calls, "edge_column_gteps": g.nnz * calls / dt / 1e9,
                      "chunk": getattr(alg, "panel_chunk", None), "group": getattr(alg, "panel_group", None)}), flush=True)
    del out
