#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "closed_form_propagate or batched or sweep or panel_max" 2>&1 | tail -12
python - <<'PY'
import time, torch, pygrank_b200 as pgb
from pygrank_b200 import device_synthetic
g = device_synthetic.rmat_graph_device(22, 16, seed=1)
n = g.n
gen = torch.Generator(device="cuda").manual_seed(0)
B = 32
P = torch.zeros((n, B), dtype=torch.float32, device="cuda")
idx = torch.randint(0, n, (10, B), device="cuda", generator=gen)
P[idx, torch.arange(B, device="cuda")[None, :].expand(10, B)] = 1.0
import os
for name, mk in (("heat3", lambda: pgb.HeatKernel(3, tol=1e-9, dtype=torch.float32)), ("gen40", lambda: pgb.GenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters", max_iters=41, dtype=torch.float32))):
    res = {}
    for route in ("1", "0"):
        os.environ["PGB_PANEL"] = route
        alg = mk()
        alg.propagate(g, P[:, :8]); torch.cuda.synchronize()
        t0 = time.perf_counter(); out = alg.propagate(g, P); torch.cuda.synchronize()
        res[route] = (time.perf_counter() - t0, sum(i - 1 for i in alg.convergence.iterations), out)
    err = float((res["1"][2] - res["0"][2]).abs().sum() / res["0"][2].abs().sum())
    print(name, "RMAT-22 32 columns: panels %.4f s (%.0f G edge-col/s), column by column %.4f s (%.0f), rel L1 %.2e" % (
        res["1"][0], g.nnz * res["1"][1] / res["1"][0] / 1e9, res["0"][0], g.nnz * res["0"][1] / res["0"][0] / 1e9, err))
PY
