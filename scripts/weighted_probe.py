"""Weighted RMAT-<scale>: ms per fused PPR step on the hub-blocked form with edge values against the item-stream kernel."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from pygrank_b200 import _capi as C  # noqa: E402
from pygrank_b200 import device_synthetic  # noqa: E402

scale = int(os.environ.get("PROBE_SCALE", "24"))
n = 1 << scale
src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=1)
wts = torch.rand(src.numel(), dtype=torch.float64, device="cuda") * 4 + 0.25
g = pgb.DeviceGraph.from_edges(n, src, dst, weights=wts, directed=False, drop_self_loops=True, normalization="symmetric")
del src, dst, wts
p = torch.zeros(n, dtype=torch.float32, device="cuda")
p[torch.randint(0, n, (10,), device="cuda")] = 1.0
steps = 30
out = {"nnz": g.nnz, "weighted": bool(g.in_view.weighted)}
for dtype in (torch.float32, torch.float64):
    for variant, name in ((4, "hsell"), (3, "item_stream")):
        C.check(C.lib().pgb_set_kernel_variant(variant))
        alg = pgb.PageRank(0.85, error_type="iters", max_iters=steps + 1, dtype=dtype)
        alg(g, p.to(dtype))
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            r = alg(g, p.to(dtype))
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        out[f"{name}_{'f32' if dtype == torch.float32 else 'f64'}_ms_per_step"] = best / steps * 1e3
    form = g.in_view.hsell(dtype)
    out[f"form_MB_{'f32' if dtype == torch.float32 else 'f64'}"] = form.nbytes() / 1e6
C.check(C.lib().pgb_set_kernel_variant(4))
print(json.dumps(out), flush=True)
