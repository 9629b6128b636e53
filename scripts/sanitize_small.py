"""Tiny end-to-end pass over every per-iteration kernel, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py
(hsell with small blocks so that many blocks, a tail, heavy slices and the second-level reduction appear)."""
import os
import sys

os.environ.setdefault("PGB_HSELL_BLOCK_COLS", "256")
os.environ.setdefault("PGB_HSELL_BLOCKS", "6")
os.environ.setdefault("PGB_HSELL_MIN_ENTRIES", "8")
os.environ.setdefault("PGB_HSELL_HEAVY_PARTS", "4")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from pygrank_b200 import _capi as C  # noqa: E402
from pygrank_b200 import device_synthetic  # noqa: E402

scale = 12
n = 1 << scale
g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
p = torch.zeros(n, dtype=torch.float64, device="cuda")
p[torch.arange(0, n, n // 10, device="cuda")] = 1.0
for dtype in (torch.float32, torch.float64):
    a = pgb.PageRank(0.85, tol=1e-9, max_iters=200, dtype=dtype)
    r = a(g, p.to(dtype))
    h = pgb.HeatKernel(3, dtype=dtype)(g, p.to(dtype))
    w = pgb.AbsorbingWalks(0.85, tol=1e-9, max_iters=200, dtype=dtype)(g, p.to(dtype))
    m = pgb.PageRank(0.85, tol=1e-8, max_iters=200, dtype=dtype, error_type="max")(g, p.to(dtype))
    y = g.conv(p.to(dtype))
    print(dtype, a.convergence.iteration, float(r.np.sum()), float(h.np.sum()), float(w.np.sum()), float(y.sum()))
# round 2: hub-signature order + tail windows, in-kernel dropout, the deterministic partial-row path, the plugin route
d = pgb.GenericGraphFilter([0.5, 0.3, 0.2], error_type="iters", max_iters=4)(g, p, graph_dropout=0.3)
print("dropout", float(d.np.sum()), float(g.dropout(0.5).conv(p).sum()))
os.environ["PGB_HSELL_TAIL_WINDOWS"] = "3"
os.environ["PGB_HSELL_TAIL_WINDOW_MIN"] = "4"
g3 = device_synthetic.rmat_graph_device(scale, 16, seed=2)
print("windows", float(pgb.PageRank(0.85, tol=1e-9, max_iters=200, dtype=torch.float32)(g3, p.float()).np.sum()))
os.environ["PGB_DETERMINISTIC"] = "1"
g4 = device_synthetic.rmat_graph_device(scale, 16, seed=2)
print("deterministic", float(pgb.PageRank(0.85, tol=1e-9, max_iters=200)(g4, p).np.sum()))
del os.environ["PGB_DETERMINISTIC"], os.environ["PGB_HSELL_TAIL_WINDOWS"]
from pygrank_b200 import backend as B  # noqa: E402
x = B.to_array(p)
num = B.conv(x, g) * 0.85 + x * 0.15
nxt = num / B.sum(num)
print("lazy", float(B.sum(B.abs(x - nxt)) / n), float(B.sum(nxt)))
for family in ("1", "csr"):          # hub-blocked panels (pgb_affine_steps_panel), item-stream panels
    os.environ["PGB_PANEL"] = family
    for dtype in (torch.float32, torch.float64):
        feats = torch.stack([p, p * 2, p * 0, p + 1, p * 3], 1).to(dtype)
        out = pgb.PageRank(0.85, tol=1e-9, max_iters=200, dtype=dtype).propagate(g, feats)
        print("panel", family, dtype, tuple(out.shape), float(out.sum()))
    if family == "1":                # polynomial mode of the panel kernel (closed-form filters)
        feats = torch.stack([p, p * 2, p * 0, p + 1, p * 3], 1)
        print("poly panel", float(pgb.HeatKernel(3, tol=1e-9).propagate(g, feats).sum()),
              float(pgb.GenericGraphFilter([0.5, 0.3, 0.2], tol=1e-9, dtype=torch.float32).propagate(g, feats.float()).sum()))
del os.environ["PGB_PANEL"]
print("sweep", float(pgb.PageRank(0.85, tol=1e-9, max_iters=200, dtype=torch.float32).sweep(g, p.float(), [0.5, 0.7, 0.9]).sum()))
# weighted graph on the hub-blocked form (edge values in the form)
src, dst = device_synthetic.rmat_edges_device(scale, 16, seed=5)
gw = pgb.DeviceGraph.from_edges(n, src, dst, weights=torch.rand(src.numel(), dtype=torch.float64, device="cuda") + 0.5,
                                normalization="symmetric")
for dtype in (torch.float32, torch.float64):
    assert gw.in_view.hsell(dtype) is not None and gw.in_view.hsell(dtype).hub_vals is not None
    print("weighted hsell", dtype, float(pgb.PageRank(0.85, tol=1e-9, max_iters=200, dtype=dtype)(gw, p.to(dtype)).np.sum()),
          float(gw.conv(p.to(dtype)).sum()))
C.check(C.lib().pgb_set_kernel_variant(3))
a = pgb.PageRank(0.85, tol=1e-9, max_iters=200)
print("item stream", a(g, p).np.sum().item(), a.convergence.iteration)
torch.cuda.synchronize()
print("sanitize_small done")
