"""One timing of the hub-blocked panel step on RMAT-<scale> (4 x fp32 or 2 x fp64 columns, fixed 40 steps) together with
the shape of the panel form; knobs come from the PGB_HSELL_* environment (one process per variant)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from pygrank_b200 import device_synthetic  # noqa: E402

scale = int(os.environ.get("PROBE_SCALE", "24"))
dtype = torch.float64 if os.environ.get("PROBE_DTYPE") == "f64" else torch.float32
width = 4 if dtype == torch.float32 else 2
g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
n = g.n
form = g.in_view.hsell_panel()
torch.cuda.synchronize()
stats = {"H": form.block_cols, "K": form.n_blocks, "hub_chunks": form.n_hub_chunks, "tail_chunks": form.n_tail_chunks,
         "pieces": form.n_pieces, "hub_slots": form.n_hub_words * 2, "tail_slots": form.n_tail_words,
         "tail_entries": int((form.tail_cols >= 0).sum()), "nnz": g.nnz, "form_MB": form.nbytes() / 1e6}
stats["hub_entries"] = g.nnz - stats["tail_entries"]
gen = torch.Generator(device="cuda").manual_seed(0)
P = torch.zeros((n, width), dtype=dtype, device="cuda")
idx = torch.randint(0, n, (10, width), device="cuda", generator=gen)
P[idx, torch.arange(width, device="cuda")[None, :].expand(10, width)] = 1.0
steps = 40
alg = pgb.PageRank(0.85, error_type="iters", max_iters=steps + 1, dtype=dtype)
alg.propagate(g, P)
torch.cuda.synchronize()
best = 1e9
for _ in range(int(os.environ.get("PROBE_REPS", "3"))):
    t0 = time.perf_counter()
    alg.propagate(g, P)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
stats["ms_per_panel_step"] = best / steps * 1e3
stats["edge_column_gteps"] = g.nnz * steps * width / best / 1e9
stats["env"] = {k: v for k, v in os.environ.items() if k.startswith("PGB_")}
print(json.dumps(stats), flush=True)
