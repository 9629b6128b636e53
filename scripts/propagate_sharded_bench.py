"""propagate() of B seed sets over the GPUs of one box (config 3 at N GPUs): every rank holds a replica of the RMAT graph
and runs its share of the columns through the hub-blocked panel path (propagate_sharded: no data-path collective).
torchrun --nproc-per-node N scripts/propagate_sharded_bench.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import pygrank_b200 as pgb  # noqa: E402
from pygrank_b200 import device_synthetic  # noqa: E402
from pygrank_b200.dist import column_shard, propagate_sharded  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dist.init_process_group("nccl")
scale = int(os.environ.get("PROBE_SCALE", "24"))
B = int(os.environ.get("PROBE_COLUMNS", "256"))
g = device_synthetic.rmat_graph_device(scale, 16, seed=1)
n = g.n
gen = torch.Generator(device="cuda").manual_seed(0)
P = torch.zeros((n, B), dtype=torch.float32, device="cuda")
idx = torch.randint(0, n, (10, B), device="cuda", generator=gen)
P[idx, torch.arange(B, device="cuda")[None, :].expand(10, B)] = 1.0
alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float32)
alg.propagate(g, P[:, :8])
torch.cuda.synchronize()
best = None
for rep in range(2):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    local, its = propagate_sharded(alg, g, P, gather=False)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    steps = torch.tensor([sum(i - 1 for i in its)], device="cuda", dtype=torch.float64)
    dist.all_reduce(steps)
    if best is None or float(dt) < best[0]:
        best = (float(dt), float(steps))
# spot check: this rank's first column against a single solve
mine = column_shard(B, rank, world)
one = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float32)
ref = one(g, P[:, mine.start].contiguous()).np
err = float((local[:, 0] - ref).abs().sum(dtype=torch.float64) / ref.abs().sum(dtype=torch.float64))
errs = torch.tensor([err], device="cuda", dtype=torch.float64)
dist.all_reduce(errs, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"gpus": world, "graph": f"RMAT scale {scale} replicated", "seed_sets": B, "seconds": best[0],
                      "column_steps": best[1], "edge_column_gteps": g.nnz * best[1] / best[0] / 1e9,
                      "worst_first_column_rel_l1_vs_single": float(errs)}), flush=True)
dist.destroy_process_group()
