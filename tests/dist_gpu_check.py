"""torchrun entry: row-partitioned PPR on N GPUs vs the single-GPU engine and the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_gpu_check.py [scale]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pygrank_b200 as pgb
    from pygrank_b200 import device_synthetic, synthetic
    from pygrank_b200.dist import DistGraph, DistPageRank
    n = 1 << scale
    g = DistGraph.rmat(scale, 16, seed=1)
    peer = g.peer_buffers(torch.float64)
    if rank == 0:
        print("peer exchange:", "off (" + getattr(g, "_peer_error", "disabled") + ")" if peer is None else
              ("multicast" if peer["multicast"] else "unicast stores"))
    seeds = synthetic.seed_sets(n, 2, 10, seed=0)
    ok = True
    single = device_synthetic.rmat_graph_device(scale, 16, seed=1) if rank == 0 else None
    if rank == 0:
        assert single.nnz == g.nnz_global, (single.nnz, g.nnz_global)
    for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
        for s in seeds:
            alg = DistPageRank(0.85, tol=1e-9, max_iters=1000, dtype=dtype)
            local_scores = alg.rank(g, s)
            full = alg.gather_user_order(g, local_scores)
            if rank == 0:
                ref_alg = pgb.PageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float64)
                ref = ref_alg(single, [int(v) for v in s]).np
                err = float((full.double() - ref).abs().sum() / ref.abs().sum())
                same_iters = alg.iteration == ref_alg.convergence.iteration
                print(f"dtype={dtype} iters dist={alg.iteration} single={ref_alg.convergence.iteration} relL1={err:.3e}")
                ok &= err <= tol and (same_iters or dtype == torch.float32)
    if rank == 0 and scale <= 18:
        from oracle import reference_port as orc
        M = orc.to_sparse_matrix(synthetic.rmat_graph_host(scale, 16, seed=1), "symmetric", False)
        p = np.zeros(n)
        p[seeds[0]] = 1.0
        ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
    alg = DistPageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float64)
    full = alg.gather_user_order(g, alg.rank(g, seeds[0]))
    if rank == 0 and scale <= 18:
        err = float(np.abs(full.cpu().numpy() - ref).sum() / np.abs(ref).sum())
        print(f"vs oracle: iters {alg.iteration} vs {iters}, relL1={err:.3e}")
        ok &= err <= 1e-10 and alg.iteration == iters
    # independent-unit sharding: columns of propagate() dealt to the ranks, graph replicated, no data-path collective
    from pygrank_b200.dist import propagate_sharded
    rep = device_synthetic.rmat_graph_device(min(scale, 16), 16, seed=1)
    nr = rep.n
    gen = torch.Generator(device="cuda").manual_seed(7)
    feats = torch.zeros((nr, 5), dtype=torch.float64, device="cuda")
    idx = torch.randint(0, nr, (10, 5), device="cuda", generator=gen)
    feats[idx, torch.arange(5, device="cuda")[None, :].expand(10, 5)] = 1.0
    algp = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
    sharded, its = propagate_sharded(algp, rep, feats)
    if rank == 0:
        alg1 = pgb.PageRank(0.85, tol=1e-9, max_iters=1000)
        whole = alg1.propagate(rep, feats)
        errp = float((sharded - whole).abs().sum() / whole.abs().sum())
        print(f"propagate_sharded: relL1={errp:.3e} iterations {its} vs {list(alg1.convergence.iterations)}")
        ok &= errp <= 1e-13 and its == list(alg1.convergence.iterations)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
