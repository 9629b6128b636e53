"""torchrun entry: row-partitioned PPR on N GPUs vs the single-GPU engine and the oracle.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tests/dist_gpu_check.py [scale] [--no-shard]

Exchange variants are chosen by the caller through the environment (scripts/gpu_dist_parity.sh runs them all):
PGB_PEER=0 (NCCL all-gather), PGB_PEER_MULTICAST=1 (multimem.st), PGB_PEER_MASK=0/1 (reader masks off / forced on;
default: on from 4 ranks up).  Both symmetric buffers are NaN-filled before every solve (PGB_PEER_POISON), so a
value a peer failed to deliver cannot go unnoticed.  Required: fp64 iteration counts EQUAL to the single-GPU
engine's (and to the oracle's, scale <= 18), fp64 scores <= 1e-10 relative L1, fp32 <= 1e-5 with counts within one
step of the fp64 count (the fp32 error sequence crosses tol at a rounding-dependent step; printed).
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("PGB_PEER_POISON", "1")


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    scale = int(args[0]) if args else 16
    shard_check = "--no-shard" not in sys.argv
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import pygrank_b200 as pgb
    from pygrank_b200 import device_synthetic, synthetic
    from pygrank_b200.dist import DistGraph, DistPageRank
    n = 1 << scale
    g = DistGraph.rmat(scale, 16, seed=1)
    peer = g.peer_buffers(torch.float64)
    variant = "nccl all-gather" if peer is None else ("multicast" if peer["multicast"] else
                                                      "unicast stores" + (" + reader mask" if peer["mask"] is not None else ""))
    if rank == 0:
        print(f"world {world} scale {scale} exchange: {variant}" +
              ("" if peer is not None else " (" + getattr(g, "_peer_error", "PGB_PEER=0") + ")"))
    seeds = synthetic.seed_sets(n, 2, 10, seed=0)
    ok = True
    report = {"world": world, "scale": scale, "exchange": variant, "runs": []}
    single = device_synthetic.rmat_graph_device(scale, 16, seed=1) if rank == 0 else None
    if rank == 0:
        assert single.nnz == g.nnz_global, (single.nnz, g.nnz_global)
    for alpha in (0.85, 0.9):
        for dtype, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
            for s in seeds[: (2 if alpha == 0.85 else 1)]:
                alg = DistPageRank(alpha, tol=1e-9, max_iters=1000, dtype=dtype)
                local_scores = alg.rank(g, s)
                full = alg.gather_user_order(g, local_scores)
                if rank == 0:
                    ref_alg = pgb.PageRank(alpha, tol=1e-9, max_iters=1000, dtype=torch.float64)
                    ref = ref_alg(single, [int(v) for v in s]).np
                    err = float((full.double() - ref).abs().sum() / ref.abs().sum())
                    it_d, it_s = alg.iteration, ref_alg.convergence.iteration
                    good = bool(np.isfinite(err)) and err <= tol and \
                        (it_d == it_s if dtype == torch.float64 else abs(it_d - it_s) <= 1)
                    print(f"alpha={alpha} dtype={dtype} iters dist={it_d} single(fp64)={it_s} relL1={err:.3e} {'ok' if good else 'FAIL'}")
                    report["runs"].append({"alpha": alpha, "dtype": str(dtype), "iters": it_d, "iters_single": it_s,
                                           "rel_l1": err, "ok": good})
                    ok &= good
    if scale <= 18:
        alg = DistPageRank(0.85, tol=1e-9, max_iters=1000, dtype=torch.float64)
        full = alg.gather_user_order(g, alg.rank(g, seeds[0]))
        if rank == 0:
            from oracle import reference_port as orc
            M = orc.to_sparse_matrix(synthetic.rmat_graph_host(scale, 16, seed=1), "symmetric", False)
            p = np.zeros(n)
            p[seeds[0]] = 1.0
            ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
            err = float(np.abs(full.cpu().numpy() - ref).sum() / np.abs(ref).sum())
            good = err <= 1e-10 and alg.iteration == iters
            print(f"vs oracle: iters {alg.iteration} vs {iters}, relL1={err:.3e} {'ok' if good else 'FAIL'}")
            report["oracle"] = {"iters": alg.iteration, "iters_oracle": iters, "rel_l1": err, "ok": bool(good)}
            ok &= good
    if shard_check:
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
        report["ok"] = bool(ok)
        print("DIST REPORT " + json.dumps(report))
        print("DIST CHECK", "PASS" if ok else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
