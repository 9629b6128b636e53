"""torchrun entry: the golden graphs (outputs of the real reference, tests/golden/) through the ROW-PARTITIONED path —
DistGraph.from_scipy + the distributed filters — on N GPUs: iteration counts equal, fp64 scores <= 1e-10.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 \
        tests/dist_golden_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("PGB_PEER_POISON", "1")


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from conftest import load_golden, rel_l1
    from pygrank_b200 import dist as D
    f64 = torch.float64
    runs = {
        "ppr85": lambda: D.DistPageRank(0.85, tol=1e-9, max_iters=1000, dtype=f64),
        "ppr90_noq": lambda: D.DistPageRank(0.9, tol=1e-9, max_iters=1000, use_quotient=False, dtype=f64),
        "ppr85_l1": lambda: D.DistPageRank(0.85, tol=1e-7, error_type="L1", max_iters=1000, dtype=f64),
        "heat3_tol9": lambda: D.DistHeatKernel(3, tol=1e-9, dtype=f64),
        "gen40": lambda: D.DistGenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters", max_iters=41, dtype=f64),
        "pprclosed": lambda: D.DistPageRankClosed(0.85, tol=1e-9, max_iters=1000, dtype=f64),
        "absorb85": lambda: D.DistAbsorbingWalks(0.85, tol=1e-9, max_iters=1000, dtype=f64),
    }
    ok = True
    for name in ("ba2000", "rmat10", "gnp600d"):
        z, A, directed = load_golden(name)
        g = D.DistGraph.from_scipy(A, directed=directed, normalization="auto")
        for rname, make in runs.items():
            for c in (0, 3):
                alg = make()
                full = alg.gather_user_order(g, alg.rank(g, personalization=z["P"][:, c]))
                want_it = int(z[f"run_{rname}_iters"][c])
                err = rel_l1(full.cpu().numpy(), z[f"run_{rname}_scores"][:, c])
                good = alg.iteration == want_it and err <= 1e-10
                ok &= bool(good)
                if rank == 0 and not good:
                    print(f"FAIL {name} {rname} col {c}: iters {alg.iteration} vs {want_it}, relL1 {err:.3e}")
        if rank == 0:
            print(f"{name}: world {dist.get_world_size()} directed={directed} normalization={g.normalization} "
                  f"hsell={'yes' if g.hsell(f64) is not None else 'no'} blocks={getattr(g.hsell(f64), 'n_blocks', 0)}")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("DIST GOLDEN", "PASS" if int(flag.item()) == 1 else "FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
