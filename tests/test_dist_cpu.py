"""N>1 host logic on CPU: world_size-2 gloo processes run the partition helpers of
pygrank_b200/dist.py and a stand-in for the per-rank kernel (the oracle's arithmetic, test-only),
exchanging slices with all_gather / all_reduce exactly as the GPU driver does."""
import os
import sys

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, scale, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygrank_b200 import synthetic
    from pygrank_b200.dist import interleaved_ids, local_entries, padded_size
    n = 1 << scale
    src, dst = synthetic.rmat_edges_host(scale, 8, seed=3)
    total = len(src)
    n_global = padded_size(n, world)
    n_local = n_global // world
    # pass 1: each rank counts its slice of the edge stream, then all-reduce
    per = (total + world - 1) // world
    lo, hi = rank * per, min((rank + 1) * per, total)
    deg = torch.zeros(n_global, dtype=torch.int32)
    ones = torch.ones(hi - lo, dtype=torch.int32)
    deg.index_add_(0, torch.from_numpy(src[lo:hi]).long(), ones)
    deg.index_add_(0, torch.from_numpy(dst[lo:hi]).long(), ones)
    dist.all_reduce(deg)
    new_id = interleaved_ids(deg, world)
    rows, cols = local_entries(torch.from_numpy(src), torch.from_numpy(dst), new_id, rank, n_local)
    A_loc = sp.coo_matrix((np.ones(len(rows)), (rows.numpy(), cols.numpy())), shape=(n_local, n_global)).tocsr()
    A_loc.sum_duplicates()
    A_loc.data[:] = 1.0
    # row-partitioned PPR with the S-in-advance normaliser (mirrors DistPageRank + finalize_state)
    alpha, tol = 0.85, 1e-9
    d_loc = np.asarray(A_loc.sum(axis=1)).ravel()
    sq = np.where(d_loc > 0, np.sqrt(d_loc), 1.0)
    R_loc = np.where(d_loc > 0, 1.0 / np.sqrt(np.maximum(d_loc, 1)), 0.0)
    R_full = torch.empty(n_global, dtype=torch.float64)
    dist.all_gather_into_tensor(R_full, torch.from_numpy(R_loc))
    degM = R_loc * (A_loc @ R_full.numpy())
    c = sq * degM
    seeds = synthetic.seed_sets(n, 1, 10, seed=0)[0]
    p_user = np.zeros(n_global)
    p_user[seeds] = 1.0
    p_int = np.zeros(n_global)
    p_int[new_id.numpy()] = p_user
    off = rank * n_local
    pn = p_int[off:off + n_local] / 10.0
    z = pn / sq
    q = (1 - alpha) * pn / sq
    acc = torch.tensor([float((z * c).sum()), float((q * sq).sum())], dtype=torch.float64)
    dist.all_reduce(acc)
    invS = 1.0 / (alpha * acc[0].item() + acc[1].item())
    bias = acc[1].item()
    # masked exchange (pgb_peers.row_mask): a rank's buffer receives a new value only where the rank reads it;
    # everything else stays NaN here, so a missing reader bit would poison the result
    from pygrank_b200.dist import reader_mask_from_need
    need = torch.zeros(n_global, dtype=torch.uint8)
    need[torch.from_numpy(A_loc.indices.astype(np.int64))] = 1
    need[off:off + n_local] = 1
    mask = reader_mask_from_need(need, rank, world, n_local)
    assert bool(((mask >> rank) & 1).all())
    all_masks = torch.empty(n_global, dtype=torch.int32)
    dist.all_gather_into_tensor(all_masks, mask)
    reads_me = ((all_masks >> rank) & 1).bool().numpy()           # entries of the full vector this rank receives
    assert np.array_equal(reads_me, need.numpy().astype(bool))

    def exchange(local):
        full = torch.empty(n_global, dtype=torch.float64)
        dist.all_gather_into_tensor(full, torch.from_numpy(np.ascontiguousarray(local)))
        out = np.full(n_global, np.nan)
        out[reads_me] = full.numpy()[reads_me]
        return torch.from_numpy(out)

    zfull = exchange(z)
    iteration, steps = 1, 0
    w = np.where(d_loc > 0, 1.0 / np.maximum(d_loc, 1), 0.0)
    while True:
        znew = (alpha * w * (A_loc @ zfull.numpy()) + q) * invS
        zold = zfull.numpy()[off:off + n_local]
        acc = torch.tensor([float((znew * c).sum()), float((sq * np.abs(znew - zold)).sum())], dtype=torch.float64)
        zfull = exchange(znew)
        dist.all_reduce(acc)
        steps += 1
        iteration = steps + 1
        if acc[1].item() / n <= tol:
            break
        invS = 1.0 / (alpha * acc[0].item() + bias)
    scores_int = zfull.numpy() * np.concatenate([np.empty(0)] + [None] * 0) if False else None
    sq_full = torch.empty(n_global, dtype=torch.float64)
    dist.all_gather_into_tensor(sq_full, torch.from_numpy(sq))
    zall = torch.empty(n_global, dtype=torch.float64)             # results: every rank's own (always valid) slice
    dist.all_gather_into_tensor(zall, zfull[off:off + n_local].clone())
    assert not np.isnan(zall.numpy()).any()
    scores_user = (zall.numpy() * sq_full.numpy() * 10.0)[new_id.numpy()][:n]
    gathered = [None] * world
    dist.all_gather_object(gathered, (A_loc.indptr, A_loc.indices, A_loc.nnz))
    if rank == 0:
        np.savez(out, new_id=new_id.numpy(), scores=scores_user, iteration=iteration,
                 nnz=np.array([g[2] for g in gathered]),
                 **{f"indptr{r}": gathered[r][0] for r in range(world)},
                 **{f"indices{r}": gathered[r][1] for r in range(world)})
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_ppr(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle import reference_port as orc
    from pygrank_b200 import synthetic
    world, scale = 2, 10
    out = str(tmp_path / "dist.npz")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, scale, out), nprocs=world, join=True)
    z = np.load(out)
    n = 1 << scale
    new_id = z["new_id"]
    assert sorted(new_id.tolist()) == list(range(n)), "the relabelling must be a permutation"
    A = synthetic.rmat_graph_host(scale, 8, seed=3)
    P = sp.coo_matrix((np.ones(n), (new_id, np.arange(n))), shape=(n, n)).tocsr()
    A_int = (P @ A @ P.T).tocsr()
    A_int.sort_indices()
    n_local = n // world
    for r in range(world):                       # every rank holds exactly its rows of the relabelled graph
        block = A_int[r * n_local:(r + 1) * n_local]
        assert np.array_equal(block.indptr, z[f"indptr{r}"])
        assert np.array_equal(block.indices, z[f"indices{r}"])
    nnz = z["nnz"]
    assert abs(int(nnz[0]) - int(nnz[1])) <= 0.02 * nnz.sum(), nnz   # degree-interleaving balances the ranks
    M = orc.to_sparse_matrix(A, "symmetric", False)
    p = np.zeros(n)
    p[synthetic.seed_sets(n, 1, 10, seed=0)[0]] = 1.0
    ref, iters, _ = orc.pagerank(M, p, 0.85, tol=1e-9, max_iters=1000)
    assert int(z["iteration"]) == iters
    assert np.abs(z["scores"] - ref).sum() / np.abs(ref).sum() < 1e-10


def test_virtual_columns_are_a_bijection_with_contiguous_hub_blocks():
    """Host arithmetic of the multi-GPU hsell form (pygrank_b200.dist.virtual_columns): every hub block
    is one contiguous virtual range holding the same positions of every rank, rank-major; the map is
    a bijection and real_columns (mirrored by hsell_real_col in csrc/hsell.cu) inverts it."""
    import torch
    from pygrank_b200.dist import real_columns, virtual_columns
    for world, n_local, H, K in [(2, 1000, 64, 5), (4, 4096, 256, 16), (8, 777, 64, 0), (2, 640, 128, 10), (1, 500, 64, 3)]:
        n = world * n_local
        cols = torch.arange(n)
        v = virtual_columns(cols, n_local, world, H, K)
        assert sorted(v.tolist()) == list(range(n)) or (K * H > 0 and v.max() < max(n, K * H + world * max(n_local - K * (H // world), 0)))
        assert len(set(v.tolist())) == n
        assert torch.equal(real_columns(v, n_local, world, H, K), cols)
        hs = H // world
        for b in range(K):
            inside = (v >= b * H) & (v < (b + 1) * H)
            pos = cols[inside] % n_local
            assert int(inside.sum()) == world * hs and int(pos.min()) == b * hs and int(pos.max()) == (b + 1) * hs - 1
            # rank-major inside the block: local index = rank*hs + (pos - b*hs)
            rnk = cols[inside] // n_local
            assert torch.equal(v[inside] - b * H, rnk * hs + (pos - b * hs))
        tail = v >= K * H
        assert int(tail.sum()) == world * (n_local - K * hs)


def test_column_shards_partition_the_columns():
    from pygrank_b200.dist import column_shard
    for B in (0, 1, 7, 8, 256, 257):
        for world in (1, 2, 3, 8):
            shards = [column_shard(B, r, world) for r in range(world)]
            assert [c for sh in shards for c in sh] == list(range(B))
            assert max(len(sh) for sh in shards) - min(len(sh) for sh in shards) <= 1


class _FakeFilter:
    """Stands in for a pygrank_b200 filter on CPU: propagate() doubles the features, one "iteration" per column index."""

    class _CM:
        iterations = []

    def __init__(self):
        self.convergence = self._CM()

    def propagate(self, graph, features):
        self.convergence.iterations = [graph + c for c in range(features.shape[1])]
        return features * 2.0


def _shard_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pygrank_b200.dist import column_shard, propagate_sharded
    feats = torch.arange(5 * 7, dtype=torch.float64).reshape(5, 7)
    full, its = propagate_sharded(_FakeFilter(), 100, feats, gather=True)
    local, lits = propagate_sharded(_FakeFilter(), 100, feats, gather=False)
    mine = column_shard(7, rank, world)
    ok = torch.equal(full, feats * 2.0) and torch.equal(local, feats[:, mine.start:mine.stop] * 2.0)
    ok = ok and its == [100 + c for r in range(world) for c in range(len(column_shard(7, r, world)))]
    ok = ok and lits == [100 + c for c in range(len(mine))]
    flags = [None] * world
    dist.all_gather_object(flags, bool(ok))
    if rank == 0:
        np.save(out, np.array(flags))
    dist.barrier()
    dist.destroy_process_group()


def test_propagate_sharded_two_ranks(tmp_path):
    world = 2
    out = str(tmp_path / "shard.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_shard_worker, args=(world, port, out), nprocs=world, join=True)
    assert np.load(out).all()
