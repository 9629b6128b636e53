"""Generates tests/golden/*.npz by running the REAL reference (pygrank 0.2.12 from
/root/reference, numpy backend) in the build container.  The reference cannot travel to
the GPU box, so its inputs and outputs are committed as small fixtures; the oracle port
(oracle/reference_port.py) and the CUDA path are both checked against them.

Run (build container only):
    python tests/golden/make_golden.py

The reference imports `wget` (pygrank/benchmarks/download.py:3), which is not installed and
there is no network: a one-line stub module is placed on sys.path for the import.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

stub = tempfile.mkdtemp()
with open(os.path.join(stub, "wget.py"), "w") as f:
    f.write("def download(*a, **k):\n    raise RuntimeError('no network')\n")
os.environ["pygrankBackend"] = "numpy"
os.environ["HOME"] = tempfile.mkdtemp()
sys.path.insert(0, stub)
sys.path.insert(0, "/root/reference")

import networkx as nx  # noqa: E402
import scipy  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import pygrank as pg  # noqa: E402

from pygrank_b200 import synthetic  # noqa: E402


def canonical(M):
    M = sp.csr_matrix(M).copy()
    M.sort_indices()
    return M


def graphs():
    G = nx.barabasi_albert_graph(2000, 5, seed=0)
    yield "ba2000", nx.to_scipy_sparse_array(G, dtype=float, format="csr"), False
    yield "rmat10", synthetic.rmat_graph_host(10, 8, seed=3), False           # has isolated vertices
    D = nx.gnp_random_graph(600, 0.01, seed=5, directed=True)
    yield "gnp600d", nx.to_scipy_sparse_array(D, dtype=float, format="csr"), True  # dangling rows/cols
    W = sp.csr_matrix(nx.to_scipy_sparse_array(nx.barabasi_albert_graph(300, 3, seed=7), dtype=float, format="csr"))
    rng = np.random.default_rng(11)
    Wc = sp.triu(W, 1).tocoo()
    w = rng.uniform(0.25, 4.0, size=Wc.nnz)
    Wsym = sp.coo_matrix((np.concatenate([w, w]), (np.concatenate([Wc.row, Wc.col]), np.concatenate([Wc.col, Wc.row]))),
                         shape=W.shape).tocsr()
    yield "weighted300", Wsym, False


def personalizations(n, rng):
    P = np.zeros((n, 4))
    for c in range(3):
        P[rng.choice(n, size=10, replace=False), c] = 1.0
    P[:, 3] = rng.uniform(0, 1, size=n) * (rng.uniform(0, 1, size=n) < 0.05)
    P[0, 3] = 2.5
    return P


def main():
    out = {}
    meta = {"pygrank": "0.2.12", "scipy": scipy.__version__, "numpy": np.__version__, "networkx": nx.__version__}
    for name, A, directed in graphs():
        A = sp.csr_matrix(A)
        A.sort_indices()
        n = A.shape[0]
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31) if False else len(name) * 7919)
        P = personalizations(n, rng)
        store = {"indptr": A.indptr.astype(np.int32), "indices": A.indices.astype(np.int32), "data": A.data,
                 "directed": np.array(directed), "P": P}
        graph = pg.AdjacencyWrapper(A, directed=directed)
        for norm in ["auto", "symmetric", "col", "laplacian", "both", "none"]:
            for renorm in ([False, True] if norm in ("auto", "symmetric") else [False]):
                pre = pg.preprocessor(normalization=norm, renormalize=renorm, assume_immutability=True)
                M = pre(graph)
                C = canonical(M.array)
                key = f"norm_{norm}_{int(renorm)}"
                store[key + "_indptr"] = C.indptr.astype(np.int32)
                store[key + "_indices"] = C.indices.astype(np.int32)
                store[key + "_data"] = C.data
                store[key + "_degrees"] = pg.degrees(M)
        runs = {
            "ppr85": (lambda pre: pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "auto"),
            "ppr90_noq": (lambda pre: pg.PageRank(0.9, tol=1e-9, use_quotient=False, preprocessor=pre, max_iters=1000), "auto"),
            "ppr85_sym": (lambda pre: pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "symmetric"),
            "ppr85_col": (lambda pre: pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "col"),
            "ppr85_tol6_mod3": (lambda pre: pg.PageRank(0.85, tol=1e-6, end_modulo=3, preprocessor=pre), "auto"),
            "ppr85_iters20": (lambda pre: pg.PageRank(0.85, error_type="iters", max_iters=20, preprocessor=pre), "auto"),
            "ppr85_l1": (lambda pre: pg.PageRank(0.85, tol=1e-7, error_type=pg.L1, preprocessor=pre, max_iters=1000), "auto"),
            "ppr85_msq": (lambda pre: pg.PageRank(0.85, tol=1e-16, error_type=pg.MSQ, preprocessor=pre, max_iters=1000), "auto"),
            "heat3": (lambda pre: pg.HeatKernel(3, preprocessor=pre), "auto"),
            "heat3_tol9": (lambda pre: pg.HeatKernel(3, tol=1e-9, preprocessor=pre), "auto"),
            "heat5_sym": (lambda pre: pg.HeatKernel(5, tol=1e-9, preprocessor=pre), "symmetric"),
            "gen40": (lambda pre: pg.GenericGraphFilter([0.9 ** k for k in range(40)], error_type="iters", max_iters=41,
                                                        preprocessor=pre), "auto"),
            "gen3_tol": (lambda pre: pg.GenericGraphFilter([0.5, 0.25, 0.125], tol=1e-9, preprocessor=pre), "auto"),
            "pprclosed": (lambda pre: pg.PageRankClosed(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "auto"),
            "absorb": (lambda pre: pg.AbsorbingWalks(tol=1e-9, preprocessor=pre, max_iters=1000), "auto"),
            "absorb85": (lambda pre: pg.AbsorbingWalks(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "auto"),
            "absorb85_col": (lambda pre: pg.AbsorbingWalks(0.85, tol=1e-9, preprocessor=pre, max_iters=1000), "col"),
        }
        for rname, (make, norm) in runs.items():
            pre = pg.preprocessor(normalization=norm, assume_immutability=True)
            M = pre(graph)
            scores, iters = [], []
            for c in range(P.shape[1]):
                alg = make(pre)
                r = alg(pg.to_signal(M, P[:, c].copy()))
                scores.append(np.asarray(r.np, dtype=np.float64))
                iters.append(alg.convergence.iteration)
            store[f"run_{rname}_scores"] = np.column_stack(scores)
            store[f"run_{rname}_iters"] = np.array(iters, dtype=np.int64)
            print(name, rname, iters)
        # a custom absorption vector
        pre = pg.preprocessor(normalization="auto", assume_immutability=True)
        M = pre(graph)
        absorption = 0.5 + (np.arange(n) % 7) / 7.0
        alg = pg.AbsorbingWalks(0.9, tol=1e-9, preprocessor=pre, max_iters=1000)
        r = alg(pg.to_signal(M, P[:, 0].copy()), absorption=pg.to_signal(M, absorption.copy()))
        store["run_absorb90_custom_scores"] = np.asarray(r.np)
        store["run_absorb90_custom_iters"] = np.array([alg.convergence.iteration])
        store["absorption"] = absorption
        # propagate (signals.py:225-226)
        alg = pg.PageRank(0.85, tol=1e-9, preprocessor=pre, max_iters=1000)
        store["run_propagate_ppr85"] = np.asarray(alg.propagate(M, P.copy()))
        # bare conv
        x = np.random.default_rng(1).uniform(-1, 1, size=n)
        store["conv_x"] = x
        store["conv_y"] = np.asarray(pg.conv(x, M))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **store)
    with open(os.path.join(HERE, "VERSIONS.txt"), "w") as f:
        for k, v in meta.items():
            f.write(f"{k} {v}\n")


if __name__ == "__main__":
    main()
