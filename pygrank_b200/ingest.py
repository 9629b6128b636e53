"""Graph ingestion straight to device COO: the SNAP ``pairs.txt`` format and ``fastgraph.Graph``.

Mirrors /root/reference/pygrank/benchmarks/loader.py:18-88 (``import_snap_format_dataset``) and the edge-list
semantics of /root/reference/pygrank/fastgraph/fastgraph.py:40-78 — node ids in first-appearance order (``add_node``),
an undirected edge stored in both directions, repeated edges ADDING their weights (``coo_array(...).tocsr()`` sums
duplicates), masked edges kept as explicit zeros — but the CSR is built by ``pgb_csr_build`` on the device: no scipy
matrix, no networkx object, no per-edge Python work.  There is no download step (the reference's
``download_dataset`` needs the network).
"""
from __future__ import annotations

import os
from typing import Iterable, Optional, Union

import numpy as np
import torch

from .graph import DeviceGraph, _dev


def read_pairs(path: str, comments: str = "#"):
    """Parse a ``pairs.txt``: one edge per line, two whitespace-separated node names, lines starting with ``#`` (or with
    fewer than two fields) skipped (loader.py:62-67).  Returns (names in first-appearance order, src ids, dst ids)."""
    with open(path, "r", encoding="utf-8") as f:
        text = f.read()
    rows = [ln.split() for ln in text.splitlines() if ln and not ln.startswith(comments)]
    rows = [r for r in rows if len(r) > 1]
    if not rows:
        return [], np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    tokens = np.empty(2 * len(rows), dtype=object)
    tokens[0::2] = [r[0] for r in rows]
    tokens[1::2] = [r[1] for r in rows]
    names, first, inverse = np.unique(tokens.astype(str), return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")            # fastgraph.add_node: ids by first appearance (u before v)
    rank = np.empty(len(names), dtype=np.int64)
    rank[order] = np.arange(len(names))
    ids = rank[inverse]
    return [str(x) for x in names[order]], ids[0::2].copy(), ids[1::2].copy()


def graph_from_pairs(names, src: np.ndarray, dst: np.ndarray, directed: bool = False, normalization: str = "auto",
                     renormalize=False, relabel: str = "hub", device=None) -> DeviceGraph:
    dev = _dev(device)
    n = len(names)
    s = torch.from_numpy(np.ascontiguousarray(src, dtype=np.int32)).to(dev)
    d = torch.from_numpy(np.ascontiguousarray(dst, dtype=np.int32)).to(dev)
    node2id = {v: i for i, v in enumerate(names)}
    # an undirected edge is stored in both directions (a self loop twice, too) and repeated edges add up, exactly like
    # fastgraph.add_edge + coo -> csr; from_edges drops the value array again when every summed weight is 1
    if not directed:
        s, d = torch.cat([s, d]), torch.cat([d, s])
    w = torch.ones(s.numel(), dtype=torch.float64, device=dev)
    return DeviceGraph.from_edges(n, s, d, w, directed=directed, symmetrize=False, drop_self_loops=False,
                                  binary=False, normalization=normalization, renormalize=renormalize, relabel=relabel,
                                  node2id=node2id)


def from_fastgraph(G, normalization: str = "auto", renormalize=False, relabel: str = "hub", device=None) -> DeviceGraph:
    """A ``pygrank.fastgraph.Graph`` (its edge lists, already holding both directions of undirected edges) -> device."""
    dev = _dev(device)
    n = len(G.node_map)
    row = torch.tensor(G.edge_row, dtype=torch.int32, device=dev)
    col = torch.tensor(G.edge_col, dtype=torch.int32, device=dev)
    w = torch.ones(row.numel(), dtype=torch.float64, device=dev)
    masked = getattr(G, "_masked_out", None)
    if masked:                                            # fastgraph.py:73-76: removed edges stay as explicit zeros
        keep = [0.0 if (u in masked and v in masked[u]) else 1.0 for u, v in zip(G.edge_row, G.edge_col)]
        w = torch.tensor(keep, dtype=torch.float64, device=dev)
    node2id = dict(G.node_map)
    return DeviceGraph.from_edges(n, row, col, w, directed=bool(G.directed), symmetrize=False, drop_self_loops=False,
                                  binary=False, normalization=normalization, renormalize=renormalize, relabel=relabel,
                                  node2id=node2id)


def _select_path(path, dataset):
    paths = [path] if isinstance(path, str) else list(path)
    for p in paths:
        if os.path.isdir(p) and os.path.isdir(os.path.join(p, dataset)):
            return p
    return paths[0]


def import_snap_format_dataset(dataset: str,
                               path: Union[Iterable[str], str] = (os.path.join(os.path.expanduser("~"), ".pygrank/data"), ".", "data"),
                               pair_file: str = "pairs.txt", group_file: Optional[str] = "groups.txt",
                               directed: bool = False, min_group_size: float = 0.01, min_group_id: int = 0,
                               max_group_number: int = 20, prepend_all_nodes: bool = False,
                               normalization: str = "auto", relabel: str = "hub", device=None):
    """loader.py:18-88 with the graph built on the device; returns (DeviceGraph, groups) with ``groups`` the same
    dictionary of node-name lists.  The dataset must already be on disk."""
    path = _select_path(path, dataset)
    names, src, dst = read_pairs(os.path.join(path, dataset, pair_file))
    g = graph_from_pairs(names, src, dst, directed=directed, normalization=normalization, relabel=relabel, device=device)
    known = g._pygrank_node2id
    groups = {}
    if prepend_all_nodes:
        groups[0] = []                                   # loader.py:59-60 lists the (still empty) graph here
    n = len(names)
    if min_group_size < 1:
        min_group_size *= n
    gfile = None if group_file is None else os.path.join(path, dataset, group_file)
    if gfile is not None and os.path.isfile(gfile):
        with open(gfile, "r", encoding="utf-8") as f:
            for line in f:
                if line[0] != "#":
                    group = [item for item in line[:-1].split() if len(item) > 0 and item in known]
                    if len(group) >= min_group_size:
                        if min_group_id > 0:
                            min_group_id -= 1
                            continue
                        groups[len(groups)] = group
                        if len(groups) >= max_group_number:
                            break
    return g, groups
