"""Device twins of the generators in ``synthetic.py`` (kernels in csrc/graphgen.cu)."""
from __future__ import annotations

import torch

from . import _capi as C
from .graph import DeviceGraph, _dev
from .synthetic import rmat_thresholds


def rmat_edges_device(scale: int, edge_factor: int = 16, seed: int = 1, a=0.57, b=0.19, c=0.19,
                      first_edge: int = 0, num_edges: int | None = None, device=None):
    dev = _dev(device)
    total = edge_factor << scale
    if num_edges is None:
        num_edges = total - first_edge
    t1, t2, t3 = rmat_thresholds(a, b, c)
    src = torch.empty(num_edges, dtype=torch.int32, device=dev)
    dst = torch.empty(num_edges, dtype=torch.int32, device=dev)
    C.check(C.lib().pgb_rmat_edges(scale, first_edge, num_edges, seed, t1, t2, t3, C.ptr(src), C.ptr(dst),
                                   C.stream_ptr()))
    return src, dst


def ba_edges_device(n: int, m: int, seed: int = 1, device=None):
    dev = _dev(device)
    num = (n - m) * m
    src = torch.empty(num, dtype=torch.int32, device=dev)
    dst = torch.empty(num, dtype=torch.int32, device=dev)
    C.check(C.lib().pgb_ba_edges(n, m, 0, num, seed, C.ptr(src), C.ptr(dst), C.stream_ptr()))
    return src, dst


def rmat_graph_device(scale: int, edge_factor: int = 16, seed: int = 1, normalization: str = "symmetric",
                      relabel: str = "hub", device=None) -> DeviceGraph:
    """Undirected RMAT graph (symmetrised, self loops dropped, duplicates collapsed to weight 1)."""
    src, dst = rmat_edges_device(scale, edge_factor, seed, device=device)
    return DeviceGraph.from_edges(1 << scale, src, dst, directed=False, drop_self_loops=True, binary=True,
                                  normalization=normalization, relabel=relabel)


def ba_graph_device(n: int, m: int, seed: int = 1, normalization: str = "symmetric", relabel: str = "hub",
                    device=None) -> DeviceGraph:
    src, dst = ba_edges_device(n, m, seed, device=device)
    return DeviceGraph.from_edges(n, src, dst, directed=False, drop_self_loops=True, binary=True,
                                  normalization=normalization, relabel=relabel)
