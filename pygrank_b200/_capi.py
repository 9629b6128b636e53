"""ctypes binding of libpgb200.so (the C-ABI declared in include/pgb200.h).

The library is the product: if it is missing or a call fails this module raises — there is
no CPU or PyTorch fallback anywhere in the package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libpgb200.so")
CSRC = os.path.join(_HERE, "csrc")

PGB_F32, PGB_F64 = 0, 1
BUILD_SYMMETRIZE, BUILD_DROP_SELF_LOOPS, BUILD_BINARY = 1, 2, 4
SCALE_ONE, SCALE_RECIP, SCALE_RSQRT = 0, 1, 2
ERR_MABS, ERR_L1, ERR_MSQ, ERR_ITERS, ERR_MAX = 0, 1, 2, 3, 4
RUNNING, CONVERGED, MAX_ITERS = 0, 1, 2
# state_f64 / state_i32 slots
SF_ALPHA, SF_BIAS, SF_INVS, SF_TACC, SF_EACC, SF_TOL, SF_MEAN, SF_LASTERR, SF_NORM, SF_PSUM, SF_AMUL = range(11)
SI_TICKET, SI_STEPS, SI_STOP, SI_ITERATION, SI_MAX_ITERS, SI_END_MODULO, SI_ERR_MODE, SI_QUOTIENT = range(8)
STATE_LEN = 16
ABI_VERSION = 4
SIGNATURE_WORDS = 8
HSELL_MAX_WINDOWS = 16
HSELL_CHUNK = 32   # PGB_HSELL_CHUNK: rounds per chunk of the hsell streams


class Csr(Structure):
    _fields_ = [("n", c_int64), ("nnz", c_int64), ("indptr", c_void_p), ("indices", c_void_p), ("values", c_void_p),
                ("tile_row", c_void_p), ("n_tiles", c_int32), ("tile_items", c_int32), ("istream", c_void_p),
                ("vstream", c_void_p), ("hsell", c_void_p)]


class PanelJob(Structure):
    """pgb_panel_job (include/pgb200.h): seed columns scheduled over the slots of a hub-blocked panel."""
    _fields_ = [("n_cols", c_int32), ("hist_stride", c_int32), ("cols", c_void_p), ("row_stride", c_int64),
                ("col_stride", c_int64), ("out", c_void_p), ("out_row_stride", c_int64), ("out_col_stride", c_int64),
                ("perm", c_void_p),
                ("sq", c_void_p), ("coefvec", c_void_p), ("col_params", c_void_p), ("alpha", c_double),
                ("alpha_s", c_double), ("coef", c_double), ("tol", c_double), ("mean", c_double),
                ("max_iters", c_int32), ("end_modulo", c_int32), ("err_mode", c_int32), ("quotient", c_int32),
                ("preserve_norm", c_int32), ("sched", c_void_p), ("slot_col", c_void_p), ("slot_plan", c_void_p),
                ("plan_norm", c_void_p), ("col_result", c_void_p), ("col_err", c_void_p), ("poly", c_int32),
                ("reserved0", c_int32), ("ranks", c_void_p), ("coef_table", c_void_p)]


class Hsell(Structure):
    """pgb_hsell (include/pgb200.h): hub-blocked sliced-ELL form of a pull CSR (edge values optional)."""
    _fields_ = [("n_rows", c_int64), ("n_slices", c_int64), ("n_partials", c_int64), ("seg_len", c_int64),
                ("n_segments", c_int32), ("block_cols", c_int32), ("n_blocks", c_int32), ("n_ctas", c_int32),
                ("n_hub_chunks", c_int32), ("n_tail_chunks", c_int32), ("n_heavy", c_int32), ("heavy_parts", c_int32),
                ("n_reduce", c_int32), ("reserved0", c_int32),
                ("hub_chunks", c_void_p), ("tail_chunks", c_void_p), ("hub_words", c_void_p), ("tail_cols", c_void_p),
                ("piece_row", c_void_p), ("upd_rows", c_void_p), ("heavy_slices", c_void_p),
                ("reduce_items", c_void_p),
                ("block_chunk_begin", c_void_p), ("cta_hub_begin", c_void_p), ("cta_tail_begin", c_void_p),
                ("piece_slice", c_void_p), ("hub_vals", c_void_p), ("tail_vals", c_void_p)]


MAX_PEERS = 16


class Peers(Structure):
    """pgb_peers (include/pgb200.h): peer-mapped buffers of the fused multi-GPU exchange."""
    _fields_ = [("n", c_int32), ("rank", c_int32), ("zbuf0", c_void_p * MAX_PEERS), ("zbuf1", c_void_p * MAX_PEERS),
                ("mc_zbuf0", c_void_p), ("mc_zbuf1", c_void_p), ("acc", c_void_p * MAX_PEERS), ("row_mask", c_void_p)]


class SpanWs(Structure):
    _fields_ = [("acc", c_void_p), ("cnt", c_void_p), ("partials", c_void_p), ("yacc", c_void_p)]


_SIGNATURES = {
    "pgb_abi_version": (c_int, []),
    "pgb_last_error": (c_char_p, []),
    "pgb_tile_items": (c_int, []),
    "pgb_device_sm_count": (c_int, [c_int]),
    "pgb_set_kernel_variant": (c_int, [c_int]),
    "pgb_hsell_max_block_cols": (c_int, [c_int]),
    "pgb_hsell_set_tail_warps": (c_int, [c_int]),
    "pgb_hsell_set_dropout": (c_int, [c_double, c_uint64]),
    "pgb_hsell_count": (c_int, [c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_double, c_int32, c_int64,
                                c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p]),
    "pgb_hsell_fill": (c_int, [c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_int32, c_int32, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_build_item_stream": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p]),
    "pgb_gather_probe": (c_int, [POINTER(Csr), c_int, c_void_p, c_void_p, c_void_p]),
    "pgb_rmat_edges": (c_int, [c_int, c_int64, c_int64, c_uint64, c_uint32, c_uint32, c_uint32, c_void_p, c_void_p,
                               c_void_p]),
    "pgb_ba_edges": (c_int, [c_int64, c_int, c_int64, c_int64, c_uint64, c_void_p, c_void_p, c_void_p]),
    "pgb_csr_build_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int]),
    "pgb_csr_build": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_size_t, c_void_p,
                              c_void_p, c_void_p, POINTER(c_int64), c_void_p]),
    "pgb_csr_expand_rows": (c_int, [c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "pgb_degree_order_workspace_bytes": (c_size_t, [c_int64]),
    "pgb_degree_order": (c_int, [c_int64, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "pgb_relabel_coo": (c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_hub_order_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "pgb_hub_order": (c_int, [c_int64, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_size_t, c_void_p, c_void_p,
                              c_void_p]),
    "pgb_mergepath_partition": (c_int, [c_int64, c_int64, c_void_p, c_int32, c_void_p, c_void_p]),
    "pgb_csr_row_sums": (c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_make_scales": (c_int, [c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "pgb_csr_normalized_values": (c_int, [c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                          c_void_p]),
    "pgb_csr_row_sums_numpy": (c_int, [c_int64, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "pgb_spmv": (c_int, [POINTER(Csr), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, SpanWs, c_void_p]),
    "pgb_affine_steps": (c_int, [POINTER(Csr), c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int64, c_void_p, c_void_p, c_void_p, SpanWs, c_int, c_int, c_int,
                                 c_void_p]),
    "pgb_poly_steps": (c_int, [POINTER(Csr), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int64, c_void_p, c_void_p, c_void_p, SpanWs, c_int, c_int, c_int, c_void_p]),
    "pgb_panel_width": (c_int, [c_int]),
    "pgb_affine_steps_batched": (c_int, [POINTER(Csr), c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int32, SpanWs,
                                         c_int, c_int, c_void_p]),
    "pgb_hsell_panel_width": (c_int, [c_int]),
    "pgb_hsell_panel_block_cols": (c_int, []),
    "pgb_affine_steps_panel": (c_int, [POINTER(Hsell), c_void_p, c_int, POINTER(PanelJob), c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_void_p]),
    "pgb_panel_stage": (c_int, [c_int64, c_int, c_int, c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int32, c_void_p,
                                c_void_p]),
    "pgb_state_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_affine_step_peer": (c_int, [POINTER(Csr), c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                     c_void_p, c_int64, c_void_p, c_void_p, c_void_p, SpanWs, c_int, POINTER(Peers),
                                     c_void_p]),
    "pgb_poly_step_peer": (c_int, [POINTER(Csr), c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int64, c_void_p, c_void_p, c_void_p, SpanWs, c_int, POINTER(Peers), c_void_p]),
    "pgb_state_finalize_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p]),
    "pgb_affine_init_peer": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                     c_void_p, c_int64, c_void_p, c_void_p, POINTER(Peers), c_void_p]),
    "pgb_scale": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "pgb_unscale": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p, c_void_p]),
    "pgb_reduce3": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_affine_init": (c_int, [c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p, c_void_p,
                                c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "pgb_affine_init_finish": (c_int, [c_void_p, c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None

# number of kernels of OURS enqueued since import (bench.py reports the delta over the timed region)
LAUNCHES = [0]


def count_launches(k: int) -> None:
    LAUNCHES[0] += int(k)


def build(verbose: bool = False) -> str:
    """Compile libpgb200.so for sm_100a with nvcc (in-tree; needs no GPU)."""
    res = subprocess.run(["make", "-C", CSRC, "-j4"], capture_output=True, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
        print(res.stderr)
    if res.returncode != 0:
        raise Exception("building libpgb200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    return LIB_PATH


def lib() -> ctypes.CDLL:
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Exception(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(pygrank_b200 has no CPU fallback)")
        handle = ctypes.CDLL(os.environ.get("PGB_LIB", LIB_PATH))   # PGB_LIB: A/B builds of the same ABI
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.pgb_abi_version() != ABI_VERSION:
            raise Exception("libpgb200.so ABI version mismatch")
        variant = os.environ.get("PGB_KERNEL_VARIANT")
        if variant:   # A/B timing aid: 3 = item-stream kernel, 4 = hsell when available (default)
            if handle.pgb_set_kernel_variant(int(variant)) != 0:
                raise Exception("pgb200: " + handle.pgb_last_error().decode())
        tail_warps = os.environ.get("PGB_HSELL_TAIL_WARPS")
        if tail_warps:
            if handle.pgb_hsell_set_tail_warps(int(tail_warps)) != 0:
                raise Exception("pgb200: " + handle.pgb_last_error().decode())
        _lib = handle
    return _lib


def check(status: int) -> None:
    """Raise the library's last error (plain Exception, the reference's error style)."""
    if status != 0:
        raise Exception("pgb200: " + lib().pgb_last_error().decode("utf-8", "replace"))


def ptr(t) -> int:
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return 0 if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
