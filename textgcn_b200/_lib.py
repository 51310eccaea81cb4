"""ctypes binding of libtgcn_b200.so (the C ABI declared in include/tgcn_b200.h).

There is no CPU fallback: if the library has not been built, importing a compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int32, c_int64, c_void_p

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "libtgcn_b200.so")

ABI_VERSION = 3
MAX_TOPK = 128
ADV_MAX_CANDIDATES = 2048

_P = c_void_p  # device / host pointers travel as integers

# name -> (restype, argtypes); every symbol include/tgcn_b200.h declares
SIGNATURES = {
    "tgcn_abi_version": (c_int32, []),
    "tgcn_last_error": (c_char_p, []),
    "tgcn_graph_create": (c_int32, [POINTER(c_void_p), c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "tgcn_graph_create_block": (c_int32, [POINTER(c_void_p), c_int64, c_int64, c_int64, c_int64, c_int64, _P, _P, _P, _P]),
    "tgcn_graph_build_transpose_perm": (c_int32, [_P, _P]),
    "tgcn_graph_destroy": (None, [_P]),
    "tgcn_graph_num_segments": (c_int64, [_P]),
    "tgcn_graph_set_mask_col_offset": (c_int32, [_P, c_int64]),
    "tgcn_propagate_workspace_bytes": (c_int64, [_P, c_int64, c_int32]),
    "tgcn_spmm_fwd": (c_int32, [_P, c_int64, _P, _P, _P, c_int64, _P]),
    "tgcn_spmm_ex": (c_int32, [_P, c_int64, _P, _P, _P, c_float, c_int32, c_int32, POINTER(c_void_p), POINTER(c_void_p),
                               c_float, c_int32, _P, _P, c_int64, _P]),
    "tgcn_layer_mean": (c_int32, [c_int64, c_int32, POINTER(c_void_p), c_float, _P, _P]),
    "tgcn_propagate_fwd": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, c_float, _P, _P, c_int64, _P]),
    "tgcn_propagate_sliced": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, c_float, c_int64, c_int64, c_int32, c_int64,
                                        POINTER(c_void_p), POINTER(c_void_p), _P, c_int64, _P]),
    "tgcn_spmm_scatter": (c_int32, [_P, c_int64, _P, _P, c_int32, POINTER(c_void_p), POINTER(c_void_p), c_float, c_int64, c_int64,
                                    c_int32, c_int64, c_int64, POINTER(c_void_p), POINTER(c_void_p), _P, c_int64, _P]),
    "tgcn_layer_mean_scatter": (c_int32, [c_int64, c_int64, c_int32, POINTER(c_void_p), c_float, c_int64, c_int64, c_int64, c_int32,
                                          POINTER(c_void_p), _P]),
    "tgcn_peer_alloc": (c_int32, [c_int64, POINTER(c_void_p), _P]),
    "tgcn_peer_open": (c_int32, [_P, POINTER(c_void_p)]),
    "tgcn_peer_close": (c_int32, [_P]),
    "tgcn_peer_free": (c_int32, [_P]),
    "tgcn_peer_barrier": (c_int32, [c_int32, c_int32, POINTER(c_void_p), c_int64, _P]),
    "tgcn_comm_unique_id": (c_int32, [_P]),
    "tgcn_comm_init_rank": (c_int32, [POINTER(c_void_p), c_int32, c_int32, _P]),
    "tgcn_comm_destroy": (c_int32, [_P]),
    "tgcn_allreduce_sum_f32": (c_int32, [_P, _P, c_int64, _P]),
    "tgcn_allgather_f32": (c_int32, [_P, _P, _P, c_int64, _P]),
    "tgcn_topk_exchange": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, _P, _P, _P]),
    "tgcn_propagate_bwd": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, c_float, c_int32, _P, _P, c_int64, _P]),
    "tgcn_propagate_host": (c_int32, [_P, c_int64, c_int32, c_int32, _P, _P, _P, _P, _P, c_int64, _P]),
    "tgcn_bpr_workspace_bytes": (c_int64, [c_int64]),
    "tgcn_bpr_fwd_bwd": (c_int32, [c_int64, c_int64, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, c_float,
                                   _P, _P, _P, _P, c_int64, _P]),
    "tgcn_eval_workspace_bytes": (c_int64, [c_int64, c_int64, c_int64, c_int32]),
    "tgcn_eval_screen_queue_offset": (c_int64, [c_int64, c_int64, c_int64, c_int32, c_int32]),
    "tgcn_eval_resolve_precision": (c_int32, [c_int64, c_int64, c_int32, c_int32, c_int32]),
    "tgcn_eval_topk": (c_int32, [_P, c_int64, _P, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P,
                                 c_int32, c_int32, c_int32, c_int32, _P, _P, _P, c_int64, _P]),
    "tgcn_topk_merge": (c_int32, [_P, c_int64, _P, c_int32, c_int32, _P, _P, c_int32, _P, _P, _P]),
    "tgcn_adv_select": (c_int32, [_P, c_int64, c_int64, c_int32, _P, _P, _P, c_int32, _P, _P, _P, _P]),
    "tgcn_ltr_pairwise_features": (c_int32, [c_int64, c_int64, c_int64, c_int64, c_int32, _P, _P, _P, _P, _P, _P, _P,
                                             _P, _P, _P, _P]),
    "tgcn_ltr_pairwise_emb_bwd": (c_int32, [c_int64, c_int64, c_int64, _P, _P, _P, _P, _P, _P]),
    "tgcn_ltr_pack_items": (c_int32, [c_int64, c_int64, c_int64, _P, _P, _P, POINTER(c_float), _P, _P]),
    "tgcn_ltr_pack_users": (c_int32, [c_int64, _P, c_int64, c_int64, _P, _P, _P, _P, _P]),
    "tgcn_score_batchwise": (c_int32, [c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, _P, _P, c_int64, c_int64, _P]),
    "tgcn_score_pairwise_adv": (c_int32, [c_int64, c_int32, c_int64, _P, c_int64, _P, _P, _P]),
    "tgcn_ltr_features_rows": (c_int32, [c_int64, c_int64, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64, _P, c_int64,
                                         _P, c_int64, _P, c_int64, _P]),
    "tgcn_topk_metrics_workspace_bytes": (c_int64, []),
    "tgcn_topk_metrics": (c_int32, [c_int64, c_int32, _P, _P, _P, c_int32, POINTER(c_int32), _P, _P, c_int64, _P]),
    "tgcn_sample_bpr_batch": (c_int32, [_P, c_int64, c_int32, _P, ctypes.c_uint64, c_int32, _P, _P, _P]),
    "tgcn_sample_candidates": (c_int32, [c_int64, c_int64, c_int32, _P, ctypes.c_uint64, _P, _P]),
    "tgcn_dropout_mask": (c_int32, [c_int64, c_float, ctypes.c_uint64, _P, _P]),
    "tgcn_sample_positives": (c_int32, [_P, c_int64, c_int32, _P, ctypes.c_uint64, _P, _P]),
    "tgcn_adam_step": (c_int32, [c_int64, _P, _P, _P, _P, c_float, c_float, c_float, c_float, c_int64, _P]),
    "tgcn_adam_prepare": (c_int32, [_P, _P, c_float, c_float, _P]),
    "tgcn_adam_step_dev": (c_int32, [c_int64, _P, _P, _P, _P, c_float, c_float, c_float, c_float, _P, _P]),
    "tgcn_counter_inc": (c_int32, [_P, _P]),
    "tgcn_dropout_mask_dev": (c_int32, [c_int64, c_float, ctypes.c_uint64, _P, _P, _P]),
}

_lib = None


class TgcnError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once) and attach the signatures.  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("TGCN_B200_LIB") or LIB_PATH   # TGCN_B200_LIB: the debug-assert flavour (build.py --debug), tests only
    if not os.path.exists(path):
        raise TgcnError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(textgcn_b200 has no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    got = lib.tgcn_abi_version()
    if got != ABI_VERSION:
        raise TgcnError(f"libtgcn_b200 ABI version {got} != expected {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        raise TgcnError(load().tgcn_last_error().decode("utf-8", "replace"))
