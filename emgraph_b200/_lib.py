"""ctypes binding of include/kge_b200.h.  There is no CPU fallback: a missing library or a failing
call raises.  torch tensors only carry the buffers (``tensor.data_ptr()``)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libkge_b200.so")

KGE_MAX_SHARDS = 8
ABI_VERSION = 6

MODEL_IDS = {"TransE": 0, "TransE_L2": 1, "DistMult": 2, "ComplEx": 3, "HolE": 4}
LOSS_IDS = {"pairwise": 0, "nll": 1, "multiclass_nll": 2, "absolute_margin": 3, "self_adversarial": 4}
OPT_IDS = {"adam": 0, "adagrad": 1, "momentum": 2, "sgd": 3}
TRAIN_SIDE_IDS = {"s,o": 0, "s+o": 0, "s": 1, "o": 2}
RANK_SIDE_IDS = {"s,o": 0, "s+o": 1, "s": 2, "o": 3}
STRATEGY_IDS = {"worst": 0, "best": 1, "middle": 2}
NL_IDS = {"linear": 0, "tanh": 1, "sigmoid": 2, "softplus": 3}
F_RESET_STATE = 1
F_NO_UPDATE = 2
F_PIPELINE = 4


class KgeTable(C.Structure):
    _fields_ = [
        ("shard", C.c_void_p * KGE_MAX_SHARDS),
        ("rows", C.c_int64),
        ("rows_per_shard", C.c_int64),
        ("n_shards", C.c_int32),
        ("K", C.c_int32),
    ]


class KgeTrainArgs(C.Structure):
    _fields_ = [
        ("model", C.c_int32), ("loss", C.c_int32), ("opt", C.c_int32), ("side", C.c_int32),
        ("flags", C.c_uint32),
        ("k", C.c_int32), ("eta", C.c_int32),
        ("margin", C.c_float),
        ("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("momentum", C.c_float),
        ("seed", C.c_uint64), ("step", C.c_uint64), ("neg_index_base", C.c_uint64),
        ("ent", KgeTable), ("ent_m", KgeTable), ("ent_v", KgeTable),
        ("rel", C.c_void_p), ("rel_m", C.c_void_p), ("rel_v", C.c_void_p),
        ("R", C.c_int64),
        ("pos", C.c_void_p), ("n_pos", C.c_int64),
        ("repl", C.c_void_p), ("keep_subj", C.c_void_p),
        ("loss_out", C.c_void_p), ("dbg_scores", C.c_void_p), ("dbg_grad_ent", C.c_void_p), ("dbg_grad_rel", C.c_void_p),
        ("stage", C.c_void_p), ("grad_tails", C.c_void_p), ("grad_tail_stride", C.c_int64),
        ("alpha", C.c_float),
        ("reg_p", C.c_int32), ("reg_lambda_ent", C.c_float), ("reg_lambda_rel", C.c_float),
        ("neg_entities", C.c_void_p), ("neg_entities_n", C.c_int64),
        ("non_linearity", C.c_int32),
        ("k_model", C.c_int32),
    ]


# every symbol include/kge_b200.h declares: name -> (restype, argtypes)
_P, _I, _L = C.c_void_p, C.c_int, C.c_int64
SYMBOLS = {
    "kge_abi_version": (_I, []),
    "kge_last_error": (C.c_char_p, []),
    "kge_has_tensor_core_rank": (_I, []),
    "kge_train_grad_floats": (_L, [_I, _L, _I]),
    "kge_train_grad_head_floats": (_L, [_I, _L, _I]),
    "kge_train_push_rows": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _L, C.POINTER(KgeTable), _L, _L, _P]),
    "kge_ctx_create": (_I, [_I, C.POINTER(_P)]),
    "kge_ctx_destroy": (_I, [_P]),
    "kge_ctx_workspace_bytes": (_L, [_P]),
    "kge_ctx_set_timing": (_I, [_P, _I]),
    "kge_ctx_get_timing": (_I, [_P, C.POINTER(C.c_float), C.POINTER(_I)]),
    "kge_ctx_get_timing_ex": (_I, [_P, C.POINTER(C.c_float), _I, C.POINTER(_I)]),
    "kge_sort_entries": (_I, [_P, _P, _L, _L, _I, _P, _P]),
    "kge_score": (_I, [_P, _I, _I, C.POINTER(KgeTable), _P, _L, _P, _L, _P, _P]),
    "kge_predict": (_I, [_P, _I, _I, C.POINTER(KgeTable), _P, _L, _P, _L, _I, _P, _P]),
    "kge_train_step": (_I, [_P, C.POINTER(KgeTrainArgs), _P]),
    "kge_train_emit": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _P]),
    "kge_train_fwd_bwd": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _P]),
    "kge_train_apply": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _L, C.POINTER(KgeTable), _L, _L, _P]),
    "kge_train_partial": (_I, [_P, C.POINTER(KgeTrainArgs), _L, _L, _P, _P]),
    "kge_train_partial_sorted": (_I, [_P, C.POINTER(KgeTrainArgs), _I, _P, _P]),
    "kge_train_backward": (_I, [_P, C.POINTER(KgeTrainArgs), _L, _L, _P, _P]),
    "kge_train_reduce": (_I, [_P, C.POINTER(KgeTrainArgs), _P]),
    "kge_allreduce_p2p": (_I, [_P, C.POINTER(KgeTable), C.POINTER(KgeTable), C.POINTER(KgeTable), _I, _L, _L, C.c_uint32, _P]),
    "kge_train_step_host": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _P, _P]),
    "kge_train_step_host_async": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _P, _P, C.POINTER(_I)]),
    "kge_train_host_wait": (_I, [_P, _I]),
    "kge_train_select": (_I, [_P, C.POINTER(KgeTrainArgs), _P, _L, _L, _L, _P]),
    "kge_normalize_rows": (_I, [_P, _P, _L, _I, _P]),
    "kge_filter_build": (_I, [_P, _P, _L, _L, _L, _P]),
    "kge_filter_clear": (_I, [_P]),
    "kge_filter_size_sync": (_L, [_P]),
    "kge_rank_counts": (_I, [_P, _I, _I, C.POINTER(KgeTable), _P, _L, _P, _L, _L, _P, _L, _I, _I, _I, _I, _P, _P]),
    "kge_rank_counts_rows": (_I, [_P, _I, _I, _L, _P, _L, _P, _P, _P, _L, _L, _P, _L, _I, _I, _I, _I, _P, _P]),
    "kge_rank_finalize": (_I, [_P, _P, _L, _I, _I, _I, _P, _P, _P]),
    "kge_rank_host": (_I, [_P, _I, _I, C.POINTER(KgeTable), _P, _L, _P, _L, _I, _I, _I, _I, _I, _P, _P]),
    "kge_dev_alloc": (_I, [_L, C.POINTER(_P)]),
    "kge_dev_free": (_I, [_P]),
    "kge_ipc_export": (_I, [_P, _P]),
    "kge_ipc_open": (_I, [_P, C.POINTER(_P)]),
    "kge_ipc_close": (_I, [_P]),
    "kge_enable_peer_access": (_I, [_I, _I]),
}

_lib = None


class KgeError(RuntimeError):
    pass


def load():
    """dlopen the in-tree library and bind every symbol; raises if it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KgeError(
            "libkge_b200.so not found at %s -- build it with `python -m emgraph_b200.build` "
            "(or __graft_entry__.build()); emgraph_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = res
        fn.argtypes = args
    v = lib.kge_abi_version()
    if v != ABI_VERSION:
        raise KgeError("libkge_b200.so ABI %d != binding ABI %d; rebuild" % (v, ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise KgeError(load().kge_last_error().decode("utf-8", "replace") or ("kge error %d" % rc))
