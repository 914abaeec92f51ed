"""ctypes binding of libncf_b200.so (include/ncf_b200.h).

The library is the product: if it is missing or a call fails this module raises - there is no
CPU or PyTorch fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libncf_b200.so")

NCF_FP32, NCF_BF16_TC = 0, 1
EMB_NONE, EMB_MATERIALIZE, EMB_ADAM_SPARSE, EMB_ADAM_DENSE_EQUIV = 0, 1, 2, 3
MAX_S = 8
HAS_BF16_TC = True

# dense flat layout ids, in ncf_dense_id order, with the reference state_dict key of each
DENSE_KEYS = (
    "mf_norm.weight", "mf_norm.bias", "mlp_norm.weight", "mlp_norm.bias",
    "user_product_attention.q_proj.weight", "user_product_attention.k_proj.weight",
    "user_product_attention.v_proj.weight", "user_product_attention.out_proj.weight",
    "user_product_attention.q_proj.bias", "user_product_attention.k_proj.bias",
    "user_product_attention.v_proj.bias", "user_product_attention.out_proj.bias",
    "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
    "mlp.4.weight", "mlp.4.bias", "mlp.6.weight", "mlp.6.bias",
    "mlp.8.weight", "mlp.8.bias", "mlp.10.weight", "mlp.10.bias",
    "mf_output.weight", "mf_output.bias", "mlp_output.weight", "mlp_output.bias",
    "final.0.weight", "final.0.bias",
)


class NcfError(RuntimeError):
    pass


class Tables(C.Structure):
    _fields_ = [("w", C.c_void_p * 4), ("m", C.c_void_p * 4), ("v", C.c_void_p * 4), ("g", C.c_void_p * 4),
                ("touched", C.c_void_p * 2), ("rows_user", C.c_int64), ("rows_item", C.c_int64), ("status", C.c_void_p)]


MAX_WORLD = 16


class ShardPlan(C.Structure):
    """ncf_shard_plan of include/ncf_b200.h (the per-step plan of the one-sided sharded step; lives in device memory)."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("begin", (C.c_int64 * (MAX_WORLD + 1)) * 2),
                ("tab", (C.c_void_p * 4) * MAX_WORLD), ("push_rows", (C.c_void_p * MAX_WORLD) * 2),
                ("push_ids", (C.c_void_p * MAX_WORLD) * 2)]


STATUS_WORDS = 4
STATUS_NAMES = ("user id outside [0, num_users)", "product id outside [0, num_products)", "hour outside [0, 24)")


class StatusWord:
    """The sticky error words of include/ncf_b200.h ("Id validation") in PINNED HOST memory: under unified addressing the
    kernels store into them through the same pointer, so the host reads them without a copy.  `raise_if_set()` is called
    where the host has synchronised anyway (after reading a loss / a result) and at the start of the next call."""

    def __init__(self):
        import torch
        self.t = torch.zeros(STATUS_WORDS, dtype=torch.int32).pin_memory()
        self.np = self.t.numpy()

    def ptr(self):
        return self.t.data_ptr()

    def raise_if_set(self, what=""):
        if self.np.any():
            bad = [STATUS_NAMES[k] for k in range(len(STATUS_NAMES)) if self.np[k]]
            self.np[:] = 0
            raise IndexError(f"{what or 'libncf_b200'}: " + "; ".join(bad) +
                             " (nn.EmbeddingBag / nn.Embedding raise on such an index; the kernels clamped it)")


class RunCfg(C.Structure):
    _fields_ = [("S", C.c_int32), ("training", C.c_int32), ("precision", C.c_int32), ("dropout_p", C.c_float),
                ("seed", C.c_uint64), ("step", C.c_uint64)]


class AdamCfg(C.Structure):
    _fields_ = [("lr", C.c_float), ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
                ("weight_decay", C.c_float), ("step", C.c_int32), ("emb_mode", C.c_int32)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_SIGS = {
    "ncf_version": (C.c_int, []),
    "ncf_last_error": (C.c_char_p, []),
    "ncf_launch_count": (_I64, []),
    "ncf_dense_numel": (_I64, []),
    "ncf_dense_offset": (_I64, [_I32]),
    "ncf_dense_size": (_I64, [_I32]),
    "ncf_workspace_bytes": (_I64, [_I64, C.POINTER(RunCfg)]),
    "ncf_forward": (C.c_int, [C.POINTER(RunCfg), C.POINTER(Tables), _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P]),
    "ncf_set_aux_stream": (C.c_int, [_P]),
    "ncf_set_sm_reserve": (C.c_int, [C.c_int32]),
    "ncf_set_loss_readback": (C.c_int, [_P, _P]),
    "ncf_check_ids": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _P]),
    "ncf_attn_fwd": (C.c_int, [C.POINTER(RunCfg), _P, _I64, _P, _I64, _P]),
    "ncf_mlp_fwd": (C.c_int, [C.POINTER(RunCfg), _P, _I64, _P, _P, _I64, _P]),
    "ncf_mlp_bwd": (C.c_int, [C.POINTER(RunCfg), _P, _P, _I64, _P, _P, _I64, _P]),
    "ncf_attn_bwd": (C.c_int, [C.POINTER(RunCfg), _P, _P, _I64, _P, _I64, _P]),
    "ncf_backward": (C.c_int, [C.POINTER(RunCfg), C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _P, _P, _I64, _P, _P,
                               _I64, _P]),
    "ncf_bce_loss": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "ncf_dense_adam": (C.c_int, [_P, _P, _P, _P, _I64, C.POINTER(AdamCfg), _P]),
    "ncf_train_step": (C.c_int, [C.POINTER(RunCfg), C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _P, _P, _P, _P, _P,
                                 _I64, _P, _P, _P, _I64, _P]),
    "ncf_gather_ln_gmf_fwd": (C.c_int, [C.POINTER(Tables), _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ncf_gather_ln_gmf_fwd_bf16": (C.c_int, [C.POINTER(Tables), _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ncf_gather_ln": (C.c_int, [C.POINTER(Tables), _P, _I32, _P, _I64, _P, _P, _P]),
    "ncf_emb_bwd_workspace_bytes": (_I64, [_I64]),
    "ncf_emb_bwd_adam": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _I32, _P, _P, _I64, _P, _P, _P, _P,
                                   _I64, _P]),
    "ncf_emb_bwd_adam_both": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P,
                                        _I64, _P]),
    "ncf_emb_bwd_adam_both_bf16": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _P,
                                        _I64, _P]),
    "ncf_emb_adam_sweep": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P]),
    "ncf_temporal_fwd": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _P, _P]),
    "ncf_temporal_tables": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "ncf_dropout_mask": (C.c_int, [C.POINTER(RunCfg), _I32, _I64, _P, _P]),
    "ncf_item_fold_workspace_bytes": (_I64, [_I64]),
    "ncf_item_fold": (C.c_int, [C.POINTER(Tables), _P, _P, _P, _P, _I64, _P]),
    "ncf_score_topk_workspace_bytes": (_I64, [_I64, _I64, _I32]),
    "ncf_score_topk": (C.c_int, [C.POINTER(Tables), _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _I64, _P]),
    "ncf_item_image_bytes": (_I64, [_I64]),
    "ncf_item_image": (C.c_int, [_P, _P, _I64, _P, _P]),
    "ncf_score_topk_tc_workspace_bytes": (_I64, [_I64, _I64, _I32]),
    "ncf_score_topk_tc": (C.c_int, [C.POINTER(Tables), _P, _P, _P, _P, _P, _I64, _I64, _I32, _P, _P, _P, _I64, _P]),
    "ncf_dot_topk_workspace_bytes": (_I64, [_I64, _I64, _I32, _I32]),
    "ncf_dot_topk": (C.c_int, [_P, _I64, _P, _P, _P, _I64, _I32, _P, _P, _P, _I64, _P]),
    "ncf_shard_bucketize_workspace_bytes": (_I64, [_I64, _I32]),
    "ncf_shard_bucketize": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _P, _I64, _P]),
    "ncf_shard_bucketize_runs_workspace_bytes": (_I64, [_I64, _I32]),
    "ncf_shard_bucketize_runs": (C.c_int, [_P, _I64, _I64, _I32, _P, _P, _P, _P, _I64, _P]),
    "ncf_shard_owner_rows": (C.c_int, [C.POINTER(Tables), _P, _I32, _P, _I64, _P, _P]),
    "ncf_shard_forward": (C.c_int, [C.POINTER(RunCfg), _P, _P, _P, _P, _P, _I64, _P, _P, _I64, _P]),
    "ncf_shard_backward": (C.c_int, [C.POINTER(RunCfg), _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P]),
    "ncf_ipc_export": (C.c_int, [_P, _P, C.POINTER(C.c_int64)]),
    "ncf_ipc_open": (C.c_int, [_P, _I64, C.POINTER(C.c_void_p)]),
    "ncf_ipc_close": (C.c_int, [_P]),
    "ncf_shard_pull_rows": (C.c_int, [_P, _P, _I32, _P, _I64, _P, _P]),
    "ncf_shard_backward_push": (C.c_int, [C.POINTER(RunCfg), _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _I64, _P]),
    "ncf_shard_route_workspace_bytes": (_I64, [_I64]),
    "ncf_shard_route": (C.c_int, [_P, _P, _I64, _I64, _I64, _I32, _P, _P, _P, _P, _I64, _P]),
    "ncf_sample_batch": (C.c_int, [_P, _P, _I64, _I32, _P, _I64, _P, _P, C.c_uint64, C.c_uint64, _P, _P, _P, _P]),
    "ncf_rank_metrics": (C.c_int, [_P, _P, _I64, _I32, C.POINTER(C.c_int32), _I32, _P, _P]),
    "ncf_auc_workspace_bytes": (_I64, [_I64, _I64]),
    "ncf_auc": (C.c_int, [_P, _P, _I64, _I64, C.c_float, _P, _P, _I64, _P]),
    "ncf_tc_selftest": (C.c_int, [_I32, _I32, _I32, _P, _P, _P, _P]),
    "ncf_shard_owner_update": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _I32, _P, _I64, _P, _P, _I64, _P]),
    "ncf_shard_owner_update_sorted": (C.c_int, [C.POINTER(AdamCfg), C.POINTER(Tables), _P, _P, _I32, _P, _I64, _P, _P, _I64, _P]),
    "ncf_shard_owner_sort": (C.c_int, [C.POINTER(Tables), _I32, _P, _I64, _P, _I64, _P]),
}
EXPORTS = tuple(_SIGS)

_lock = threading.Lock()
_lib = None


def load():
    """dlopen libncf_b200.so (built in-tree by build.py); raises NcfError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise NcfError(f"{LIB_PATH} is missing: run `python __graft_entry__.py build` (nvcc, sm_100a). "
                               "There is no CPU fallback for the AdvancedNCF hot path.")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGS.items():
                fn = getattr(lib, name)      # AttributeError here = header/library mismatch
                fn.restype = res
                fn.argtypes = args
            if lib.ncf_version() != 2:
                raise NcfError("libncf_b200.so ABI version mismatch")
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().ncf_last_error().decode("utf-8", "replace")
        raise NcfError(f"{what or 'libncf_b200'} failed ({rc}): {msg}")


def dense_layout():
    """[(state_dict key, offset, size)] of the flat dense buffer, and its padded length."""
    lib = load()
    return [(k, int(lib.ncf_dense_offset(i)), int(lib.ncf_dense_size(i))) for i, k in enumerate(DENSE_KEYS)], \
        int(lib.ncf_dense_numel())


def ptr(t):
    """device pointer of a torch tensor (or None)"""
    return None if t is None else C.c_void_p(t.data_ptr())
