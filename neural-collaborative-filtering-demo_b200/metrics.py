"""Ranking metrics of the reference (`src/utils/metrics.py`) on the device that holds the scores:
HR@K, NDCG@K, MRR@K, MAP@K over [batch, 1+neg] groups (:110-242), AUC (:244-265, sklearn.roc_auc_score) and accuracy
(:267-275).  Same keys and argument meaning as `calculate_metrics` (:9-108).

CUDA tensors go through two kernels of libncf_b200 (`ncf_rank_metrics`: the rank of every positive inside its group, one
pass, nothing sorted; `ncf_auc`: exact Mann-Whitney count in 64-bit integers).  Host tensors (the reference's own
`validate` gathers everything to the CPU, trainer.py:387-400) use the vectorised torch code below - this is the
reference's host function, not part of the GPU hot path."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import _lib


def _run_auc(P: torch.Tensor, T: torch.Tensor, out4: torch.Tensor, cap: int) -> None:
    lib = _lib.load()
    n = P.numel()
    nbytes = int(lib.ncf_auc_workspace_bytes(n, cap))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=P.device)
    _lib.check(lib.ncf_auc(_lib.ptr(P), _lib.ptr(T), n, cap, 0.5, _lib.ptr(out4), _lib.ptr(ws), nbytes,
                           C.c_void_p(torch.cuda.current_stream(P.device).cuda_stream)), "ncf_auc")


def _device_metrics(P: torch.Tensor, T: torch.Tensor, k_values: List[int], small_class_cap: int = 0) -> Dict[str, float]:
    """all metrics for CUDA score / target matrices [groups, M] with one D2H read of 4 * n_k + 4 doubles"""
    lib = _lib.load()
    dev = P.device
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    P = P.contiguous().float()
    T = T.contiguous().float()
    groups, M = P.shape
    out: Dict[str, float] = {}
    res = torch.zeros(4 * len(k_values) + 4, dtype=torch.float64, device=dev)
    for s in range(0, len(k_values), 8):
        ks = k_values[s:s + 8]
        kv = (C.c_int32 * len(ks))(*ks)
        _lib.check(lib.ncf_rank_metrics(_lib.ptr(P), _lib.ptr(T), groups, M, kv, len(ks),
                                        C.c_void_p(res.data_ptr() + 32 * s), stream), "ncf_rank_metrics")
    auc_out = res[4 * len(k_values):]
    _run_auc(P, T, auc_out, small_class_cap)
    host = res.cpu()
    if small_class_cap and float(host[4 * len(k_values)]) == -1.0:      # the capacity hint was too small: full-size network
        _run_auc(P, T, auc_out, 0)
        host = res.cpu()
    for q, k in enumerate(k_values):
        h = host[4 * q:4 * q + 4] / max(groups, 1)
        out[f"hit_rate@{k}"], out[f"ndcg@{k}"], out[f"mrr@{k}"], out[f"map@{k}"] = (float(x) for x in h)
    out["auc"] = float(host[4 * len(k_values)])
    out["accuracy"] = float(host[4 * len(k_values) + 1])
    return out


def calculate_auc(preds: torch.Tensor, targets: torch.Tensor) -> float:
    if preds.is_cuda:
        out4 = torch.zeros(4, dtype=torch.float64, device=preds.device)
        _run_auc(preds.reshape(-1).contiguous().float(), targets.reshape(-1).to(preds.device).contiguous().float(), out4, 0)
        return float(out4[0])
    preds = preds.reshape(-1).double()
    pos = targets.reshape(-1) > 0.5
    n_pos = int(pos.sum())
    n_neg = preds.numel() - n_pos
    if n_pos == 0 or n_neg == 0:
        return float("nan")          # sklearn: undefined with a single class (reference validate() hits this)
    order = torch.argsort(preds, stable=True)
    sp = preds[order]
    _, inv, cnt = torch.unique_consecutive(sp, return_inverse=True, return_counts=True)
    end = torch.cumsum(cnt, 0).double()
    avg_rank = ((end - cnt.double() + 1.0 + end) / 2.0)[inv]
    ranks = torch.empty_like(avg_rank)
    ranks[order] = avg_rank
    return float((ranks[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


def calculate_accuracy(preds: torch.Tensor, targets: torch.Tensor, threshold: float = 0.5) -> float:
    return float(((preds >= threshold) == (targets > 0.5)).float().mean())


def calculate_metrics(predictions: torch.Tensor, targets: torch.Tensor, k_values: List[int] = [1, 5, 10],
                      batch_size: Optional[int] = None, negative_samples: Optional[int] = None) -> Dict[str, float]:
    predictions = predictions.detach()
    targets = targets.detach()
    if predictions.dim() == 2 and predictions.size(1) == 1:
        predictions = predictions.squeeze(1)
    if targets.dim() == 2 and targets.size(1) == 1:
        targets = targets.squeeze(1)
    if batch_size is None or negative_samples is None:
        raise ValueError("Please provide both batch_size and negative_samples "
                         "to reshape predictions into [batch_size, 1+negative_samples].")
    M = 1 + negative_samples
    if predictions.numel() != batch_size * M:
        raise ValueError(f"Size mismatch: got {predictions.numel()} total preds, "
                         f"but expected batch_size*M = {batch_size * M}.")
    P = predictions.reshape(batch_size, M).float()
    T = targets.reshape(batch_size, M).float()
    dev = P.device
    if P.is_cuda:
        out = _device_metrics(P, T.to(dev), list(k_values), small_class_cap=2 * batch_size)
        flat_p, flat_t = P.reshape(-1), T.reshape(-1).to(dev)
        pm, nm = flat_t == 1, flat_t == 0
        if bool(pm.any()):
            out["pos_accuracy"] = calculate_accuracy(flat_p[pm], flat_t[pm])
        if bool(nm.any()):
            out["neg_accuracy"] = calculate_accuracy(flat_p[nm], flat_t[nm])
        return out
    out: Dict[str, float] = {}
    order = torch.sort(P, dim=1, descending=True, stable=True).indices      # ties: lower index first, as the kernel
    rel_sorted = torch.gather(T, 1, order)
    ideal = torch.sort(T, dim=1, descending=True).values
    for k in k_values:
        kk = min(k, M)
        rel = rel_sorted[:, :kk]
        pos = torch.arange(1, kk + 1, device=dev, dtype=torch.float32)
        disc = 1.0 / torch.log2(pos + 1.0)
        out[f"hit_rate@{k}"] = float((rel.sum(1) > 0).float().mean())
        dcg = (rel * disc).sum(1)
        idcg = (ideal[:, :kk] * disc).sum(1)
        out[f"ndcg@{k}"] = float(torch.where(idcg > 0, dcg / idcg.clamp_min(1e-30), torch.zeros_like(dcg)).mean())
        is_pos = rel == 1
        first = torch.where(is_pos, 1.0 / pos, torch.zeros((), device=dev)).max(dim=1).values
        out[f"mrr@{k}"] = float(first.mean())
        cum = torch.cumsum(is_pos.float(), dim=1)
        prec = torch.where(is_pos, cum / pos, torch.zeros((), device=dev)).sum(1)
        nrel = is_pos.float().sum(1)
        out[f"map@{k}"] = float(torch.where(nrel > 0, prec / nrel.clamp_min(1.0), torch.zeros_like(prec)).mean())
    flat_p, flat_t = P.reshape(-1), T.reshape(-1)
    out["auc"] = calculate_auc(flat_p, flat_t)
    out["accuracy"] = calculate_accuracy(flat_p, flat_t)
    pm, nm = flat_t == 1, flat_t == 0
    if bool(pm.any()):
        out["pos_accuracy"] = calculate_accuracy(flat_p[pm], flat_t[pm])
    if bool(nm.any()):
        out["neg_accuracy"] = calculate_accuracy(flat_p[nm], flat_t[nm])
    return out
