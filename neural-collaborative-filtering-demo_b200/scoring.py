"""Full-catalogue scoring + top-k retrieval (reference app.py:43-77 `get_recommendations`).

`CatalogueScorer` folds the item side once (`ncf_item_fold`: eval-mode attention sees one key, so
logit(u,i) = LN_mf(U_mf[u]).P_hat[i] + g[i]) and then ranks users against the whole catalogue with
`ncf_score_topk`; order = score descending, ties -> lowest item index (pandas nlargest keep='first').

Large catalogues with many users per call go through `ncf_score_topk_tc`: a tcgen05 bf16 GEMM bounds every logit
from above and only the pairs that could enter a user's list are re-scored exactly, so the output is bit-identical
to `ncf_score_topk` (tests/test_gpu_parity.py) at a fraction of the CUDA-core work.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch

from . import _lib
from .architecture import AdvancedNCF, _stream


class CatalogueScorer:
    def __init__(self, model: AdvancedNCF, tc_min_items: Optional[int] = None, tc_min_users: Optional[int] = None):
        self.model = model
        self.lib = _lib.load()
        self.p_hat = None
        self.g = None
        self.img = None            # item tile images of the tensor-core pre-filter (large catalogues)
        self.use_tc = os.environ.get("NCF_SCORE_TC", "1") != "0"
        if tc_min_items is not None:
            self.TC_MIN_ITEMS = tc_min_items
        if tc_min_users is not None:
            self.TC_MIN_USERS = tc_min_users
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        """Recompute the folded item table (call after the weights change)."""
        m = self.model
        m._ensure_flat()
        dev = m._flat.device
        I = m.num_products
        self.p_hat = torch.empty(I, 64, device=dev)
        self.g = torch.empty(I, device=dev)
        nbytes = int(self.lib.ncf_item_fold_workspace_bytes(I))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        tables = m._tables_struct()
        _lib.check(self.lib.ncf_item_fold(C.byref(tables), _lib.ptr(m._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                          _lib.ptr(ws), nbytes, _stream(dev)), "ncf_item_fold")
        self.img = None
        if self.use_tc and I >= self.TC_MIN_ITEMS:
            self.img = torch.empty(int(self.lib.ncf_item_image_bytes(I)), dtype=torch.uint8, device=dev)
            _lib.check(self.lib.ncf_item_image(_lib.ptr(self.p_hat), _lib.ptr(self.g), I, _lib.ptr(self.img), _stream(dev)),
                       "ncf_item_image")

    TC_MIN_ITEMS = 1 << 16      # below this the per-CTA warm-up of the pre-filter does not pay
    TC_MIN_USERS = 64

    @torch.no_grad()
    def topk(self, user_ids: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(indices int64 [n,k], scores fp32 [n,k]) over the whole catalogue."""
        m = self.model
        dev = m._flat.device
        u = user_ids.reshape(-1).to(device=dev, dtype=torch.long).contiguous()
        n, I = u.numel(), m.num_products
        k_eff = min(k, I)
        idx = torch.empty(n, k_eff, dtype=torch.long, device=dev)
        sc = torch.empty(n, k_eff, dtype=torch.float32, device=dev)
        if n == 0:
            return idx, sc
        tables = m._tables_struct()
        if self.img is not None and n >= self.TC_MIN_USERS:
            nbytes = int(self.lib.ncf_score_topk_tc_workspace_bytes(n, I, k_eff))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.check(self.lib.ncf_score_topk_tc(C.byref(tables), _lib.ptr(m._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                                  _lib.ptr(self.img), _lib.ptr(u), n, I, k_eff, _lib.ptr(idx), _lib.ptr(sc),
                                                  _lib.ptr(ws), nbytes, _stream(dev)), "ncf_score_topk_tc")
            return idx, sc
        nbytes = int(self.lib.ncf_score_topk_workspace_bytes(n, I, k_eff))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(self.lib.ncf_score_topk(C.byref(tables), _lib.ptr(m._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                           _lib.ptr(u), n, I, k_eff, _lib.ptr(idx), _lib.ptr(sc), _lib.ptr(ws), nbytes,
                                           _stream(dev)), "ncf_score_topk")
        return idx, sc


def get_recommendations(model: AdvancedNCF, customer_id: int, num_products: int, top_k: int, selected_hour=None):
    """reference app.py:43-77 without the DataFrame: item row indices and scores of the top_k rows,
    rows = arange(num_products) % model.num_products (:48), ordered like DataFrame.nlargest (:75)."""
    dev = next(model.parameters()).device
    all_products = torch.arange(num_products, device=dev) % model.num_products
    customer = torch.full((num_products,), int(customer_id), device=dev, dtype=torch.long)
    hour = None
    if selected_hour is not None:
        hour = torch.full((num_products,), int(selected_hour), device=dev, dtype=torch.long)
    scores = model.forward_simple(customer, all_products, hour)
    order = torch.argsort(scores, descending=True, stable=True)[:top_k]
    return order, scores[order]
