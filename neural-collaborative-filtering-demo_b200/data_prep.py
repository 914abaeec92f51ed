"""Device-side input pipeline: what the reference does per sample in Python
(`SheetzDataset.__getitem__` + `_sample_negative`, data_prep.py:134-161, 181-212), per batch in
`collate_recommender_batch` (:230-320) and per epoch in `ConsistentBatchSampler` (:397-444), done for
a whole batch by one kernel (`ncf_sample_batch`) with the interactions resident on the GPU.

`InteractionSampler` iterates like the reference DataLoader: it yields `(KeyedJaggedTensor, targets[N,1])`
on the device, rows interaction-major (positive first), ids key-major.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterator, Optional, Tuple

import torch

from . import _lib
from .architecture import _stream
from .kjt import KeyedJaggedTensor


def first_appearance_index(values) -> dict:
    """`{v: idx for idx, v in enumerate(unique())}` of SheetzDataset.user_to_idx / product_to_idx
    (data_prep.py:65-71): dense ids in first-appearance order."""
    out = {}
    for v in values:
        if v not in out:
            out[v] = len(out)
    return out


def remap_product_id(pid: str, num_products: int) -> int:
    """inference-time id remap (local_inference.py:55,65; generate_embeddings.py:104)."""
    return int(pid.lstrip("P"), 16) % num_products


def remap_cardnumber(card, num_users: int) -> int:
    """local_inference.py:51."""
    return int(card) % num_users


class InteractionSampler:
    def __init__(self, user_idx: torch.Tensor, item_idx: torch.Tensor, num_users: int, num_products: int,
                 negative_samples: int = 4, batch_size: int = 256, shuffle: bool = True, seed: int = 42,
                 use_history: bool = True, device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.NcfError("InteractionSampler runs on CUDA only (no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.users = user_idx.to(self.device, torch.long).contiguous()
        self.items = item_idx.to(self.device, torch.long).contiguous()
        self.num_users, self.num_products = num_users, num_products
        self.S = 1 + negative_samples
        self.batch_size, self.shuffle = batch_size, shuffle
        self.seed, self.step = seed, 0
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        # inverse-popularity weights (data_prep.py:95-102): w = 1/max(count,1), normalised; CDF for the search
        counts = torch.bincount(self.items, minlength=num_products).clamp_min(1).double()
        w = 1.0 / counts
        self.weights = w / w.sum()
        self.cdf = torch.cumsum(self.weights, 0).contiguous()
        # user -> set of training items as a sorted CSR (data_prep.py:163-179)
        self.hist_off = self.hist_items = None
        if use_history:
            key = torch.unique(self.users * num_products + self.items)          # sorted, de-duplicated
            hu = torch.div(key, num_products, rounding_mode="floor")
            self.hist_items = (key % num_products).contiguous()
            self.hist_off = torch.zeros(num_users + 1, dtype=torch.long, device=self.device)
            self.hist_off[1:] = torch.cumsum(torch.bincount(hu, minlength=num_users), 0)

    def __len__(self) -> int:
        return (self.users.numel() + self.batch_size - 1) // self.batch_size

    def batch_indices(self) -> Iterator[torch.Tensor]:
        """ConsistentBatchSampler.__iter__ (data_prep.py:419-440): shuffled indices, last batch padded with its
        own first indices."""
        n = self.users.numel()
        idx = torch.randperm(n, device=self.device, generator=self.gen) if self.shuffle else torch.arange(n, device=self.device)
        for i in range(len(self)):
            b = idx[i * self.batch_size:(i + 1) * self.batch_size]
            if b.numel() < self.batch_size:
                b = torch.cat([b, b[:self.batch_size - b.numel()]])
            yield b

    def sample(self, pos_user: torch.Tensor, pos_item: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """(user_ids [B*S], item_ids [B*S], targets [B*S]) for B positive interactions."""
        B = pos_user.numel()
        n = B * self.S
        u = torch.empty(n, dtype=torch.long, device=self.device)
        it = torch.empty(n, dtype=torch.long, device=self.device)
        t = torch.empty(n, dtype=torch.float32, device=self.device)
        self.step += 1
        _lib.check(self.lib.ncf_sample_batch(_lib.ptr(pos_user.contiguous()), _lib.ptr(pos_item.contiguous()), B, self.S,
                                             _lib.ptr(self.cdf), self.num_products, _lib.ptr(self.hist_off),
                                             _lib.ptr(self.hist_items), self.seed, self.step, _lib.ptr(u), _lib.ptr(it),
                                             _lib.ptr(t), _stream(self.device)), "ncf_sample_batch")
        return u, it, t

    def __iter__(self):
        for b in self.batch_indices():
            u, it, t = self.sample(self.users[b], self.items[b])
            values = torch.cat([u, it])
            kjt = KeyedJaggedTensor.from_lengths_sync(keys=["user_id", "product_id"], values=values,
                                                      lengths=torch.ones(values.numel(), dtype=torch.long, device=self.device))
            yield kjt, t.view(-1, 1)
