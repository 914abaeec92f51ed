"""KeyedJaggedTensor: the input container of AdvancedNCF.forward.

The reference takes `torchrec.sparse.jagged_tensor.KeyedJaggedTensor` (torchrec==0.8.0, reference
Dockerfile:22-27).  When torchrec is importable that class is re-exported unchanged; otherwise this
minimal class provides the surface the reference's callers use: `keys()`, `values()`,
`lengths()`, `to(device)`, `from_lengths_sync`, and the `offsets=` constructor of
generate_embeddings.py:107-112.  Layout: `values` is key-major (`[users..., products...]`,
data_prep.py:286-298) with one id per (key, sample), i.e. `lengths` is all ones.
"""
from __future__ import annotations

from typing import List, Optional

import torch

try:  # pragma: no cover - torchrec is not part of this image
    from torchrec.sparse.jagged_tensor import KeyedJaggedTensor as _TorchrecKJT
except Exception:  # noqa: BLE001
    _TorchrecKJT = None


class _KeyedJaggedTensor:
    def __init__(self, keys: List[str], values: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, weights: Optional[torch.Tensor] = None):
        if lengths is None:
            if offsets is None:
                raise ValueError("KeyedJaggedTensor needs lengths or offsets")
            lengths = offsets[1:] - offsets[:-1]
        self._keys = list(keys)
        self._values = values
        self._lengths = lengths

    @staticmethod
    def from_lengths_sync(keys, values, lengths, weights=None):
        return _KeyedJaggedTensor(keys, values, lengths=lengths)

    @staticmethod
    def from_offsets_sync(keys, values, offsets, weights=None):
        return _KeyedJaggedTensor(keys, values, offsets=offsets)

    def keys(self):
        return self._keys

    def values(self):
        return self._values

    def lengths(self):
        return self._lengths

    def stride(self):
        return self._lengths.numel() // max(1, len(self._keys))

    def to(self, device, non_blocking: bool = False):
        return _KeyedJaggedTensor(self._keys, self._values.to(device, non_blocking=non_blocking),
                                  lengths=self._lengths.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return _KeyedJaggedTensor(self._keys, self._values.pin_memory(), lengths=self._lengths.pin_memory())

    def __repr__(self):
        return f"KeyedJaggedTensor(keys={self._keys}, values={tuple(self._values.shape)})"


KeyedJaggedTensor = _TorchrecKJT if _TorchrecKJT is not None else _KeyedJaggedTensor


def make_kjt(user_ids: torch.Tensor, product_ids: torch.Tensor) -> "KeyedJaggedTensor":
    """key-major KJT of one id per sample, as collate_recommender_batch builds it (data_prep.py:286-298)."""
    values = torch.cat([user_ids.reshape(-1), product_ids.reshape(-1)]).long()
    return KeyedJaggedTensor.from_lengths_sync(keys=["user_id", "product_id"], values=values,
                                               lengths=torch.ones(values.numel(), dtype=torch.long, device=values.device))
