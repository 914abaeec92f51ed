"""KeyedJaggedTensor / JaggedTensor stand-ins (TEST INFRASTRUCTURE ONLY, see torchrec/__init__.py)."""
from typing import List, Optional

import torch


class JaggedTensor:
    def __init__(self, values: torch.Tensor, lengths: torch.Tensor):
        self._values = values
        self._lengths = lengths

    def values(self) -> torch.Tensor:
        return self._values

    def lengths(self) -> torch.Tensor:
        return self._lengths

    def offsets(self) -> torch.Tensor:
        z = torch.zeros(1, dtype=self._lengths.dtype, device=self._lengths.device)
        return torch.cat([z, torch.cumsum(self._lengths, 0)])


class KeyedJaggedTensor:
    def __init__(self, keys: List[str], values: torch.Tensor,
                 lengths: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                 weights: Optional[torch.Tensor] = None):
        if lengths is None:
            if offsets is None:
                raise ValueError("need lengths or offsets")
            lengths = offsets[1:] - offsets[:-1]
        self._keys = list(keys)
        self._values = values
        self._lengths = lengths

    @staticmethod
    def from_lengths_sync(keys, values, lengths, weights=None):
        return KeyedJaggedTensor(keys=keys, values=values, lengths=lengths)

    @staticmethod
    def from_offsets_sync(keys, values, offsets, weights=None):
        return KeyedJaggedTensor(keys=keys, values=values, offsets=offsets)

    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def lengths(self) -> torch.Tensor:
        return self._lengths

    def stride(self) -> int:
        return self._lengths.numel() // len(self._keys)

    def to(self, device, non_blocking: bool = False):
        return KeyedJaggedTensor(self._keys, self._values.to(device, non_blocking=non_blocking),
                                 self._lengths.to(device, non_blocking=non_blocking))

    def pin_memory(self):
        return KeyedJaggedTensor(self._keys, self._values.pin_memory(), self._lengths.pin_memory())

    def __getitem__(self, key: str) -> JaggedTensor:
        i = self._keys.index(key)
        s = self.stride()
        lens = self._lengths[i * s:(i + 1) * s]
        start = int(self._lengths[: i * s].sum())
        n = int(lens.sum())
        return JaggedTensor(self._values[start:start + n], lens)

    def __repr__(self):
        return f"KeyedJaggedTensor(keys={self._keys}, values={self._values.shape}, lengths={self._lengths.shape})"
