"""Minimal stand-in for the `torchrec` names the reference imports.

TEST INFRASTRUCTURE ONLY.  torchrec==0.8.0 (pinned at reference Dockerfile:22-27) is not
installed in this image; this shim restates the public behaviour the reference relies on so
that /root/reference/src/model/{architecture,trainer,data_prep}.py import *unchanged* when
oracle/make_golden.py generates fixtures.  Nothing in the product package imports this.

Restated semantics (torchrec public behaviour):
  * EmbeddingBagCollection: one nn.EmbeddingBag(mode="sum", include_last_offset=True) per
    table, registered under `embedding_bags.<table name>`; weights ~ U(-sqrt(1/n), sqrt(1/n)).
  * KeyedJaggedTensor: key-major `values`, per-(key,sample) `lengths`; stride = len(lengths)/len(keys).
"""
import enum
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .sparse.jagged_tensor import KeyedJaggedTensor  # noqa: F401


class PoolingType(enum.Enum):
    SUM = "SUM"
    MEAN = "MEAN"
    NONE = "NONE"


@dataclass
class EmbeddingBagConfig:
    name: str = ""
    embedding_dim: int = 0
    num_embeddings: int = 0
    feature_names: List[str] = field(default_factory=list)
    pooling: PoolingType = PoolingType.SUM


class _KeyedResult:
    """Mapping feature name -> [stride, dim] (what `KeyedTensor` offers the reference)."""

    def __init__(self, d: Dict[str, torch.Tensor]):
        self._d = d

    def __getitem__(self, k: str) -> torch.Tensor:
        return self._d[k]

    def keys(self):
        return list(self._d.keys())

    def to_dict(self):
        return dict(self._d)


class EmbeddingBagCollection(nn.Module):
    def __init__(self, tables: List[EmbeddingBagConfig], device: Optional[torch.device] = None):
        super().__init__()
        self.embedding_bags = nn.ModuleDict()
        self._feature_to_table: Dict[str, str] = {}
        for cfg in tables:
            bag = nn.EmbeddingBag(cfg.num_embeddings, cfg.embedding_dim, mode="sum",
                                  include_last_offset=True, device=device)
            bound = math.sqrt(1.0 / cfg.num_embeddings)
            with torch.no_grad():
                bag.weight.uniform_(-bound, bound)
            self.embedding_bags[cfg.name] = bag
            for f in cfg.feature_names:
                self._feature_to_table[f] = cfg.name

    def forward(self, features: KeyedJaggedTensor) -> _KeyedResult:
        out = {}
        for f, t in self._feature_to_table.items():
            jt = features[f]
            out[f] = self.embedding_bags[t](jt.values(), jt.offsets())
        return _KeyedResult(out)
