"""Embedding export + exact nearest-neighbour hand-off.

`export_product_embeddings` mirrors the reference's `generate_embeddings` loop (src/inference/generate_embeddings.py:
184-236) without the GCS plumbing: product id -> row by the reference's hash-mod (`:104`), `get_product_embeddings`
on the device in batches instead of one product per call, the `mlp` vector L2-normalised (`:213-216`), one JSON
object `{"id", "embedding"}` per line (`:218-222`), first occurrence of a product id wins (`:199-200`).

`CosineIndex` is an EXACT cosine top-k over those vectors: the stand-in for the Vertex Tree-AH index the reference
configures (setup_tree_ah_endpoint.py:25-32: 64 dimensions, cosine distance) - useful as ground truth for an ANN
index and as the retrieval step of the serving design (src/api/routes.py:55-70)."""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch

from .architecture import AdvancedNCF
from .data_prep import remap_product_id
from .kjt import KeyedJaggedTensor


def _category_ids(model: AdvancedNCF, values: Sequence, mapping: Optional[Dict], upper: int) -> torch.Tensor:
    """generate_embeddings.py:96-101: map through the (sorted-name) dictionary, default 0, clamp into range."""
    out = []
    for v in values:
        i = mapping.get(v, 0) if mapping is not None else int(v)
        out.append(min(max(0, int(i)), upper - 1))
    return torch.tensor(out, dtype=torch.long)


@torch.no_grad()
def product_embedding_records(model: AdvancedNCF, product_ids: Sequence[str], category_ids: Optional[Sequence] = None,
                              department_ids: Optional[Sequence] = None, category_map: Optional[Dict] = None,
                              department_map: Optional[Dict] = None, batch: int = 65536,
                              which: str = "mlp") -> Tuple[List[str], torch.Tensor]:
    """(ids, L2-normalised vectors [n,64] on the model's device) for the first occurrence of every product id."""
    dev = next(model.parameters()).device
    seen, keep = set(), []
    for k, pid in enumerate(product_ids):
        if not pid or pid in seen:
            continue
        seen.add(pid)
        keep.append(k)
    ids = [str(product_ids[k]) for k in keep]
    rows = torch.tensor([remap_product_id(product_ids[k], model.num_products) for k in keep], dtype=torch.long)
    cats = _category_ids(model, [category_ids[k] for k in keep] if category_ids is not None else [0] * len(keep),
                         category_map, model.num_categories)
    deps = _category_ids(model, [department_ids[k] for k in keep] if department_ids is not None else [0] * len(keep),
                         department_map, model.num_departments)
    was_training = model.training
    model.eval()
    out = torch.empty(len(keep), 64, device=dev)
    for s in range(0, len(keep), batch):
        r = rows[s:s + batch].to(dev)
        kjt = KeyedJaggedTensor(keys=["user_id", "product_id"], values=torch.cat([torch.zeros_like(r), r]),
                                lengths=torch.ones(2 * r.numel(), dtype=torch.long, device=dev))
        emb = model.get_product_embeddings({"product_features": kjt,
                                            "category_features": {"department_ids": deps[s:s + batch].to(dev),
                                                                  "category_ids": cats[s:s + batch].to(dev)}})
        v = emb[which].reshape(r.numel(), -1)[:, :64]
        out[s:s + batch] = v / v.norm(dim=1, keepdim=True)
    model.train(was_training)
    return ids, out


def export_product_embeddings(model: AdvancedNCF, product_ids: Sequence[str], path: str, **kw) -> int:
    """Write the JSONL file the reference uploads to `embeddings/<name>` (generate_embeddings.py:218-236)."""
    ids, vecs = product_embedding_records(model, product_ids, **kw)
    host = vecs.cpu().tolist()
    with open(path, "w") as f:
        for pid, v in zip(ids, host):
            f.write(json.dumps({"id": pid, "embedding": v}) + "\n")
    return len(ids)


class CosineIndex:
    """Exact cosine top-k over L2-normalised vectors (dot product = cosine)."""

    def __init__(self, ids: Sequence[str], vectors: torch.Tensor):
        self.ids = list(ids)
        self.vectors = vectors / vectors.norm(dim=1, keepdim=True).clamp_min(1e-30)

    @classmethod
    def from_jsonl(cls, path: str, device="cuda") -> "CosineIndex":
        ids, vecs = [], []
        with open(path) as f:
            for line in f:
                rec = json.loads(line)
                ids.append(rec["id"])
                vecs.append(rec["embedding"])
        return cls(ids, torch.tensor(vecs, dtype=torch.float32, device=device))

    @torch.no_grad()
    def query(self, queries: torch.Tensor, k: int, chunk: int = 1 << 20) -> Tuple[torch.Tensor, torch.Tensor]:
        """(neighbour positions int64 [n,k], cosine similarities [n,k]); ties -> lowest position.  The catalogue is
        streamed in chunks with a running top-k merge, so it never needs an [n, catalogue] score matrix."""
        q = queries.to(self.vectors.device, torch.float32)
        q = q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)
        n, total = q.shape[0], self.vectors.shape[0]
        k = min(k, total)
        best_s = torch.full((n, 0), 0.0, device=q.device)
        best_i = torch.zeros((n, 0), dtype=torch.long, device=q.device)
        for s in range(0, total, chunk):
            sc = q @ self.vectors[s:s + chunk].t()
            idx = torch.arange(s, s + sc.shape[1], device=q.device).expand(n, -1)
            cs, ci = torch.cat([best_s, sc], 1), torch.cat([best_i, idx], 1)
            order = torch.argsort(cs, dim=1, descending=True, stable=True)[:, :k]     # stable: earlier position wins ties
            best_s, best_i = torch.gather(cs, 1, order), torch.gather(ci, 1, order)
        return best_i, best_s

    def lookup(self, positions: torch.Tensor) -> List[List[str]]:
        return [[self.ids[j] for j in row] for row in positions.cpu().tolist()]
