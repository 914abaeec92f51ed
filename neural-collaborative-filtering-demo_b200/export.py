"""Embedding export + exact nearest-neighbour hand-off.

`export_product_embeddings` mirrors the reference's `generate_embeddings` loop (src/inference/generate_embeddings.py:
184-236) without the GCS plumbing: product id -> row by the reference's hash-mod (`:104`), `get_product_embeddings`
on the device in batches instead of one product per call, the `mlp` vector L2-normalised (`:213-216`), one JSON
object `{"id", "embedding"}` per line (`:218-222`), first occurrence of a product id wins (`:199-200`).

`CosineIndex` is the stand-in for the Vertex Tree-AH index the reference configures (setup_tree_ah_endpoint.py:25-32: 64
dimensions, cosine distance): an EXACT cosine top-k running on the library's own scoring kernels (`ncf_dot_topk`), plus an
inverted-file (IVF) approximate mode checked for recall against it - the retrieval step of the serving design
(src/api/routes.py:55-70)."""
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


def _dot_topk(queries: torch.Tensor, vectors: torch.Tensor, bias: torch.Tensor, k: int, image: Optional[torch.Tensor] = None):
    """ncf_dot_topk: (positions int64 [n,k], sigmoid(dot + bias) [n,k]) - the scoring kernels of the catalogue scorer on raw rows."""
    from . import _lib
    import ctypes as C
    lib = _lib.load()
    dev = vectors.device
    q = queries.to(dev, torch.float32).contiguous()
    n, I = q.shape[0], vectors.shape[0]
    k = min(k, I)
    idx = torch.empty(n, k, dtype=torch.long, device=dev)
    sc = torch.empty(n, k, dtype=torch.float32, device=dev)
    if n == 0:
        return idx, sc
    nbytes = int(lib.ncf_dot_topk_workspace_bytes(n, I, k, 1 if image is not None else 0))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.ncf_dot_topk(_lib.ptr(q), n, _lib.ptr(vectors), _lib.ptr(bias), _lib.ptr(image), I, k, _lib.ptr(idx), _lib.ptr(sc),
                                _lib.ptr(ws), nbytes, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "ncf_dot_topk")
    return idx, sc


class CosineIndex:
    """Cosine top-k over L2-normalised 64-d vectors (dot product = cosine): the stand-in for the Vertex Tree-AH index the
    reference configures (setup_tree_ah_endpoint.py:25-32: 64 dimensions, cosine distance).

    `query` is EXACT: it runs the catalogue-scoring kernels of libncf_b200 on raw rows (`ncf_dot_topk`: fp32 dot products,
    running top-k with (similarity descending, position ascending) order, tensor-core pre-filter for large indexes) - no
    cuBLAS, no [n, catalogue] score matrix, no argsort.
    `build_ivf` + `query_ivf` are the approximate option (inverted file: spherical k-means lists, probe the nearest
    `nprobe` lists), checked for recall against `query`."""

    TC_MIN_VECTORS = 1 << 16

    def __init__(self, ids: Sequence[str], vectors: torch.Tensor):
        self.ids = list(ids)
        v = vectors.to(torch.float32)
        self.vectors = (v / v.norm(dim=1, keepdim=True).clamp_min(1e-30)).contiguous()
        if self.vectors.shape[1] != 64:
            raise NotImplementedError("the retrieval kernels are built for 64-d vectors (the reference's embedding_dim)")
        self._zero_bias = torch.zeros(self.vectors.shape[0], device=self.vectors.device)
        self._image = None
        self._ivf = None

    @classmethod
    def from_jsonl(cls, path: str, device="cuda") -> "CosineIndex":
        ids, vecs = [], []
        with open(path) as f:
            for line in f:
                rec = json.loads(line)
                ids.append(rec["id"])
                vecs.append(rec["embedding"])
        return cls(ids, torch.tensor(vecs, dtype=torch.float32, device=device))

    def _cosines(self, q: torch.Tensor, positions: torch.Tensor) -> torch.Tensor:
        """the similarities of the k winners as plain dot products (the kernels rank by the monotone sigmoid(dot))"""
        return (q.unsqueeze(1) * self.vectors[positions]).sum(-1)

    @torch.no_grad()
    def query(self, queries: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """(neighbour positions int64 [n,k], cosine similarities [n,k]); ties -> lowest position."""
        from . import _lib
        import ctypes as C
        q = queries.to(self.vectors.device, torch.float32)
        q = (q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)).contiguous()
        if not q.is_cuda:
            raise _lib.NcfError("CosineIndex.query runs on CUDA only (no CPU fallback)")
        total = self.vectors.shape[0]
        if self._image is None and total >= self.TC_MIN_VECTORS:
            lib = _lib.load()
            self._image = torch.empty(int(lib.ncf_item_image_bytes(total)), dtype=torch.uint8, device=q.device)
            _lib.check(lib.ncf_item_image(_lib.ptr(self.vectors), _lib.ptr(self._zero_bias), total, _lib.ptr(self._image),
                                          C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)), "ncf_item_image")
        idx, _ = _dot_topk(q, self.vectors, self._zero_bias, k, self._image if q.shape[0] >= 64 else None)
        return idx, self._cosines(q, idx)

    # ---- inverted-file option ---------------------------------------------------------------------------------------
    @torch.no_grad()
    def build_ivf(self, nlist: int = 0, iters: int = 8, seed: int = 0) -> "CosineIndex":
        """Spherical k-means with `nlist` centroids (default ~ sqrt(n)); every Lloyd assignment is a k = 1 call of the same
        top-k kernel (vectors as queries, centroids as the catalogue); vectors are then stored list by list."""
        V = self.vectors
        n = V.shape[0]
        nlist = nlist or max(1, int(n ** 0.5))
        g = torch.Generator(device=V.device).manual_seed(seed)
        cent = V[torch.randperm(n, device=V.device, generator=g)[:nlist]].clone()
        zero = torch.zeros(nlist, device=V.device)
        for _ in range(iters):
            a = _dot_topk(V, cent, zero, 1)[0][:, 0]
            new = torch.zeros_like(cent).index_add_(0, a, V)
            empty = new.norm(dim=1) == 0
            new[empty] = cent[empty]
            cent = new / new.norm(dim=1, keepdim=True).clamp_min(1e-30)
        a = _dot_topk(V, cent, zero, 1)[0][:, 0]
        order = torch.argsort(a, stable=True)
        counts = torch.bincount(a, minlength=nlist)
        offsets = torch.zeros(nlist + 1, dtype=torch.long, device=V.device)
        offsets[1:] = torch.cumsum(counts, 0)
        self._ivf = {"centroids": cent.contiguous(), "zero": zero, "order": order, "vectors": V[order].contiguous(),
                     "offsets": offsets.cpu().tolist()}
        return self

    @torch.no_grad()
    def query_ivf(self, queries: torch.Tensor, k: int, nprobe: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
        """approximate (positions, cosines): only the `nprobe` lists whose centroids are closest to each query are scanned,
        list by list for all the queries that probe it (exact kernel on the list's slice), then the per-list winners
        are merged.  Missing neighbours are filled with position -1 / similarity -inf."""
        if self._ivf is None:
            self.build_ivf()
        f = self._ivf
        q = queries.to(self.vectors.device, torch.float32)
        q = (q / q.norm(dim=1, keepdim=True).clamp_min(1e-30)).contiguous()
        n, nlist = q.shape[0], f["centroids"].shape[0]
        nprobe = min(nprobe, nlist)
        probes = _dot_topk(q, f["centroids"], f["zero"], nprobe)[0]                       # [n, nprobe]
        cand_s = torch.full((n, nprobe * k), float("-inf"), device=q.device)
        cand_i = torch.full((n, nprobe * k), -1, dtype=torch.long, device=q.device)
        for l in torch.unique(probes).tolist():
            lo, hi = f["offsets"][l], f["offsets"][l + 1]
            if hi == lo:
                continue
            who, slot = (probes == l).nonzero(as_tuple=True)
            seg = f["vectors"][lo:hi]
            idx, _ = _dot_topk(q[who], seg, f["zero"][:1].expand(hi - lo).contiguous(), k)
            kk = idx.shape[1]
            cols = slot.unsqueeze(1) * k + torch.arange(kk, device=q.device)
            cand_s[who.unsqueeze(1), cols] = (q[who].unsqueeze(1) * seg[idx]).sum(-1)
            cand_i[who.unsqueeze(1), cols] = f["order"][lo + idx]
        top = torch.sort(cand_s, dim=1, descending=True, stable=True)
        keep = top.indices[:, :k]
        return torch.gather(cand_i, 1, keep), top.values[:, :k]

    def lookup(self, positions: torch.Tensor) -> List[List[str]]:
        return [[self.ids[j] if j >= 0 else None for j in row] for row in positions.cpu().tolist()]
