"""CPU oracle for the AdvancedNCF training + scoring hot path.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this module, and only as the checker (or as the
CPU arm that is *reported*), never as the thing shipped: the product path
(`neural-collaborative-filtering-demo_b200/`) never imports `oracle/` and raises if its CUDA
library is missing.

What it is: a plain-PyTorch fp32, single-thread-agnostic restatement of the reference's
algorithm.  Every function cites the reference `file:line` it follows (paths relative to the
reference repo root).  The arithmetic that the reference delegates to third-party code
(`torchrec==0.8.0` EmbeddingBagCollection / KeyedJaggedTensor, pinned at `Dockerfile:22-27`;
`torch.optim.Adam`, `nn.BCELoss`) is restated from the published behaviour: a SUM-pooled bag of
length 1 is a plain row gather; KJT values are key-major.

Parity pin (see DESIGN.md "Oracle"): `tests/test_oracle_golden.py` checks this file against
  * the reference's own shipped artefacts - 1000 golden `(user,item)->score` rows of
    `src/inference/demo/data/predictions.csv` produced by the shipped checkpoint
    `src/inference/demo/train_20241225_002713_model/` (committed in compact form under
    `tests/golden/golden_ckpt_compact.npz`);
  * fixtures generated in the build container by importing the UNMODIFIED reference modules
    (`oracle/make_golden.py`, through `oracle/shims/`): forward in train and eval mode, recorded
    dropout masks, BCELoss, dense gradients, two `torch.optim.Adam` steps, `forward_simple` with
    and without the hour path, `nlargest` top-k order, `calculate_metrics` values.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Params = Dict[str, torch.Tensor]

K_UMF = "mf_embedding_collection.embedding_bags.user_id.weight"
K_PMF = "mf_embedding_collection.embedding_bags.product_id.weight"
K_UMLP = "mlp_embedding_collection.embedding_bags.user_id.weight"
K_PMLP = "mlp_embedding_collection.embedding_bags.product_id.weight"
TABLE_KEYS = (K_UMF, K_PMF, K_UMLP, K_PMLP)

# parameters that `forward` actually uses (83,909 dense values + 4 tables)
ACTIVE_DENSE_KEYS = (
    "mf_norm.weight", "mf_norm.bias", "mlp_norm.weight", "mlp_norm.bias",
    "user_product_attention.q_proj.weight", "user_product_attention.q_proj.bias",
    "user_product_attention.k_proj.weight", "user_product_attention.k_proj.bias",
    "user_product_attention.v_proj.weight", "user_product_attention.v_proj.bias",
    "user_product_attention.out_proj.weight", "user_product_attention.out_proj.bias",
    "mlp.0.weight", "mlp.0.bias", "mlp.2.weight", "mlp.2.bias",
    "mlp.4.weight", "mlp.4.bias", "mlp.6.weight", "mlp.6.bias",
    "mlp.8.weight", "mlp.8.bias", "mlp.10.weight", "mlp.10.bias",
    "mf_output.weight", "mf_output.bias", "mlp_output.weight", "mlp_output.bias",
    "final.0.weight", "final.0.bias",
)


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def layer_norm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """nn.LayerNorm(d): biased variance, eps inside the sqrt, affine (architecture.py:255-256)."""
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) * torch.rsqrt(var + eps) * w + b


def linear(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> torch.Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def apply_keep_mask(x: torch.Tensor, keep: Optional[torch.Tensor], p: float) -> torch.Tensor:
    """Inverted dropout with an explicit keep mask (nn.Dropout semantics: kept values / (1-p))."""
    if keep is None or p == 0.0:
        return x
    return x * keep.to(x.dtype) / (1.0 - p)


def multi_head_attention(p: Params, prefix: str, query: torch.Tensor, key: torch.Tensor,
                         value: torch.Tensor, num_heads: int,
                         keep: Optional[torch.Tensor] = None, dropout_p: float = 0.0) -> torch.Tensor:
    """MultiHeadAttention.forward (architecture.py:35-57).

    query/key/value: [B, S, E].  keep: optional bool [B, H, S, S] mask for the dropout applied to
    the softmax probabilities (architecture.py:51).
    """
    B, S, E = query.shape
    hd = E // num_heads
    q = linear(query, p[prefix + ".q_proj.weight"], p[prefix + ".q_proj.bias"]).view(B, -1, num_heads, hd).transpose(1, 2)
    k = linear(key, p[prefix + ".k_proj.weight"], p[prefix + ".k_proj.bias"]).view(B, -1, num_heads, hd).transpose(1, 2)
    v = linear(value, p[prefix + ".v_proj.weight"], p[prefix + ".v_proj.bias"]).view(B, -1, num_heads, hd).transpose(1, 2)
    scores = torch.matmul(q, k.transpose(-2, -1)) / math.sqrt(hd)          # :45 (scale = sqrt(head_dim) :33)
    w = torch.softmax(scores, dim=-1)                                      # :50
    w = apply_keep_mask(w, keep, dropout_p)                                # :51
    out = torch.matmul(w, v).transpose(1, 2).contiguous().view(B, -1, E)    # :54-55
    return linear(out, p[prefix + ".out_proj.weight"], p[prefix + ".out_proj.bias"])  # :57


def mlp_tower(p: Params, x: torch.Tensor, keeps: Optional[Sequence[Optional[torch.Tensor]]] = None,
              dropout_p: float = 0.0) -> torch.Tensor:
    """self.mlp: 3 x (Linear -> ReLU -> LayerNorm -> Dropout) (architecture.py:230-242)."""
    for li, base in enumerate((0, 4, 8)):
        x = linear(x, p[f"mlp.{base}.weight"], p[f"mlp.{base}.bias"])
        x = torch.relu(x)
        x = layer_norm(x, p[f"mlp.{base + 2}.weight"], p[f"mlp.{base + 2}.bias"])
        x = apply_keep_mask(x, None if keeps is None else keeps[li], dropout_p)
    return x


# --------------------------------------------------------------------------------------------
# AdvancedNCF.forward / forward_simple
# --------------------------------------------------------------------------------------------
def forward(p: Params, user_ids: torch.Tensor, item_ids: torch.Tensor, *, training: bool,
            negative_samples: int = 4, num_heads: int = 4, temporal_dim: int = 32,
            dropout_p: float = 0.0, masks: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """AdvancedNCF.forward (architecture.py:258-381) on the two id columns of the key-major KJT
    (`values = [users..., items...]`, data_prep.py:286-298).  Returns probabilities [N, 1].

    masks (train only): {"attn": bool [B,H,S,S], "mlp0": bool [N,256], "mlp1": [N,128], "mlp2": [N,64]}.
    """
    N = user_ids.numel()
    S = 1 + negative_samples if training else 1                            # :275
    if N % S != 0:
        raise ValueError(f"{N} sample rows do not form groups of {S}")
    B = N // S                                                              # :276
    dp = dropout_p if training else 0.0
    masks = masks or {}

    u_mf = layer_norm(p[K_UMF][user_ids], p["mf_norm.weight"], p["mf_norm.bias"])      # :286,305
    i_mf = layer_norm(p[K_PMF][item_ids], p["mf_norm.weight"], p["mf_norm.bias"])      # :306
    mf_pred = linear(u_mf * i_mf, p["mf_output.weight"], p["mf_output.bias"])          # :307-308

    u_mlp = layer_norm(p[K_UMLP][user_ids], p["mlp_norm.weight"], p["mlp_norm.bias"]).view(B, S, -1)   # :311,315
    i_mlp = layer_norm(p[K_PMLP][item_ids], p["mlp_norm.weight"], p["mlp_norm.bias"]).view(B, S, -1)   # :312,316
    attn = multi_head_attention(p, "user_product_attention", u_mlp, i_mlp, i_mlp, num_heads,
                                masks.get("attn"), dp).reshape(N, -1)                  # :319-326
    x = torch.cat([attn, torch.zeros(N, temporal_dim, dtype=attn.dtype, device=attn.device)], dim=1)   # :329-340
    h = mlp_tower(p, x, [masks.get("mlp0"), masks.get("mlp1"), masks.get("mlp2")], dp) # :344
    mlp_pred = linear(h, p["mlp_output.weight"], p["mlp_output.bias"])                 # :345
    z = linear(torch.cat([mf_pred, mlp_pred], dim=1), p["final.0.weight"], p["final.0.bias"])  # :353
    return torch.sigmoid(z)                                                            # :354


def forward_simple(p: Params, user_ids: torch.Tensor, item_ids: torch.Tensor,
                   hour: Optional[torch.Tensor] = None,
                   temporal_proj: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                   num_heads: int = 4, temporal_dim: int = 32) -> torch.Tensor:
    """AdvancedNCF.forward_simple (architecture.py:409-485); eval-mode semantics (no dropout).

    `temporal_proj` = (weight [64,32], bias [64]) of the *fresh* nn.Linear the reference builds on
    every call when temporal_dim != mf_embedding_dim (architecture.py:436-442); the caller draws it.
    """
    n = user_ids.numel()
    u_mf = layer_norm(p[K_UMF][user_ids], p["mf_norm.weight"], p["mf_norm.bias"])      # :429
    i_mf = layer_norm(p[K_PMF][item_ids], p["mf_norm.weight"], p["mf_norm.bias"])      # :430
    u_mlp = layer_norm(p[K_UMLP][user_ids], p["mlp_norm.weight"], p["mlp_norm.bias"])  # :451
    i_mlp = layer_norm(p[K_PMLP][item_ids], p["mlp_norm.weight"], p["mlp_norm.bias"])  # :452
    if hour is not None:
        t = p["temporal_encoding.hour_embed.weight"][hour]                             # :434
        tt = t
        if t.shape[-1] != u_mf.shape[-1]:                                              # :436
            if temporal_proj is None:
                raise ValueError("hour path needs the fresh temporal projection")
            tt = linear(t, temporal_proj[0], temporal_proj[1])                         # :442
        i_mf = i_mf * (1 + 0.3 * tt)                                                   # :444
        i_mlp = i_mlp * (1 + 0.3 * tt)                                                 # :456
        tail = t                                                                       # :467
    else:
        tail = torch.zeros(n, temporal_dim, dtype=u_mf.dtype, device=u_mf.device)     # :471-476
    mf_pred = linear(u_mf * i_mf, p["mf_output.weight"], p["mf_output.bias"])          # :447-448
    attn = multi_head_attention(p, "user_product_attention", u_mlp.unsqueeze(1), i_mlp.unsqueeze(1),
                                i_mlp.unsqueeze(1), num_heads).squeeze(1)              # :459-463
    h = mlp_tower(p, torch.cat([attn, tail], dim=1))                                   # :468/477,480
    mlp_pred = linear(h, p["mlp_output.weight"], p["mlp_output.bias"])                 # :481
    z = linear(torch.cat([mf_pred, mlp_pred], dim=1), p["final.0.weight"], p["final.0.bias"])
    return torch.sigmoid(z).squeeze(-1)                                                # :484-485


def temporal_encoding(p: Params, hour, day, month, days_since, max_period: int = 365) -> torch.Tensor:
    """TemporalEncoding.forward (architecture.py:86-94)."""
    t = (p["temporal_encoding.hour_embed.weight"][hour] + p["temporal_encoding.day_embed.weight"][day]
         + p["temporal_encoding.month_embed.weight"][month])
    return t + p["temporal_encoding.pe"][days_since.long() % max_period]


def sinusoid_table(dim: int, max_period: int = 365) -> torch.Tensor:
    """The `pe` buffer (architecture.py:79-84)."""
    position = torch.arange(max_period).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, dim, 2) * (-math.log(10000.0) / dim))
    pe = torch.zeros(max_period, dim)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def category_hierarchy(p: Params, department_ids: torch.Tensor, category_ids: torch.Tensor,
                       num_heads: int = 4) -> torch.Tensor:
    """CategoryHierarchy.forward in eval mode (architecture.py:111-119)."""
    d = p["category_hierarchy.department_embed.weight"][department_ids]
    c = p["category_hierarchy.category_embed.weight"][category_ids]
    # MultiHeadAttention views its inputs as (batch=shape[0], -1, H, hd) (architecture.py:40-42), so
    # for [n, E] ids the result is [n, 1, E]; the residual `+ cat_embeds` then BROADCASTS to
    # [n, n, E] (architecture.py:119).  Restated as-is: callers use n == 1 (generate_embeddings.py:107-120).
    if d.dim() == 2:
        h = multi_head_attention(p, "category_hierarchy.hierarchy_attn", c.unsqueeze(1), d.unsqueeze(1),
                                 d.unsqueeze(1), num_heads)
    else:
        h = multi_head_attention(p, "category_hierarchy.hierarchy_attn", c, d, d, num_heads)
    return layer_norm(h + c, p["category_hierarchy.norm.weight"], p["category_hierarchy.norm.bias"])


def get_user_embeddings(p: Params, user_ids: torch.Tensor) -> Dict[str, torch.Tensor]:
    """architecture.py:383-391."""
    return {"mf": layer_norm(p[K_UMF][user_ids], p["mf_norm.weight"], p["mf_norm.bias"]),
            "mlp": layer_norm(p[K_UMLP][user_ids], p["mlp_norm.weight"], p["mlp_norm.bias"])}


def get_product_embeddings(p: Params, item_ids, department_ids, category_ids) -> Dict[str, torch.Tensor]:
    """architecture.py:393-407."""
    return {"mf": layer_norm(p[K_PMF][item_ids], p["mf_norm.weight"], p["mf_norm.bias"]),
            "mlp": layer_norm(p[K_PMLP][item_ids], p["mlp_norm.weight"], p["mlp_norm.bias"]),
            "category": category_hierarchy(p, department_ids, category_ids)}


# --------------------------------------------------------------------------------------------
# loss / optimizer (trainer.py:71-81, 271-285)
# --------------------------------------------------------------------------------------------
def bce_loss(pred: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.BCELoss(): mean of -[y log p + (1-y) log(1-p)], each log clamped at -100 (trainer.py:78,271)."""
    lp = torch.clamp(torch.log(pred), min=-100.0)
    l1p = torch.clamp(torch.log(1.0 - pred), min=-100.0)
    return -(target * lp + (1.0 - target) * l1p).mean()


def adam_step_(params: Params, grads: Params, state: Dict[str, Dict[str, torch.Tensor]], *, step: int,
               lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
               weight_decay: float = 1e-5) -> None:
    """torch.optim.Adam single-tensor update, L2-coupled weight decay, applied in place to every
    tensor that has a gradient - including every row of the dense table gradients
    (trainer.py:71-75, 285).  `step` is 1-based."""
    b1, b2 = betas
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    for k, g in grads.items():
        if g is None:
            continue
        w = params[k]
        st = state.setdefault(k, {"m": torch.zeros_like(w), "v": torch.zeros_like(w)})
        if weight_decay != 0.0:
            g = g + weight_decay * w
        st["m"].lerp_(g, 1.0 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        w.addcdiv_(st["m"], denom, value=-(lr / bc1))


def train_step(params: Params, state, step: int, user_ids, item_ids, targets, *, negative_samples=4,
               dropout_p=0.0, masks=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5,
               update_keys: Optional[Sequence[str]] = None):
    """One iteration of ModelTrainer.train_epoch's loop body (trainer.py:253-289): forward, BCELoss,
    zero_grad, backward (dense table grads), Adam.  Returns (loss, probabilities, grads)."""
    keys = list(update_keys) if update_keys is not None else list(TABLE_KEYS) + list(ACTIVE_DENSE_KEYS)
    leaves = {k: params[k].detach().clone().requires_grad_(True) for k in keys}
    p = dict(params)
    p.update(leaves)
    out = forward(p, user_ids, item_ids, training=True, negative_samples=negative_samples,
                  dropout_p=dropout_p, masks=masks)
    loss = bce_loss(out, targets.view_as(out))
    gl = torch.autograd.grad(loss, [leaves[k] for k in keys], allow_unused=True)
    grads = {k: g for k, g in zip(keys, gl)}
    with torch.no_grad():
        adam_step_(params, grads, state, step=step, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
    return loss.detach(), out.detach(), grads


# --------------------------------------------------------------------------------------------
# scoring / top-k (app.py:43-77) and the eval-mode factorisation
# --------------------------------------------------------------------------------------------
def topk_stable(scores: torch.Tensor, k: int) -> torch.Tensor:
    """DataFrame.nlargest(k, 'score') order (app.py:75): score descending, ties -> lowest index."""
    order = torch.argsort(scores, dim=-1, descending=True, stable=True)
    return order[..., :k]


def item_fold(p: Params, num_heads: int = 4, temporal_dim: int = 32):
    """Eval mode has ONE key per query so softmax == 1 and the MLP tower depends on the item only
    (architecture.py:275-276, 315-323).  Returns (P_hat [I,64], g [I]) with
        logit(u, i) = LN_mf(U_mf[u]) . P_hat[i] + g[i]
    P_hat = a * LN_mf(P_mf) * w_mf ;  g = a*b_mf + c*mlp_pred(i) + d   (final.0 = [a, c], bias d)."""
    I = p[K_PMF].shape[0]
    a = p["final.0.weight"][0, 0]
    c = p["final.0.weight"][0, 1]
    d = p["final.0.bias"][0]
    i_mf = layer_norm(p[K_PMF], p["mf_norm.weight"], p["mf_norm.bias"])
    p_hat = a * i_mf * p["mf_output.weight"][0]
    i_mlp = layer_norm(p[K_PMLP], p["mlp_norm.weight"], p["mlp_norm.bias"])
    v = linear(i_mlp, p["user_product_attention.v_proj.weight"], p["user_product_attention.v_proj.bias"])
    attn = linear(v, p["user_product_attention.out_proj.weight"], p["user_product_attention.out_proj.bias"])
    h = mlp_tower(p, torch.cat([attn, torch.zeros(I, temporal_dim)], dim=1))
    mlp_pred = linear(h, p["mlp_output.weight"], p["mlp_output.bias"])[:, 0]
    g = a * p["mf_output.bias"][0] + c * mlp_pred + d
    return p_hat, g


def score_catalogue(p: Params, user_ids: torch.Tensor, chunk: int = 1 << 20) -> torch.Tensor:
    """Full-catalogue scores [len(user_ids), I] exactly as app.py:48-67 computes them
    (forward_simple over all items for each user), evaluated through forward_simple."""
    I = p[K_PMF].shape[0]
    items = torch.arange(I)
    rows = []
    for u in user_ids.tolist():
        parts = []
        for s in range(0, I, chunk):
            it = items[s:s + chunk]
            parts.append(forward_simple(p, torch.full((it.numel(),), u, dtype=torch.long), it))
        rows.append(torch.cat(parts))
    return torch.stack(rows)


# --------------------------------------------------------------------------------------------
# sharding (torchrec ROW_WISE convention) and id remapping (8a A2)
# --------------------------------------------------------------------------------------------
def row_shard(ids: torch.Tensor, rows: int, world: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """block = ceil(rows/world); owner = id // block; local = id % block (integer, bit-exact)."""
    block = (rows + world - 1) // world
    return ids // block, ids % block, block


def remap_product_id(pid: str, num_products: int) -> int:
    """local_inference.py:55,65 / generate_embeddings.py:104: int(pid.lstrip('P'), 16) % num_products."""
    return int(pid.lstrip("P"), 16) % num_products


def remap_cardnumber(card: str, num_users: int) -> int:
    """local_inference.py:51: int(cardnumber) % num_users."""
    return int(card) % num_users


def first_appearance_index(values: Sequence) -> Dict:
    """SheetzDataset.user_to_idx / product_to_idx: enumerate(unique()) in first-appearance order
    (data_prep.py:65-71)."""
    out: Dict = {}
    for v in values:
        if v not in out:
            out[v] = len(out)
    return out


# --------------------------------------------------------------------------------------------
# ranking metrics (utils/metrics.py) - next-row N2, restated for "matched AUC/HR@10"
# --------------------------------------------------------------------------------------------
def ranking_metrics(pred: torch.Tensor, target: torch.Tensor, k_values: Sequence[int], batch_size: int,
                    negative_samples: int) -> Dict[str, float]:
    """HR@K / NDCG@K / MRR@K over [batch_size, 1+negative_samples] groups with exactly one positive
    per group (utils/metrics.py:110-205), plus AUC (rank-sum form of sklearn.roc_auc_score, :244-265)
    and accuracy at threshold 0.5 (:267-275)."""
    M = 1 + negative_samples
    P = pred.reshape(batch_size, M).float()
    T = target.reshape(batch_size, M).float()
    out: Dict[str, float] = {}
    for k in k_values:
        kk = min(k, M)
        top = torch.topk(P, kk, dim=1).indices                       # metrics.py:124
        hit = torch.gather(T, 1, top)                                 # [batch, kk]
        out[f"hit_rate@{k}"] = float((hit.sum(1) > 0).float().mean())
        disc = 1.0 / torch.log2(torch.arange(kk, dtype=torch.float32) + 2.0)
        dcg = (hit * disc).sum(1)
        ideal = torch.sort(T, dim=1, descending=True).values[:, :kk]
        idcg = (ideal * disc).sum(1)
        ndcg = torch.where(idcg > 0, dcg / idcg.clamp_min(1e-12), torch.zeros_like(dcg))
        out[f"ndcg@{k}"] = float(ndcg.mean())
        first = torch.where(hit > 0, 1.0 / (torch.arange(kk, dtype=torch.float32) + 1.0), torch.zeros(1))
        out[f"mrr@{k}"] = float(first.max(dim=1).values.mean())
    out["auc"] = auc(pred.reshape(-1), target.reshape(-1))
    out["accuracy"] = float(((pred.reshape(-1) >= 0.5).float() == target.reshape(-1).float()).float().mean())
    return out


def auc(pred: torch.Tensor, target: torch.Tensor) -> float:
    """ROC AUC by the Mann-Whitney rank-sum with average ranks for ties (== sklearn.roc_auc_score)."""
    pred = pred.double()
    pos = target > 0.5
    n_pos = int(pos.sum())
    n_neg = pred.numel() - n_pos
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = torch.argsort(pred, stable=True)
    sp = pred[order]
    ranks = torch.empty_like(sp)
    # average ranks over tie groups
    uniq, inv, cnt = torch.unique_consecutive(sp, return_inverse=True, return_counts=True)
    end = torch.cumsum(cnt, 0).double()
    start = end - cnt.double() + 1.0
    ranks = ((start + end) / 2.0)[inv]
    r = torch.empty_like(ranks)
    r[order] = ranks
    return float((r[pos].sum() - n_pos * (n_pos + 1) / 2.0) / (n_pos * n_neg))


# --------------------------------------------------------------------------------------------
# input pipeline (data_prep.py) - next-row N1
# --------------------------------------------------------------------------------------------
def product_weights(item_idx, num_products: int):
    """inverse-popularity sampling weights (data_prep.py:95-102)."""
    import numpy as np
    counts = np.zeros(num_products)
    for i in item_idx:
        counts[int(i)] += 1
    counts = np.maximum(counts, 1)
    w = 1 / counts
    return w / w.sum()


def sample_negative(user_idx: int, positive_product: int, weights, history: dict, num_products: int) -> int:
    """SheetzDataset._sample_negative (data_prep.py:134-161), drawing from numpy's global RNG like the reference."""
    import numpy as np
    user_positives = history.get(user_idx, set())
    for _ in range(10):
        product_idx = np.random.choice(num_products, p=weights)
        if product_idx != positive_product and product_idx not in user_positives:
            return int(product_idx)
    valid = list(set(range(num_products)) - user_positives - {positive_product})
    if not valid:
        while True:
            idx = np.random.randint(0, num_products)
            if idx != positive_product:
                return int(idx)
    return int(np.random.choice(valid))


def consistent_batches(dataset_size: int, batch_size: int, shuffle: bool = True):
    """ConsistentBatchSampler.__iter__ (data_prep.py:419-440)."""
    import numpy as np
    indices = list(range(dataset_size))
    if shuffle:
        np.random.shuffle(indices)
    for i in range((dataset_size + batch_size - 1) // batch_size):
        b = indices[i * batch_size:min((i + 1) * batch_size, dataset_size)]
        if len(b) < batch_size:
            b = b + b[:batch_size - len(b)]
        yield b


def collate(samples):
    """collate_recommender_batch (data_prep.py:230-320): samples = [(user_ids[M], item_ids[M], targets[M])];
    returns (values key-major, lengths, targets [N,1])."""
    users = [int(u) for s in samples for u in s[0]]
    items = [int(i) for s in samples for i in s[1]]
    targets = [float(t) for s in samples for t in s[2]]
    values = torch.tensor(users + items, dtype=torch.long)
    return values, torch.ones(values.numel(), dtype=torch.long), torch.tensor(targets, dtype=torch.float32).unsqueeze(1)
