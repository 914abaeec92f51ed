"""AdvancedNCF with the reference's PyTorch surface and a CUDA (sm_100a) hot path.

Mirrors `src/model/architecture.py` of the reference: same constructor (:122-133), same 62-key
state_dict, same `forward(KeyedJaggedTensor) -> [N,1]` (:258),
`forward_simple(user_ids, product_ids, hour=None) -> [N]` (:409), `get_user_embeddings` (:383),
`get_product_embeddings` (:393), and the sub-modules the demo dashboard pokes (app.py:156-183).
All GPU work of forward/backward goes through libncf_b200.so (include/ncf_b200.h); there is no
CPU or eager-PyTorch fallback: calling the model with CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import _lib
from .kjt import KeyedJaggedTensor

TABLE_KEYS = ("mf_embedding_collection.embedding_bags.user_id.weight",
              "mf_embedding_collection.embedding_bags.product_id.weight",
              "mlp_embedding_collection.embedding_bags.user_id.weight",
              "mlp_embedding_collection.embedding_bags.product_id.weight")


class MultiHeadAttention(nn.Module):
    """Parameter container + eager forward of reference MultiHeadAttention (architecture.py:18-57).
    The hot path never calls this forward (the attention runs inside the CUDA library); it serves
    CategoryHierarchy (export path) and the dashboard (app.py:160-175)."""

    def __init__(self, embed_dim: int, num_heads: int = 4, dropout: float = 0.1):
        super().__init__()
        assert embed_dim % num_heads == 0, "embed_dim must be divisible by num_heads"
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.head_dim = embed_dim // num_heads
        self.q_proj = nn.Linear(embed_dim, embed_dim)
        self.k_proj = nn.Linear(embed_dim, embed_dim)
        self.v_proj = nn.Linear(embed_dim, embed_dim)
        self.out_proj = nn.Linear(embed_dim, embed_dim)
        self.dropout = nn.Dropout(dropout)
        self.scale = math.sqrt(self.head_dim)

    def forward(self, query, key, value, mask=None):
        b = query.shape[0]
        q = self.q_proj(query).view(b, -1, self.num_heads, self.head_dim).transpose(1, 2)
        k = self.k_proj(key).view(b, -1, self.num_heads, self.head_dim).transpose(1, 2)
        v = self.v_proj(value).view(b, -1, self.num_heads, self.head_dim).transpose(1, 2)
        scores = torch.matmul(q, k.transpose(-2, -1)) / self.scale
        if mask is not None:
            scores = scores.masked_fill(mask == 0, float("-inf"))
        w = self.dropout(torch.softmax(scores, dim=-1))
        out = torch.matmul(w, v).transpose(1, 2).contiguous().view(b, -1, self.embed_dim)
        return self.out_proj(out)


class TemporalEncoding(nn.Module):
    """reference architecture.py:59-94; forward runs ncf_temporal_fwd."""

    def __init__(self, embed_dim: int, max_period: int = 365):
        super().__init__()
        self.embed_dim = embed_dim
        self.max_period = max_period
        self.hour_embed = nn.Embedding(24, embed_dim)
        self.day_embed = nn.Embedding(7, embed_dim)
        self.month_embed = nn.Embedding(12, embed_dim)
        position = torch.arange(max_period).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, embed_dim, 2) * (-math.log(10000.0) / embed_dim))
        pe = torch.zeros(max_period, embed_dim)
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)

    def forward(self, hour, day, month, days_since):
        w = self.hour_embed.weight
        if not w.is_cuda:
            raise _lib.NcfError("TemporalEncoding runs on CUDA only (no CPU fallback)")
        if self.embed_dim != 32 or self.max_period != 365:
            raise NotImplementedError("the CUDA kernel is built for temporal_dim=32, max_period=365")
        lib = _lib.load()
        shape = hour.shape
        ids = [t.reshape(-1).to(device=w.device, dtype=torch.long).contiguous() for t in (hour, day, month, days_since)]
        for t, hi, name in zip(ids[:3], (24, 7, 12), ("hour", "day", "month")):     # nn.Embedding raises (architecture.py:88-90)
            if t.numel() and (int(t.min()) < 0 or int(t.max()) >= hi):
                raise IndexError(f"TemporalEncoding: {name} index out of range [0, {hi})")
        n = ids[0].numel()
        out = torch.empty(n, self.embed_dim, device=w.device, dtype=torch.float32)
        _lib.check(lib.ncf_temporal_fwd(_lib.ptr(w), _lib.ptr(self.day_embed.weight), _lib.ptr(self.month_embed.weight),
                                        _lib.ptr(self.pe), *[_lib.ptr(t) for t in ids], n, _lib.ptr(out),
                                        _stream(w.device)), "ncf_temporal_fwd")
        return out.view(*shape, self.embed_dim)


class CategoryHierarchy(nn.Module):
    """reference architecture.py:96-119 (export path only; eager, incl. its [n,1,E]+[n,E] broadcast)."""

    def __init__(self, num_departments: int, num_categories: int, embed_dim: int, dropout: float = 0.1):
        super().__init__()
        self.department_embed = nn.Embedding(num_departments, embed_dim)
        self.category_embed = nn.Embedding(num_categories, embed_dim)
        self.hierarchy_attn = MultiHeadAttention(embed_dim, num_heads=4, dropout=dropout)
        self.norm = nn.LayerNorm(embed_dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, department_ids, category_ids):
        dept = self.department_embed(department_ids)
        cat = self.category_embed(category_ids)
        h = self.dropout(self.hierarchy_attn(cat, dept, dept))
        return self.norm(h + cat)


class _KeyedRows:
    def __init__(self, d):
        self._d = d

    def __getitem__(self, k):
        return self._d[k]

    def keys(self):
        return list(self._d.keys())


class EmbeddingTables(nn.Module):
    """Stand-in for torchrec.EmbeddingBagCollection (architecture.py:153-190): one [rows,dim] table
    per feature under `embedding_bags.<name>.weight`, init U(+-sqrt(1/rows)).  Calling it returns the
    raw rows per feature (bags of length 1); the model's own forward does not use this call."""

    def __init__(self, sizes: Dict[str, int], dim: int):
        super().__init__()
        self.embedding_bags = nn.ModuleDict()
        for name, rows in sizes.items():
            bag = nn.EmbeddingBag(rows, dim, mode="sum", include_last_offset=True)
            with torch.no_grad():
                bag.weight.uniform_(-math.sqrt(1.0 / rows), math.sqrt(1.0 / rows))
            self.embedding_bags[name] = bag

    def forward(self, features):
        vals = features.values()
        keys = list(features.keys())
        n = vals.numel() // len(keys)
        return _KeyedRows({k: self.embedding_bags[k].weight.index_select(0, vals[i * n:(i + 1) * n])
                           for i, k in enumerate(keys)})


def _stream(device):
    """torch's current stream of `device` as the C ABI's stream argument.  The raw accessor is one C call (~0.3 us);
    torch.cuda.current_stream builds a Stream object (~4 us), and a step passes the stream to two dozen calls."""
    index = device.index if isinstance(device, torch.device) else device
    if index is None:
        index = torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(index))


class _NCFFunction(torch.autograd.Function):
    """forward = ncf_forward, backward = ncf_backward (dense grads returned as views of one flat
    buffer; table grads materialised or fused with Adam inside the library)."""

    @staticmethod
    def forward(ctx, module, user_ids, item_ids, S, training, *params):
        lib = _lib.load()
        dev = user_ids.device
        N = user_ids.numel()
        cfg = module._run_cfg(S, training)
        ws_bytes = int(lib.ncf_workspace_bytes(N, C.byref(cfg)))
        needs_grad = training and torch.is_grad_enabled() and any(p.requires_grad for p in params)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev) if needs_grad else module._scratch(ws_bytes, dev)
        out = torch.empty(N, dtype=torch.float32, device=dev)
        tables = module._tables_struct()
        _lib.check(lib.ncf_forward(C.byref(cfg), C.byref(tables), _lib.ptr(module._flat), _lib.ptr(user_ids),
                                   _lib.ptr(item_ids), N, None, None, None, _lib.ptr(out), _lib.ptr(ws), ws_bytes,
                                   _stream(dev)), "ncf_forward")
        ctx.module, ctx.cfg, ctx.ws, ctx.N = module, cfg, ws, N
        ctx.save_for_backward(user_ids, item_ids)
        ctx.nparams = len(params)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        module = ctx.module
        user_ids, item_ids = ctx.saved_tensors
        dev = user_ids.device
        grad_out = grad_out.contiguous().float()
        dense_grad = torch.zeros(module._flat.numel(), dtype=torch.float32, device=dev)
        adam, tables, table_grads = module._backward_cfg()
        _lib.check(lib.ncf_backward(C.byref(ctx.cfg), C.byref(adam), C.byref(tables), _lib.ptr(module._flat),
                                    _lib.ptr(dense_grad), _lib.ptr(user_ids), _lib.ptr(item_ids), ctx.N,
                                    _lib.ptr(grad_out), _lib.ptr(ctx.ws), ctx.ws.numel(), _stream(dev)), "ncf_backward")
        ctx.ws = None
        grads = [dense_grad[off:off + size].view(shape) for off, size, shape in module._flat_views]
        grads += table_grads
        assert len(grads) == ctx.nparams
        return (None, None, None, None, None, *grads)


class AdvancedNCF(nn.Module):
    def __init__(self,
                 num_users: int,
                 num_products: int,
                 num_departments: int,
                 num_categories: int,
                 mf_embedding_dim: int = 64,
                 mlp_embedding_dim: int = 64,
                 temporal_dim: int = 32,
                 mlp_hidden_dims: List[int] = [256, 128, 64],
                 num_heads: int = 4,
                 dropout: float = 0.2,
                 negative_samples: int = 4):
        super().__init__()
        self.num_users = num_users
        self.num_products = num_products
        self.num_departments = num_departments
        self.num_categories = num_categories
        self.mf_embedding_dim = mf_embedding_dim
        self.mlp_embedding_dim = mlp_embedding_dim
        self.temporal_dim = temporal_dim
        self.mlp_hidden_dims = list(mlp_hidden_dims)
        self.num_heads = num_heads
        self.dropout = dropout
        self.negative_samples = negative_samples

        # registration order follows the reference so that state_dict() enumerates the same keys
        self.mf_norm = nn.LayerNorm(mf_embedding_dim)
        self.mlp_norm = nn.LayerNorm(mlp_embedding_dim)
        sizes = {"user_id": num_users, "product_id": num_products}
        self.mf_embedding_collection = EmbeddingTables(sizes, mf_embedding_dim)
        self.mlp_embedding_collection = EmbeddingTables(sizes, mlp_embedding_dim)
        self.category_hierarchy = CategoryHierarchy(num_departments, num_categories, mlp_embedding_dim, dropout)
        self.temporal_encoding = TemporalEncoding(temporal_dim)
        self.user_product_attention = MultiHeadAttention(mlp_embedding_dim, num_heads, dropout)
        self.sequence_attention = MultiHeadAttention(mlp_embedding_dim, num_heads, dropout)
        combined_dim = mlp_embedding_dim + temporal_dim
        self.feature_combination = nn.Sequential(nn.Linear(combined_dim, mlp_hidden_dims[0]), nn.ReLU(),
                                                 nn.LayerNorm(mlp_hidden_dims[0]), nn.Dropout(dropout))
        layers, cur = [], combined_dim
        for h in mlp_hidden_dims:
            layers += [nn.Linear(cur, h), nn.ReLU(), nn.LayerNorm(h), nn.Dropout(dropout)]
            cur = h
        self.mlp = nn.Sequential(*layers)
        self.mf_output = nn.Linear(mf_embedding_dim, 1)
        self.mlp_output = nn.Linear(mlp_hidden_dims[-1], 1)
        self.final = nn.Sequential(nn.Linear(2, 1), nn.Sigmoid())

        # ---- CUDA-path state (not part of the state_dict contract) ----
        self.compute_precision = "fp32"          # "fp32" | "bf16" (tcgen05 towers)
        self._flat: Optional[torch.Tensor] = None
        self._flat_views = []
        self._scratch_buf: Optional[torch.Tensor] = None
        self._dropout_seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        self._fwd_calls = 0
        self._table_mode = "autograd"            # "autograd" | "fused_dense_equiv" | "fused_sparse"
        self._table_opt = None
        self._table_hparams = None
        self._table_state = None
        self._table_step = 0
        self._status: Optional[_lib.StatusWord] = None    # id-validation words the kernels set (pinned host memory)

    # ------------------------------------------------------------------------------------------
    # plumbing
    # ------------------------------------------------------------------------------------------
    def _check_geometry(self):
        if (self.mf_embedding_dim, self.mlp_embedding_dim, self.temporal_dim, self.mlp_hidden_dims, self.num_heads) != \
                (64, 64, 32, [256, 128, 64], 4):
            raise NotImplementedError(
                "libncf_b200 is built for the reference configuration (config/config.yaml:56-61): embedding 64/64, "
                "temporal_dim 32, layers [256,128,64], 4 heads")

    def _table_params(self):
        return [self.mf_embedding_collection.embedding_bags["user_id"].weight,
                self.mf_embedding_collection.embedding_bags["product_id"].weight,
                self.mlp_embedding_collection.embedding_bags["user_id"].weight,
                self.mlp_embedding_collection.embedding_bags["product_id"].weight]

    def _dense_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in _lib.DENSE_KEYS]

    def _ensure_flat(self):
        """Keep the 30 dense tensors `forward` uses as views of ONE flat fp32 buffer (the layout of
        include/ncf_b200.h) so the library takes a single pointer; re-flattens after .to()/.cuda()."""
        self._check_geometry()
        params = self._dense_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise _lib.NcfError("AdvancedNCF runs on CUDA only: move the model to a B200 (no CPU fallback)")
        layout, numel = _lib.dense_layout()
        ok = self._flat is not None and self._flat.device == dev
        if ok:
            base = self._flat.data_ptr()
            ok = all(p.data_ptr() == base + 4 * off for p, (_, off, _) in zip(params, layout))
        if not ok:
            flat = torch.zeros(numel, dtype=torch.float32, device=dev)
            views = []
            with torch.no_grad():
                for p, (_, off, size) in zip(params, layout):
                    if p.numel() != size or p.dtype != torch.float32:
                        raise _lib.NcfError("dense parameter does not match the library layout")
                    flat[off:off + size].copy_(p.detach().reshape(-1))
                    p.data = flat[off:off + size].view(p.shape)
                    views.append((off, size, tuple(p.shape)))
            self._flat, self._flat_views = flat, views
        for t in self._table_params():
            if t.device != dev or not t.is_contiguous() or t.dtype != torch.float32:
                raise _lib.NcfError("embedding tables must be contiguous fp32 on the model's device")
        if self._status is None:
            self._status = _lib.StatusWord()

    def check_status(self, what: str = "AdvancedNCF"):
        """Raise IndexError if a kernel saw an id (or hour) outside its table since the last check - what
        nn.EmbeddingBag does synchronously on the CPU (reference architecture.py:286-287).  Reads pinned host memory:
        no device synchronisation; everything launched before the caller's last synchronisation is covered."""
        if self._status is not None:
            self._status.raise_if_set(what)

    def _scratch(self, nbytes, dev):
        if self._scratch_buf is None or self._scratch_buf.numel() < nbytes or self._scratch_buf.device != dev:
            self._scratch_buf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._scratch_buf

    def _run_cfg(self, S, training):
        cfg = _lib.RunCfg()
        cfg.S = S
        cfg.training = 1 if training else 0
        cfg.precision = _lib.NCF_BF16_TC if self.compute_precision == "bf16" else _lib.NCF_FP32
        cfg.dropout_p = float(self.dropout) if training else 0.0
        cfg.seed = self._dropout_seed
        if training:
            self._fwd_calls += 1
        cfg.step = self._fwd_calls
        return cfg

    def _tables_struct(self, grads=None):
        t = _lib.Tables()
        tabs = self._table_params()
        for i in range(4):
            t.w[i] = tabs[i].data_ptr()
            t.g[i] = grads[i].data_ptr() if grads is not None else None
            if self._table_state is not None:
                t.m[i] = self._table_state["m"][i].data_ptr()
                t.v[i] = self._table_state["v"][i].data_ptr()
        if self._table_state is not None:
            t.touched[0] = self._table_state["touched"][0].data_ptr()
            t.touched[1] = self._table_state["touched"][1].data_ptr()
        t.rows_user = self.num_users
        t.rows_item = self.num_products
        t.status = self._status.ptr() if self._status is not None else None
        return t

    def _backward_cfg(self):
        adam = _lib.AdamCfg()
        tabs = self._table_params()
        if self._table_mode == "autograd":
            if not any(t.requires_grad for t in tabs):
                adam.emb_mode = _lib.EMB_NONE
                return adam, self._tables_struct(), [None] * 4
            grads = [torch.zeros_like(t) for t in tabs]        # the reference's dense table gradient
            adam.emb_mode = _lib.EMB_MATERIALIZE
            return adam, self._tables_struct(grads), grads
        hp = dict(self._table_hparams)
        if self._table_opt is not None:                        # follow the caller's optimizer (lr schedules)
            g = self._table_opt.param_groups[0]
            hp.update(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])
        self._table_step += 1
        adam.lr, (adam.beta1, adam.beta2) = hp["lr"], hp["betas"]
        adam.eps, adam.weight_decay = hp["eps"], hp["weight_decay"]
        adam.step = self._table_step
        adam.emb_mode = _lib.EMB_ADAM_DENSE_EQUIV if self._table_mode == "fused_dense_equiv" else _lib.EMB_ADAM_SPARSE
        return adam, self._tables_struct(), [None] * 4

    def configure_table_optimizer(self, mode: str = "fused_dense_equiv", optimizer=None, lr: float = 1e-3,
                                  betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 1e-5):
        """How `loss.backward()` treats the four embedding tables.

        "autograd"           dense [rows,64] gradients land in `.grad` exactly like the reference
                             (nn.EmbeddingBag sparse=False) and the caller's optimizer updates them.
        "fused_dense_equiv"  the sorted-id scatter + Adam runs inside backward and every untouched
                             row receives the reference's g = wd*w Adam step: same weights as
                             torch.optim.Adam over dense gradients (trainer.py:71-75, 285).
        "fused_sparse"       touched rows only (NOT multi-step equivalent to the reference).
        In the fused modes the tables keep `.grad = None`, so the caller's torch optimizer skips
        them; hyper-parameters follow `optimizer.param_groups[0]` when an optimizer is given.
        """
        if mode not in ("autograd", "fused_dense_equiv", "fused_sparse"):
            raise ValueError(f"unknown table optimizer mode {mode!r}")
        self._table_mode = mode
        self._table_opt = optimizer
        self._table_hparams = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        if mode == "autograd":
            self._table_state = None
            return self
        tabs = self._table_params()
        if self._table_state is None or self._table_state["m"][0].device != tabs[0].device:
            self._table_state = {
                "m": [torch.zeros_like(t) for t in tabs],
                "v": [torch.zeros_like(t) for t in tabs],
                "touched": [torch.zeros(self.num_users, dtype=torch.uint8, device=tabs[0].device),
                            torch.zeros(self.num_products, dtype=torch.uint8, device=tabs[0].device)],
            }
            self._table_step = 0
        return self

    def table_optimizer_state_dict(self):
        if self._table_state is None:
            return {}
        return {"step": self._table_step, "m": [t.clone() for t in self._table_state["m"]],
                "v": [t.clone() for t in self._table_state["v"]]}

    def load_table_optimizer_state_dict(self, sd):
        if not sd:
            return
        if self._table_state is None:
            self.configure_table_optimizer(self._table_mode if self._table_mode != "autograd" else "fused_dense_equiv")
        self._table_step = int(sd["step"])
        for i in range(4):
            self._table_state["m"][i].copy_(sd["m"][i])
            self._table_state["v"][i].copy_(sd["v"][i])

    @staticmethod
    def _split_ids(features):
        values = features.values()
        total = values.size(0) // 2                 # architecture.py:274
        if values.size(0) != 2 * total or total == 0:
            raise ValueError(f"features must hold user_id and product_id values for every sample; got {values.size(0)}")
        if values.dtype != torch.long:
            values = values.long()
        return values[:total].contiguous(), values[total:2 * total].contiguous(), total

    # ------------------------------------------------------------------------------------------
    # reference surface
    # ------------------------------------------------------------------------------------------
    def forward(self, features: KeyedJaggedTensor) -> torch.Tensor:
        """reference AdvancedNCF.forward (architecture.py:258-381): probabilities [N, 1]."""
        user_ids, item_ids, total = self._split_ids(features)
        S = 1 + self.negative_samples if self.training else 1          # :275
        if S > _lib.MAX_S:
            raise NotImplementedError(f"1 + negative_samples = {S} exceeds the library limit {_lib.MAX_S}")
        if total % S != 0:
            raise ValueError(f"{total} samples cannot be viewed as groups of {S} "
                             "(reference: view(batch_size, samples_per_interaction, -1), architecture.py:315)")
        if not user_ids.is_cuda:
            raise _lib.NcfError("AdvancedNCF.forward needs CUDA tensors (no CPU fallback); call features.to('cuda')")
        self._ensure_flat()
        self.check_status("AdvancedNCF.forward (an earlier call)")
        params = self._dense_params() + self._table_params()
        out = _NCFFunction.apply(self, user_ids, item_ids, S, self.training, *params)
        outputs = out.view(total, 1)
        if outputs.shape != (total, 1):                                # :356-362
            raise ValueError(f"Output shape mismatch: got {outputs.shape}, expected {(total, 1)}")
        return outputs

    @torch.no_grad()
    def forward_simple(self, user_ids, product_ids, hour=None):
        """reference forward_simple (architecture.py:409-485): scores [N]; inference API (no autograd).
        With `hour`, a FRESH nn.Linear(temporal_dim, 64) is drawn per call exactly like the reference
        (:436-442) and folded into two 24-row tables by ncf_temporal_tables."""
        if not user_ids.is_cuda:
            raise _lib.NcfError("AdvancedNCF.forward_simple needs CUDA tensors (no CPU fallback)")
        self._ensure_flat()
        self.check_status("AdvancedNCF.forward_simple (an earlier call)")
        lib = _lib.load()
        dev = user_ids.device
        u = user_ids.reshape(-1).long().contiguous()
        p = product_ids.reshape(-1).long().contiguous()
        if u.numel() != p.numel():
            raise ValueError("user_ids and product_ids must have the same length")
        N = u.numel()
        cfg = self._run_cfg(1, self.training)
        hour_t = tmod = tail1 = None
        if hour is not None:
            hour_t = hour.reshape(-1).to(device=dev, dtype=torch.long).contiguous()
            proj = nn.Linear(self.temporal_dim, self.mf_embedding_dim, device=dev)          # :437-441
            tmod = torch.empty(24, 64, device=dev)
            tail1 = torch.empty(24, 256, device=dev)
            _lib.check(lib.ncf_temporal_tables(_lib.ptr(self.temporal_encoding.hour_embed.weight), _lib.ptr(proj.weight),
                                               _lib.ptr(proj.bias), _lib.ptr(self._flat), _lib.ptr(tmod), _lib.ptr(tail1),
                                               _stream(dev)), "ncf_temporal_tables")
        out = torch.empty(N, dtype=torch.float32, device=dev)
        if N == 0:
            return out
        ws_bytes = int(lib.ncf_workspace_bytes(N, C.byref(cfg)))
        ws = self._scratch(ws_bytes, dev)
        tables = self._tables_struct()
        _lib.check(lib.ncf_forward(C.byref(cfg), C.byref(tables), _lib.ptr(self._flat), _lib.ptr(u), _lib.ptr(p), N,
                                   _lib.ptr(hour_t), _lib.ptr(tmod), _lib.ptr(tail1), _lib.ptr(out), _lib.ptr(ws),
                                   ws_bytes, _stream(dev)), "ncf_forward")
        return out

    def _gather_ln(self, side: int, ids: torch.Tensor):
        self._ensure_flat()
        lib = _lib.load()
        ids = ids.reshape(-1).long().contiguous()
        n = ids.numel()
        mf = torch.empty(n, 64, device=ids.device)
        mlp = torch.empty(n, 64, device=ids.device)
        tables = self._tables_struct()
        _lib.check(lib.ncf_gather_ln(C.byref(tables), _lib.ptr(self._flat), side, _lib.ptr(ids), n, _lib.ptr(mf),
                                     _lib.ptr(mlp), _stream(ids.device)), "ncf_gather_ln")
        return mf, mlp

    @torch.no_grad()
    def get_user_embeddings(self, user_features: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """reference architecture.py:383-391."""
        u, _, _ = self._split_ids(user_features["user_features"])
        mf, mlp = self._gather_ln(0, u)
        return {"mf": mf, "mlp": mlp}

    @torch.no_grad()
    def get_product_embeddings(self, product_features: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """reference architecture.py:393-407."""
        _, p, _ = self._split_ids(product_features["product_features"])
        mf, mlp = self._gather_ln(1, p)
        cat = self.category_hierarchy(product_features["category_features"]["department_ids"],
                                      product_features["category_features"]["category_ids"])
        return {"mf": mf, "mlp": mlp, "category": cat}
