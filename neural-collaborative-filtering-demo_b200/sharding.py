"""Row-sharded AdvancedNCF training across the GPUs of one NVSwitch box.

The reference's own multi-GPU path (`DistributedModelParallel(..., device_ids=...)`, trainer.py:84-88)
cannot run; this module defines the sharding it intended, torchrec ROW_WISE style:

    block = ceil(rows / world);  owner(id) = id // block;  local(id) = id % block

Tables (both towers of a side together) live on their owner; the batch is data-parallel.  One step:

    requester  sort + de-duplicate ids, bucket by owner    ncf_shard_route
    ---------  all-to-all(v) ids ------------------------  ShardRouter.exchange_ids
    owner      gather + LayerNorm local rows -> [n,128]    ncf_shard_owner_rows
    ---------  all-to-all(v) rows back ------------------  ShardRouter.return_rows
    requester  GMF + towers forward, BCELoss, backward     ncf_shard_forward / ncf_shard_backward
    ---------  all-to-all(v) gradient rows --------------  ShardRouter.send_rows
    owner      sorted-id segment sum + LN backward + Adam  ncf_shard_owner_update
    ---------  all-reduce dense gradients (336 KB) ------  dist.all_reduce
    all        Adam on the replicated dense parameters     ncf_dense_adam

The loss is the mean over the GLOBAL batch, so a world-size-G run equals a single-GPU run on the
concatenated batch.  `ShardRouter` is device-agnostic torch.distributed plumbing (tested on CPU with
gloo); every compute phase is a libncf_b200 call.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .architecture import AdvancedNCF, _stream


def shard_block(rows: int, world: int) -> int:
    return (rows + world - 1) // world


def shard_owner_local(ids: torch.Tensor, rows: int, world: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(owner rank, local row) of global ids - integer arithmetic, bit-exact on any device."""
    block = shard_block(rows, world)
    return torch.div(ids, block, rounding_mode="floor"), ids % block


def shard_rows(rows: int, world: int, rank: int) -> int:
    block = shard_block(rows, world)
    return max(0, min(rows, (rank + 1) * block) - rank * block)


def _alltoallv_many(items, group) -> None:
    """Several all-to-all(v) exchanges launched back to back (async) and awaited together.
    items = [(out, inp, recv_counts, send_counts)].  (A single batched point-to-point group was tried instead of one
    NCCL call per exchange: slower on NVSwitch.)"""
    works = [dist.all_to_all_single(out, inp, recv_counts, send_counts, group=group, async_op=True)
             for out, inp, recv_counts, send_counts in items]
    for w in works:
        w.wait()


class ShardRouter:
    """all-to-all(v) routing of one side's ids / rows between requesters and owners."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.send_counts: List[int] = []
        self.recv_counts: List[int] = []

    def exchange_ids(self, local_ids_owner_major: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
        """counts[w] ids go to rank w (the first sum(counts) entries of the owner-major id buffer); returns the
        ids this rank must serve, grouped by requesting rank."""
        return ShardRouter.exchange_ids_multi([self], [(local_ids_owner_major, counts)])[0]

    @staticmethod
    def exchange_ids_multi(routers, plans):
        """The id exchange of several sides at once: ONE all-to-all of all the counts and one host sync (the split
        sizes), then the id all-to-alls are launched back to back."""
        return ShardRouter.finish_id_exchange(ShardRouter.begin_id_exchange(routers, plans))

    @staticmethod
    def begin_id_exchange(routers, plans):
        """Enqueue the count exchange and the device->host copy of the split sizes; nothing waits yet, so the caller
        can queue more GPU work before finish_id_exchange blocks the CPU."""
        first = routers[0]
        counts = torch.stack([c for _, c in plans], dim=1).contiguous()              # [world, sides]
        if first.world == 1:
            both = torch.stack([counts, counts])
        else:
            recv = torch.empty_like(counts)
            dist.all_to_all_single(recv, counts, group=first.group)
            both = torch.stack([counts, recv])
        if both.is_cuda:
            host = torch.empty(both.shape, dtype=both.dtype, pin_memory=True)
            host.copy_(both, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        else:
            host, ev = both, None
        return dict(routers=routers, plans=plans, host=host, event=ev)

    @staticmethod
    def finish_id_exchange(h):
        routers, plans = h["routers"], h["plans"]
        if h["event"] is not None:
            h["event"].synchronize()                                                  # the one host sync of the step
        both = h["host"]
        outs, items, counts = [], [], []
        for k, (r, (ids, _)) in enumerate(zip(routers, plans)):
            send, recv = both[0, :, k].tolist(), both[1, :, k].tolist()
            counts.append((send, recv))
            if r.world == 1:
                outs.append(ids[:send[0]])
                continue
            out = ids.new_empty(sum(recv))
            items.append((out, ids[:sum(send)], recv, send))
            outs.append(out)
        if items:
            _alltoallv_many(items, routers[0].group)
        for r, (send, recv) in zip(routers, counts):
            r.send_counts, r.recv_counts = send, recv
        return outs

    def return_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """owner -> requester: rows [sum(recv_counts), W] come back in the order the ids were sent."""
        if self.world == 1:
            return rows
        out = rows.new_empty((sum(self.send_counts),) + tuple(rows.shape[1:]))
        dist.all_to_all_single(out, rows, self.send_counts, self.recv_counts, group=self.group)
        return out

    @staticmethod
    def exchange_rows_multi(routers, rows_list, to_owner: bool):
        """return_rows (to_owner=False) / send_rows (to_owner=True) of several sides launched back to back."""
        if routers[0].world == 1:
            return list(rows_list)
        outs, items = [], []
        for r, rows in zip(routers, rows_list):
            out_counts, in_counts = (r.recv_counts, r.send_counts) if to_owner else (r.send_counts, r.recv_counts)
            out = rows.new_empty((sum(out_counts),) + tuple(rows.shape[1:]))
            items.append((out, rows, out_counts, in_counts))
            outs.append(out)
        _alltoallv_many(items, routers[0].group)
        return outs

    def send_rows(self, rows: torch.Tensor) -> torch.Tensor:
        """requester -> owner: rows [sum(send_counts), W] in owner-major order (gradient rows)."""
        if self.world == 1:
            return rows
        out = rows.new_empty((sum(self.recv_counts),) + tuple(rows.shape[1:]))
        dist.all_to_all_single(out, rows, self.recv_counts, self.send_counts, group=self.group)
        return out


class ShardedNCFEngine:
    """One process per GPU.  `model` holds the replicated dense parameters (its own tables are not
    used); this engine owns the local shard of the four tables and their Adam state."""

    def __init__(self, model: AdvancedNCF, num_users: int, num_products: int, lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-5, table_mode: str = "fused_dense_equiv", group=None,
                 seed: int = 1234, init_tables: Optional[List[torch.Tensor]] = None, rank: Optional[int] = None,
                 world: Optional[int] = None):
        self.lib = _lib.load()
        self.model = model
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        self.U, self.I = num_users, num_products
        model._ensure_flat()
        self.device = model._flat.device
        self.S = 1 + model.negative_samples
        self.hp = dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay)
        self.table_mode = table_mode
        self.step = 0
        bu, bi = shard_block(num_users, self.world), shard_block(num_products, self.world)
        self.rows_u, self.rows_i = shard_rows(num_users, self.world, self.rank), shard_rows(num_products, self.world, self.rank)
        dev = self.device
        if init_tables is not None:        # global tables given (tests / small models): keep this rank's slice
            sl = [slice(self.rank * bu, self.rank * bu + self.rows_u), slice(self.rank * bi, self.rank * bi + self.rows_i)]
            self.w = [init_tables[k][sl[k & 1]].to(dev).contiguous().clone() for k in range(4)]
        else:                              # torchrec init U(+-sqrt(1/rows)) with the GLOBAL row count, on device
            g = torch.Generator(device=dev).manual_seed(seed + 7919 * self.rank)
            self.w = []
            for k in range(4):
                rows, glob = (self.rows_u, num_users) if k % 2 == 0 else (self.rows_i, num_products)
                t = torch.empty(max(rows, 1), 64, device=dev)
                t.uniform_(-(1.0 / glob) ** 0.5, (1.0 / glob) ** 0.5, generator=g)
                self.w.append(t[:rows] if rows else t[:0])
        self.m = [torch.zeros_like(t) for t in self.w]
        self.v = [torch.zeros_like(t) for t in self.w]
        self.touched = [torch.zeros(max(self.rows_u, 1), dtype=torch.uint8, device=dev),
                        torch.zeros(max(self.rows_i, 1), dtype=torch.uint8, device=dev)]
        n = model._flat.numel()
        self._dense_and_loss = torch.zeros(n + 1, device=dev)             # [dense gradients | loss]: one all-reduce
        self.dense_grad = self._dense_and_loss[:n]
        self.dense_m = torch.zeros(n, device=dev)
        self.dense_v = torch.zeros(n, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.routers = [ShardRouter(group), ShardRouter(group)]
        self._prefetched = None
        if self.world == 1 or not dist.is_initialized():
            for r in self.routers:
                r.world = 1

    # ---- small helpers ----------------------------------------------------------------------
    def _tables(self):
        t = _lib.Tables()
        for k in range(4):
            t.w[k], t.m[k], t.v[k] = self.w[k].data_ptr(), self.m[k].data_ptr(), self.v[k].data_ptr()
        t.touched[0], t.touched[1] = self.touched[0].data_ptr(), self.touched[1].data_ptr()
        t.rows_user, t.rows_item = max(self.rows_u, 1), max(self.rows_i, 1)
        return t

    def _cfg(self):
        m = self.model
        cfg = _lib.RunCfg()
        cfg.S, cfg.training = self.S, 1
        cfg.precision = _lib.NCF_BF16_TC if m.compute_precision == "bf16" else _lib.NCF_FP32
        cfg.dropout_p = float(m.dropout)
        cfg.seed = m._dropout_seed + self.rank
        cfg.step = self.step
        return cfg

    def _adam(self):
        a = _lib.AdamCfg()
        a.lr, (a.beta1, a.beta2) = self.hp["lr"], self.hp["betas"]
        a.eps, a.weight_decay, a.step = self.hp["eps"], self.hp["weight_decay"], self.step
        a.emb_mode = _lib.EMB_ADAM_DENSE_EQUIV if self.table_mode == "fused_dense_equiv" else _lib.EMB_ADAM_SPARSE
        return a

    def _s(self):
        return _stream(self.device)

    # ---- phases (the emulated-cluster test drives these one by one) -----------------------------
    def phase_bucketize(self, user_ids, item_ids):
        """requester: every DISTINCT id of the batch is exchanged once (one radix sort of both sides): owner-major
        local ids + per-owner counts + the exchanged-row position of every sample."""
        self.step += 1
        self.N = n = user_ids.numel()
        dev = self.device
        counts = torch.empty(2, self.world, dtype=torch.long, device=dev)
        local = torch.empty(2, max(n, 1), dtype=torch.long, device=dev)
        pos = torch.empty(2, max(n, 1), dtype=torch.long, device=dev)
        nbytes = int(self.lib.ncf_shard_route_workspace_bytes(n))
        self._route_ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.check(self.lib.ncf_shard_route(_lib.ptr(user_ids), _lib.ptr(item_ids), n, self.U, self.I, self.world,
                                            _lib.ptr(counts), _lib.ptr(local), _lib.ptr(pos), _lib.ptr(self._route_ws), nbytes,
                                            self._s()), "ncf_shard_route")
        self._plan = [(counts[0], local[0], pos[0]), (counts[1], local[1], pos[1])]
        return [(p[1], p[0]) for p in self._plan]        # [(local ids owner-major, counts)] per side

    def phase_owner_rows(self, served_ids: List[torch.Tensor]):
        """owner: LN'd [n,128] rows for the ids each side has to serve."""
        self._served = served_ids
        out = []
        tabs = self._tables()
        for side, ids in enumerate(served_ids):
            rows = torch.empty(ids.numel(), 128, device=self.device)
            _lib.check(self.lib.ncf_shard_owner_rows(C.byref(tabs), _lib.ptr(self.model._flat), side, _lib.ptr(ids),
                                                     ids.numel(), _lib.ptr(rows), self._s()), "ncf_shard_owner_rows")
            out.append(rows)
        return out

    def phase_forward_backward(self, rows: List[torch.Tensor], targets: torch.Tensor, global_rows: int):
        """requester: forward, BCELoss (mean over the GLOBAL batch), backward; returns the gradient rows."""
        N = self.N
        cfg = self._cfg()
        nbytes = int(self.lib.ncf_workspace_bytes(N, C.byref(cfg)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        self.outputs = torch.empty(N, device=self.device)
        pos_u, pos_i = self._plan[0][2], self._plan[1][2]
        flat = self.model._flat
        _lib.check(self.lib.ncf_shard_forward(C.byref(cfg), _lib.ptr(flat), _lib.ptr(rows[0]), _lib.ptr(rows[1]),
                                              _lib.ptr(pos_u), _lib.ptr(pos_i), N, _lib.ptr(self.outputs), _lib.ptr(ws),
                                              nbytes, self._s()), "ncf_shard_forward")
        grad_out = torch.empty(N, device=self.device)
        _lib.check(self.lib.ncf_bce_loss(_lib.ptr(self.outputs), _lib.ptr(targets), N, _lib.ptr(self.loss),
                                         _lib.ptr(grad_out), self._s()), "ncf_bce_loss")
        scale = float(N) / float(global_rows)            # local mean -> share of the global mean
        grad_out.mul_(scale)
        self.loss.mul_(scale)
        self.dense_grad.zero_()
        gu = torch.empty(rows[0].shape[0], 128, device=self.device)     # one gradient row per exchanged row
        gi = torch.empty(rows[1].shape[0], 128, device=self.device)
        _lib.check(self.lib.ncf_shard_backward(C.byref(cfg), _lib.ptr(flat), _lib.ptr(self.dense_grad), _lib.ptr(rows[0]),
                                               _lib.ptr(rows[1]), _lib.ptr(pos_u), _lib.ptr(pos_i), N, _lib.ptr(grad_out),
                                               _lib.ptr(gu), _lib.ptr(gi), _lib.ptr(self._route_ws), _lib.ptr(ws), nbytes,
                                               self._s()),
                   "ncf_shard_backward")
        return [gu, gi]

    def phase_owner_update(self, grad_rows: List[torch.Tensor]):
        """owner: segment-sum the received gradient rows per local id, LN backward, Adam."""
        adam = self._adam()
        tabs = self._tables()
        for side in (1, 0):
            ids, g = self._served[side], grad_rows[side]
            n = ids.numel()
            if n:
                nbytes = int(self.lib.ncf_emb_bwd_workspace_bytes(n))
                ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
                _lib.check(self.lib.ncf_shard_owner_update(C.byref(adam), C.byref(tabs), _lib.ptr(self.model._flat),
                                                           _lib.ptr(self.dense_grad), side, _lib.ptr(ids), n, _lib.ptr(g),
                                                           _lib.ptr(ws), nbytes, self._s()), "ncf_shard_owner_update")
        if self.table_mode == "fused_dense_equiv":
            _lib.check(self.lib.ncf_emb_adam_sweep(C.byref(adam), C.byref(tabs), self._s()), "ncf_emb_adam_sweep")

    def phase_dense_adam(self):
        adam = self._adam()
        flat = self.model._flat
        _lib.check(self.lib.ncf_dense_adam(_lib.ptr(flat), _lib.ptr(self.dense_grad), _lib.ptr(self.dense_m),
                                           _lib.ptr(self.dense_v), flat.numel(), C.byref(adam), self._s()), "ncf_dense_adam")

    # ---- the real step ----------------------------------------------------------------------------
    def train_step(self, user_ids: torch.Tensor, item_ids: torch.Tensor, targets: torch.Tensor,
                   next_ids: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        """Global ids int64 [N] and targets fp32 [N] of THIS rank's batch (device tensors).  Returns the
        global mean loss (device scalar, identical on all ranks).

        next_ids = (user_ids, item_ids) the NEXT call will be given (input-pipeline look-ahead; all ranks must pass it
        or none): that batch is routed (sort, de-duplication, count exchange) right after this step's gradient exchange
        is queued, so the one host synchronisation of a step - reading the split sizes - waits while the GPU still has
        the owner update of this step to do, instead of draining the queue."""
        mark = self._mark
        mark(None)
        pre, self._prefetched = self._prefetched, None
        key = (user_ids.data_ptr(), item_ids.data_ptr(), user_ids.numel())
        if pre is not None and pre["key"] == key:
            self.step += 1
            self._plan, self.N, self._route_ws = pre["plan"], pre["N"], pre["route_ws"]
            mark("bucketize")
            served = ShardRouter.finish_id_exchange(pre["handle"])
        else:
            plan = self.phase_bucketize(user_ids, item_ids)
            mark("bucketize")
            served = ShardRouter.exchange_ids_multi(self.routers, plan)
        mark("a2a ids")
        rows_out = self.phase_owner_rows(served)
        mark("owner rows")
        rows = ShardRouter.exchange_rows_multi(self.routers, rows_out, to_owner=False)
        mark("a2a rows")
        grads = self.phase_forward_backward(rows, targets, self.N * self.world)
        mark("forward+backward")
        recv = ShardRouter.exchange_rows_multi(self.routers, grads, to_owner=True)
        mark("a2a grads")
        if next_ids is not None:
            keep = (self.step, self._plan, self.N, self._route_ws, self._served)
            plan = self.phase_bucketize(*next_ids)
            handle = ShardRouter.begin_id_exchange(self.routers, plan)
            self._prefetched = dict(key=(next_ids[0].data_ptr(), next_ids[1].data_ptr(), next_ids[0].numel()), plan=self._plan,
                                    N=self.N, route_ws=self._route_ws, handle=handle)
            self.step, self._plan, self.N, self._route_ws, self._served = keep
            mark("route next")
        self.phase_owner_update(recv)
        mark("owner update")
        if self.world > 1:
            self._dense_and_loss[-1:].copy_(self.loss)
            dist.all_reduce(self._dense_and_loss, group=self.group)      # dense gradients + the loss in one call
            self.loss.copy_(self._dense_and_loss[-1:])
        self.phase_dense_adam()
        mark("dense allreduce+adam")
        return self.loss

    # NCF_SHARD_PROFILE=1: CUDA-event time of every phase, summed over the steps (bench.py prints it to stderr)
    def _mark(self, name):
        if not os.environ.get("NCF_SHARD_PROFILE"):
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if name is None:
            self._ev = [(None, ev)]
        else:
            self._ev.append((name, ev))
        if name == "dense allreduce+adam":
            torch.cuda.synchronize()
            prof = self.__dict__.setdefault("phase_ms", {})
            for (_, a), (n, b) in zip(self._ev[:-1], self._ev[1:]):
                prof[n] = prof.get(n, 0.0) + a.elapsed_time(b)
            prof["steps"] = prof.get("steps", 0) + 1

    # ---- sharded checkpoints -------------------------------------------------------
    def save_checkpoint(self, directory: str) -> None:
        """Every rank writes its table shard (weights + Adam moments) to `shard_<rank>_of_<world>.pt`; rank 0 also
        writes `dense.pt` (the reference state_dict keys of the replicated parameters, their Adam state, step)."""
        os.makedirs(directory, exist_ok=True)
        torch.save({"rank": self.rank, "world": self.world, "num_users": self.U, "num_products": self.I,
                    "w": [t.cpu() for t in self.w], "m": [t.cpu() for t in self.m], "v": [t.cpu() for t in self.v]},
                   os.path.join(directory, f"shard_{self.rank}_of_{self.world}.pt"))
        if self.rank == 0:
            dense = {k: v.detach().cpu() for k, v in self.model.state_dict().items() if "embedding_collection" not in k}
            torch.save({"model_state_dict": dense, "dense_m": self.dense_m.cpu(), "dense_v": self.dense_v.cpu(),
                        "step": self.step, "hp": dict(self.hp), "table_mode": self.table_mode, "world": self.world},
                       os.path.join(directory, "dense.pt"))

    def load_checkpoint(self, directory: str) -> None:
        """Resume from `save_checkpoint` output written with ANY world size: each rank reads the saved shards that
        overlap its own row range (re-sharding on load)."""
        d = torch.load(os.path.join(directory, "dense.pt"), map_location="cpu", weights_only=False)
        self.model.load_state_dict(d["model_state_dict"], strict=False)
        self.model._ensure_flat()
        self.dense_m.copy_(d["dense_m"])
        self.dense_v.copy_(d["dense_v"])
        self.step = int(d["step"])
        files = sorted(glob.glob(os.path.join(directory, "shard_*_of_*.pt")))
        if not files:
            raise FileNotFoundError(f"no table shards under {directory}")
        for f in files:
            sh = torch.load(f, map_location="cpu", weights_only=False)
            if sh["num_users"] != self.U or sh["num_products"] != self.I:
                raise ValueError("checkpoint was written for different table sizes")
            for k in range(4):
                rows = self.U if k % 2 == 0 else self.I
                src_block, dst_block = shard_block(rows, sh["world"]), shard_block(rows, self.world)
                src0 = sh["rank"] * src_block
                src1 = src0 + sh["w"][k].shape[0]
                dst0 = self.rank * dst_block
                dst1 = dst0 + self.w[k].shape[0]
                lo, hi = max(src0, dst0), min(src1, dst1)
                if lo >= hi:
                    continue
                for mine, theirs in ((self.w, sh["w"]), (self.m, sh["m"]), (self.v, sh["v"])):
                    mine[k][lo - dst0:hi - dst0].copy_(theirs[k][lo - src0:hi - src0])

    def gather_tables(self) -> List[torch.Tensor]:
        """Reassemble the global tables on every rank (tests / checkpointing of small models)."""
        if self.world == 1:
            return [t.clone() for t in self.w]
        out = []
        for k in range(4):
            block = shard_block(self.U if k % 2 == 0 else self.I, self.world)
            pad = torch.zeros(block, 64, device=self.device)
            pad[:self.w[k].shape[0]] = self.w[k]
            parts = [torch.empty_like(pad) for _ in range(self.world)]
            dist.all_gather(parts, pad, group=self.group)
            out.append(torch.cat(parts)[:self.U if k % 2 == 0 else self.I])
        return out


class ShardedCatalogueScorer:
    """Full-catalogue top-k with the tables of a ShardedNCFEngine (scoring): USERS stay
    sharded, the folded item side (P_hat [I,64], g [I]) is replicated by one all-gather, and every rank
    ranks the users it owns against the whole catalogue - no merge step, so the result does not depend
    on the number of GPUs."""

    def __init__(self, engine: ShardedNCFEngine):
        self.e = engine
        self.lib = engine.lib
        self.p_hat = None
        self.g = None
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        e = self.e
        dev = e.device
        rows = max(e.rows_i, 1)
        p_local = torch.zeros(shard_block(e.I, e.world), 64, device=dev)
        g_local = torch.zeros(shard_block(e.I, e.world), device=dev)
        if e.rows_i:
            nbytes = int(self.lib.ncf_item_fold_workspace_bytes(rows))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            tabs = e._tables()
            _lib.check(self.lib.ncf_item_fold(C.byref(tabs), _lib.ptr(e.model._flat), _lib.ptr(p_local), _lib.ptr(g_local),
                                              _lib.ptr(ws), nbytes, e._s()), "ncf_item_fold")
        if e.world > 1:
            ps = [torch.empty_like(p_local) for _ in range(e.world)]
            gs = [torch.empty_like(g_local) for _ in range(e.world)]
            dist.all_gather(ps, p_local, group=e.group)
            dist.all_gather(gs, g_local, group=e.group)
            self.p_hat = torch.cat(ps)[:e.I].contiguous()
            self.g = torch.cat(gs)[:e.I].contiguous()
        else:
            self.p_hat, self.g = p_local[:e.I].contiguous(), g_local[:e.I].contiguous()

    @torch.no_grad()
    def topk_local_users(self, local_user_ids: torch.Tensor, k: int):
        """(global item indices [n,k], scores [n,k]) for users given by their LOCAL row on this rank."""
        e = self.e
        u = local_user_ids.reshape(-1).to(device=e.device, dtype=torch.long).contiguous()
        n = u.numel()
        k_eff = min(k, e.I)
        idx = torch.empty(n, k_eff, dtype=torch.long, device=e.device)
        sc = torch.empty(n, k_eff, dtype=torch.float32, device=e.device)
        if n == 0:
            return idx, sc
        nbytes = int(self.lib.ncf_score_topk_workspace_bytes(n, e.I, k_eff))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
        tabs = e._tables()
        _lib.check(self.lib.ncf_score_topk(C.byref(tabs), _lib.ptr(e.model._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                           _lib.ptr(u), n, e.I, k_eff, _lib.ptr(idx), _lib.ptr(sc), _lib.ptr(ws), nbytes,
                                           e._s()), "ncf_score_topk")
        return idx, sc
