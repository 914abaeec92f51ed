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

import numpy as np
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


def plan_offsets(counts, rank: int):
    """Host arithmetic of the one-sided step, from the all-gathered counts [world(requester), 2 * world + 1]
    (= [user counts per owner | item counts per owner | rows of the batch] of every rank):
      begin[side][o]     first of THIS rank's distinct ids (ascending = owner-major order) that owner o holds
      push_off[side][o]  row of owner o's receive buffer where this rank's segment starts (requesters in rank order)
      n_dist[side]       distinct ids this rank pulls / pushes;  n_recv[side]  rows this rank receives as an owner
      global_rows        sample rows of the whole step (the loss is the mean over them)
    Pure integer code (bit-exact on any host; tests/test_sharding.py checks it against a brute-force layout)."""
    W = len(counts)
    c = [[[int(counts[r][side * W + o]) for o in range(W)] for side in (0, 1)] for r in range(W)]
    begin = [[0] * (W + 1) for _ in (0, 1)]
    push_off = [[0] * W for _ in (0, 1)]
    n_dist, n_recv = [0, 0], [0, 0]
    for side in (0, 1):
        acc = 0
        for o in range(W):
            begin[side][o] = acc
            acc += c[rank][side][o]
            push_off[side][o] = sum(c[r][side][o] for r in range(rank))
        begin[side][W] = acc
        n_dist[side] = acc
        n_recv[side] = sum(c[r][side][rank] for r in range(W))
    return begin, push_off, n_dist, n_recv, sum(int(counts[r][2 * W]) for r in range(W))


def plan_offsets_np(counts, rank: int):
    """plan_offsets as array arithmetic (what the step runs; counts = int64 ndarray): begin int64 [2, world + 1], push_off
    int64 [2, world], n_dist / n_recv lists of two ints, global_rows int."""
    W = counts.shape[0]
    c = counts[:, :2 * W].reshape(W, 2, W)                     # [requester, side, owner]
    begin = np.zeros((2, W + 1), dtype=np.int64)
    np.cumsum(c[rank], axis=1, out=begin[:, 1:])
    push_off = c[:rank].sum(axis=0, dtype=np.int64)
    n_recv = c[:, :, rank].sum(axis=0)
    return begin, push_off, [int(begin[0, W]), int(begin[1, W])], [int(n_recv[0]), int(n_recv[1])], int(counts[:, 2 * W].sum())


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
    used); this engine owns the local shard of the four tables and their Adam state.

    exchange = "p2p" (default on a multi-GPU NCCL group): the ONE-SIDED step of include/ncf_b200.h - every rank maps its
    peers' table shards and gradient receive buffers (CUDA IPC over NVLink / NVSwitch); the forward gather pulls remote
    rows itself, the backward segment-sum pushes gradient rows into the owners' buffers, and the only collectives left
    are three small ones that double as the step's barriers (count all-gather, dense all-reduce).
    exchange = "nccl": the all-to-all(v) version (ids -> owner gather -> rows back -> gradient rows)."""

    def __init__(self, model: AdvancedNCF, num_users: int, num_products: int, lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-5, table_mode: str = "fused_dense_equiv", group=None,
                 seed: int = 1234, init_tables: Optional[List[torch.Tensor]] = None, rank: Optional[int] = None,
                 world: Optional[int] = None, exchange: Optional[str] = None):
        self.lib = _lib.load()
        self.model = model
        self.group = group
        self.world = world if world is not None else (dist.get_world_size(group) if dist.is_initialized() else 1)
        self.rank = rank if rank is not None else (dist.get_rank(group) if dist.is_initialized() else 0)
        if self.world > _lib.MAX_WORLD:
            raise ValueError(f"world size {self.world} exceeds NCF_MAX_WORLD = {_lib.MAX_WORLD}")
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
        # a rank whose block is empty (rows = 9, world = 4) keeps one dummy row per table: nothing ever addresses it
        alloc_u, alloc_i = max(self.rows_u, 1), max(self.rows_i, 1)
        if init_tables is not None:        # global tables given (tests / small models): keep this rank's slice
            sl = [slice(self.rank * bu, self.rank * bu + self.rows_u), slice(self.rank * bi, self.rank * bi + self.rows_i)]
            self.w = []
            for k in range(4):
                t = torch.zeros(alloc_u if k % 2 == 0 else alloc_i, 64, device=dev)
                t[:(self.rows_u if k % 2 == 0 else self.rows_i)] = init_tables[k][sl[k & 1]].to(dev)
                self.w.append(t)
        else:                              # torchrec init U(+-sqrt(1/rows)) with the GLOBAL row count, on device
            g = torch.Generator(device=dev).manual_seed(seed + 7919 * self.rank)
            self.w = []
            for k in range(4):
                rows, glob = (alloc_u, num_users) if k % 2 == 0 else (alloc_i, num_products)
                t = torch.empty(rows, 64, device=dev)
                t.uniform_(-(1.0 / glob) ** 0.5, (1.0 / glob) ** 0.5, generator=g)
                self.w.append(t)
        self.m = [torch.zeros_like(t) for t in self.w]
        self.v = [torch.zeros_like(t) for t in self.w]
        self.touched = [torch.zeros(alloc_u, dtype=torch.uint8, device=dev), torch.zeros(alloc_i, dtype=torch.uint8, device=dev)]
        n = model._flat.numel()
        self._dense_and_loss = torch.zeros(n + 1, device=dev)             # [dense gradients | loss]: one all-reduce
        self.dense_grad = self._dense_and_loss[:n]
        self.dense_m = torch.zeros(n, device=dev)
        self.dense_v = torch.zeros(n, device=dev)
        self.loss = torch.zeros(1, device=dev)
        self.routers = [ShardRouter(group), ShardRouter(group)]
        self._prefetched = None
        self._status = _lib.StatusWord()
        if self.world == 1 or not dist.is_initialized():
            for r in self.routers:
                r.world = 1
        if exchange is None:
            exchange = os.environ.get("NCF_SHARD_EXCHANGE") or ("p2p" if (self.world > 1 and dist.is_initialized()) else "nccl")
        if exchange not in ("p2p", "nccl"):
            raise ValueError(f"unknown exchange {exchange!r}")
        self.exchange = exchange
        # ---- one-sided state (set up lazily: needs the batch size) ----
        self._cap = 0                      # rows per side the buffers are sized for
        self._peers = None                 # peers' pointers: {"w": [world][4], "rows": [world][2], "ids": [world][2]}
        self._bufs = None
        self._plans = None
        self._plan_slot = 0
        if self.world > 1:
            # the step's small collectives must not queue behind a whole tower kernel (include/ncf_b200.h ncf_set_sm_reserve)
            _lib.check(self.lib.ncf_set_sm_reserve(int(os.environ.get("NCF_SM_RESERVE", "1"))), "ncf_set_sm_reserve")
        self.profile = bool(os.environ.get("NCF_SHARD_PROFILE"))     # CUDA-event time of every phase (bench.py sets it)
        self._plan_static = None       # peer pointers of the plan (change only when _setup_p2p remaps)
        self._barrier_word = torch.zeros(1, device=dev)
        # auxiliary stream: the item side of the segment sums / owner update next to the user side, and the routing of the
        # next batch next to whatever leaves SMs free (NCF_SHARD_AUX=0 switches it off)
        self._aux = torch.cuda.Stream(device=dev) if os.environ.get("NCF_SHARD_AUX", "1") != "0" else None

    def table(self, k: int) -> torch.Tensor:
        """this rank's REAL rows of table k (without the dummy row of an empty shard)"""
        return self.w[k][:(self.rows_u if k % 2 == 0 else self.rows_i)]

    # ---- small helpers ----------------------------------------------------------------------
    def _tables(self):
        t = _lib.Tables()
        for k in range(4):
            t.w[k], t.m[k], t.v[k] = self.w[k].data_ptr(), self.m[k].data_ptr(), self.v[k].data_ptr()
        t.touched[0], t.touched[1] = self.touched[0].data_ptr(), self.touched[1].data_ptr()
        t.rows_user, t.rows_item = max(self.rows_u, 1), max(self.rows_i, 1)
        t.status = self._status.ptr()
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

    def check_status(self):
        """IndexError if a step saw an id outside [0, num_users) / [0, num_products) (ncf_check_ids)."""
        self._status.raise_if_set("ShardedNCFEngine")

    def _buffers(self, n: int):
        """Per-step device buffers, allocated once for the largest batch seen (no torch.empty inside the step)."""
        if self._bufs is not None and self._bufs["n"] >= n:
            return self._bufs
        dev = self.device
        cfg = self._cfg()
        b = {"n": n,
             "counts": [torch.zeros(2 * self.world + 2, dtype=torch.long, device=dev) for _ in range(2)],      # [users per owner | items per owner | N | loss bits]
             "local": [torch.empty(2, max(n, 1), dtype=torch.long, device=dev) for _ in range(2)],
             "pos": [torch.empty(2, max(n, 1), dtype=torch.long, device=dev) for _ in range(2)],
             "route_ws": [torch.empty(int(self.lib.ncf_shard_route_workspace_bytes(n)), dtype=torch.uint8, device=dev)
                          for _ in range(2)],
             "ws": torch.empty(int(self.lib.ncf_workspace_bytes(n, C.byref(cfg))), dtype=torch.uint8, device=dev),
             "out": torch.empty(max(n, 1), device=dev), "grad_out": torch.empty(max(n, 1), device=dev),
             "rows": [torch.empty(max(n, 1), 128, device=dev) for _ in range(2)],
             "grads": [torch.empty(max(n, 1), 128, device=dev) for _ in range(2)],
             "counts_all": [torch.zeros(self.world, 2 * self.world + 2, dtype=torch.long, device=dev) for _ in range(2)],
             "counts_host": [torch.zeros(self.world, 2 * self.world + 2, dtype=torch.long).pin_memory() for _ in range(2)],
             "slot": 0}
        self._bufs = b
        return b

    # ---- routing (both exchange modes) --------------------------------------------------------------------------
    def _route(self, user_ids, item_ids, slot: int):
        """requester: every DISTINCT id of the batch is exchanged once (one radix sort of both sides): owner-major local
        ids + per-owner counts + the exchanged-row position of every sample.  counts tensor = [user counts | item counts |
        N] (the row count travels with them: the global mean needs every rank's N)."""
        n = user_ids.numel()
        b = self._buffers(n)
        counts, rws = b["counts"][slot], b["route_ws"][slot]
        # ncf_shard_route lays local ids / positions out as [2][N] for THIS batch's N (the buffers may be larger)
        local = b["local"][slot].view(-1)[:2 * max(n, 1)].view(2, max(n, 1))
        pos = b["pos"][slot].view(-1)[:2 * max(n, 1)].view(2, max(n, 1))
        _lib.check(self.lib.ncf_check_ids(_lib.ptr(user_ids), _lib.ptr(item_ids), n, self.U, self.I, None,
                                          C.c_void_p(self._status.ptr()), self._s()), "ncf_check_ids")
        _lib.check(self.lib.ncf_shard_route(_lib.ptr(user_ids), _lib.ptr(item_ids), n, self.U, self.I, self.world,
                                            _lib.ptr(counts), _lib.ptr(local), _lib.ptr(pos), _lib.ptr(rws), rws.numel(),
                                            self._s()), "ncf_shard_route")
        counts[2 * self.world:2 * self.world + 1].fill_(n)
        return dict(N=n, slot=slot, counts=counts, local=local, pos=pos, route_ws=rws,
                    key=(user_ids.data_ptr(), item_ids.data_ptr(), n))

    def _begin_count_gather(self, routed, loss=None):
        """all-gather of every rank's [2 * world + 2] counts (a collective every rank reaches: it also serves as a barrier
        of the one-sided step) + the device -> pinned host copy; nothing waits yet.  loss = this rank's share of the
        CURRENT step's global loss: rides in the last word (fp32 bits), see _early_loss."""
        b = self._bufs
        slot = routed["slot"]
        if loss is not None:
            routed["counts"][2 * self.world + 1:].view(torch.float32)[:1].copy_(loss)
        all_dev, host = b["counts_all"][slot], b["counts_host"][slot]
        if self.world > 1 and dist.is_initialized():
            dist.all_gather_into_tensor(all_dev.view(-1), routed["counts"], group=self.group)
        else:
            all_dev[self.rank].copy_(routed["counts"])
        host.copy_(all_dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        routed["counts_all"], routed["counts_host"], routed["event"] = all_dev, host, ev
        return routed

    # ---- phases of the all-to-all version (the emulated-cluster test drives these one by one) -------------------
    def phase_bucketize(self, user_ids, item_ids):
        self.step += 1
        r = self._route(user_ids, item_ids, 0)
        self.N = r["N"]
        self._routed = r
        self._plan = [(r["counts"][:self.world], r["local"][0], r["pos"][0]),
                      (r["counts"][self.world:2 * self.world], r["local"][1], r["pos"][1])]
        self._route_ws = r["route_ws"]
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

    def phase_forward_backward(self, rows: List[torch.Tensor], targets: torch.Tensor, global_rows: int, push_plan=None):
        """requester: forward, BCELoss (mean over the GLOBAL batch), backward.  Returns the gradient rows (all-to-all
        version), or pushes them into the owners' receive buffers when push_plan (device plan pointer) is given."""
        N = self.N
        cfg = self._cfg()
        b = self._buffers(N)
        ws, nbytes = b["ws"], b["ws"].numel()
        self.outputs = b["out"][:N]
        self.lib.ncf_set_aux_stream(C.c_void_p(self._aux.cuda_stream) if self._aux is not None else None)
        pos_u, pos_i = self._plan[0][2], self._plan[1][2]
        flat = self.model._flat
        _lib.check(self.lib.ncf_shard_forward(C.byref(cfg), _lib.ptr(flat), _lib.ptr(rows[0]), _lib.ptr(rows[1]),
                                              _lib.ptr(pos_u), _lib.ptr(pos_i), N, _lib.ptr(self.outputs), _lib.ptr(ws),
                                              nbytes, self._s()), "ncf_shard_forward")
        grad_out = b["grad_out"][:N]
        _lib.check(self.lib.ncf_bce_loss(_lib.ptr(self.outputs), _lib.ptr(targets), N, _lib.ptr(self.loss),
                                         _lib.ptr(grad_out), self._s()), "ncf_bce_loss")
        scale = float(N) / float(global_rows)            # local mean -> share of the global mean
        grad_out.mul_(scale)
        self.loss.mul_(scale)
        hook = self.__dict__.pop("_mid_hook", None)
        if hook is not None:
            hook()
        if getattr(self, "_want_loss_event", False):
            self._early_loss()
        self.dense_grad.zero_()
        if push_plan is not None:
            prof = self.profile
            if prof:
                e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                e0.record()
            _lib.check(self.lib.ncf_shard_backward_push(C.byref(cfg), _lib.ptr(flat), _lib.ptr(self.dense_grad), _lib.ptr(rows[0]),
                                                        _lib.ptr(rows[1]), _lib.ptr(pos_u), _lib.ptr(pos_i), N, _lib.ptr(grad_out),
                                                        push_plan, _lib.ptr(self._routed["local"]), _lib.ptr(self._route_ws),
                                                        _lib.ptr(ws), nbytes, self._s()), "ncf_shard_backward_push")
            if prof:
                # the towers' share of that call, re-run in place through the stage exports (gradients into a scratch buffer):
                # the rest is the requester's segment sum + push, i.e. embedding-path work
                e1.record()
                scratch = self.__dict__.setdefault("_prof_grad", torch.zeros_like(self.dense_grad))
                _lib.check(self.lib.ncf_mlp_bwd(C.byref(cfg), _lib.ptr(flat), _lib.ptr(scratch), N, _lib.ptr(grad_out), _lib.ptr(ws),
                                                nbytes, self._s()), "ncf_mlp_bwd")
                _lib.check(self.lib.ncf_attn_bwd(C.byref(cfg), _lib.ptr(flat), _lib.ptr(scratch), N, _lib.ptr(ws), nbytes, self._s()),
                           "ncf_attn_bwd")
                e2.record()
                torch.cuda.synchronize()
                p = self.__dict__.setdefault("phase_ms", {})
                p["requester segment sum + push (inside backward)"] = (p.get("requester segment sum + push (inside backward)", 0.0)
                                                                       + max(e0.elapsed_time(e1) - e1.elapsed_time(e2), 0.0))
            return None
        gu = b["grads"][0][:rows[0].shape[0]]            # one gradient row per exchanged row
        gi = b["grads"][1][:rows[1].shape[0]]
        _lib.check(self.lib.ncf_shard_backward(C.byref(cfg), _lib.ptr(flat), _lib.ptr(self.dense_grad), _lib.ptr(rows[0]),
                                               _lib.ptr(rows[1]), _lib.ptr(pos_u), _lib.ptr(pos_i), N, _lib.ptr(grad_out),
                                               _lib.ptr(gu), _lib.ptr(gi), _lib.ptr(self._route_ws), _lib.ptr(ws), nbytes,
                                               self._s()),
                   "ncf_shard_backward")
        return [gu, gi]

    def phase_owner_sort(self, served: List[torch.Tensor]):
        """owner, one-sided step: radix sort of the local ids the requesters wrote next to their pull (the id-only half of
        the update), enqueued on the current stream; phase_owner_update(presorted=True) then starts at the segment sum."""
        tabs = self._tables()
        ws_list = self.__dict__.setdefault("_own_ws", [None, None])
        for side in (1, 0):
            ids = served[side]
            n = ids.numel()
            if not n:
                continue
            nbytes = int(self.lib.ncf_emb_bwd_workspace_bytes(n))
            if ws_list[side] is None or ws_list[side].numel() < nbytes:
                ws_list[side] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            ws = ws_list[side]
            _lib.check(self.lib.ncf_shard_owner_sort(C.byref(tabs), side, _lib.ptr(ids), n, _lib.ptr(ws), nbytes, self._s()),
                       "ncf_shard_owner_sort")

    def phase_owner_update(self, grad_rows: List[torch.Tensor], served: Optional[List[torch.Tensor]] = None,
                           ln_grad: Optional[torch.Tensor] = None, presorted: bool = False):
        """owner: segment-sum the received gradient rows per local id, LN backward, Adam.  The two sides touch different
        tables (their LayerNorm-affine gradients are added atomically), so the item side runs on the auxiliary stream.
        ln_grad: where the gradients of mf_norm / mlp_norm (the first 256 floats of the flat layout - the only dense
        gradients an owner produces) are accumulated; default: the engine's dense gradient buffer."""
        adam = self._adam()
        dense_grad = ln_grad if ln_grad is not None else self.dense_grad
        tabs = self._tables()
        served = served if served is not None else self._served
        main = torch.cuda.current_stream(self.device)
        ws_list = self.__dict__.setdefault("_own_ws", [None, None])
        for side in (1, 0):
            ids, g = served[side], grad_rows[side]
            n = ids.numel()
            if not n:
                continue
            nbytes = int(self.lib.ncf_emb_bwd_workspace_bytes(n))
            if ws_list[side] is None or ws_list[side].numel() < nbytes:
                ws_list[side] = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
            ws = ws_list[side]
            stream = self._aux if (side == 1 and self._aux is not None) else main
            if stream is not main:
                stream.wait_stream(main)
            fn = self.lib.ncf_shard_owner_update_sorted if presorted else self.lib.ncf_shard_owner_update
            with torch.cuda.stream(stream):
                _lib.check(fn(C.byref(adam), C.byref(tabs), _lib.ptr(self.model._flat), _lib.ptr(dense_grad), side, _lib.ptr(ids),
                              n, _lib.ptr(g), _lib.ptr(ws), nbytes, self._s()), "ncf_shard_owner_update")
        if self._aux is not None:
            main.wait_stream(self._aux)
        if self.table_mode == "fused_dense_equiv":
            _lib.check(self.lib.ncf_emb_adam_sweep(C.byref(adam), C.byref(tabs), self._s()), "ncf_emb_adam_sweep")

    def phase_dense_adam(self):
        adam = self._adam()
        flat = self.model._flat
        _lib.check(self.lib.ncf_dense_adam(_lib.ptr(flat), _lib.ptr(self.dense_grad), _lib.ptr(self.dense_m),
                                           _lib.ptr(self.dense_v), flat.numel(), C.byref(adam), self._s()), "ncf_dense_adam")

    # ---- the one-sided step over peer memory ----------------------------------------------------------------------
    def _own_pointers(self):
        return {"w": [t.data_ptr() for t in self.w], "rows": [t.data_ptr() for t in self._recv_rows],
                "ids": [t.data_ptr() for t in self._recv_ids]}

    def _setup_p2p(self, n: int, peers: Optional[List["ShardedNCFEngine"]] = None):
        """Receive buffers for the gradient rows of every requester (+ their local ids) and the peers' pointers.
        peers = the other in-process engines (emulated cluster, one GPU): plain device pointers; otherwise the pointers
        travel as CUDA IPC handles through all_gather_object and are mapped with ncf_ipc_open."""
        if (self._peers is not None or getattr(self, "_peer_engines", None) is not None) and self._cap >= n:
            return
        dev = self.device
        cap = [self.world * min(n, max(self.rows_u, 1)), self.world * min(n, max(self.rows_i, 1))]
        self._recv_rows = [torch.empty(c, 128, device=dev) for c in cap]
        self._recv_ids = [torch.empty(c, dtype=torch.long, device=dev) for c in cap]
        self._cap = n
        if peers is None and (self.world == 1 or not dist.is_initialized()):
            peers = [self]
        if peers is not None:
            self._peer_engines = peers
            return                                   # pointers are read from the peer engines when the plan is filled
        mine = self._own_pointers()
        exported = {}
        for kind, ptrs in mine.items():
            exported[kind] = []
            for p in ptrs:
                h = (C.c_char * 64)()
                off = C.c_int64(0)
                _lib.check(self.lib.ncf_ipc_export(C.c_void_p(p), h, C.byref(off)), "ncf_ipc_export")
                exported[kind].append((bytes(h), int(off.value)))
        gathered = [None] * self.world
        dist.all_gather_object(gathered, exported, group=self.group)
        if self._peers is not None:                   # re-sized buffers: drop the old mappings
            self._close_peers()
        self._peers = {"w": [], "rows": [], "ids": []}
        self._plan_static = None
        for r in range(self.world):
            for kind in ("w", "rows", "ids"):
                if r == self.rank:
                    self._peers[kind].append(list(mine[kind]))
                    continue
                ptrs = []
                for hbytes, off in gathered[r][kind]:
                    out = C.c_void_p()
                    _lib.check(self.lib.ncf_ipc_open(hbytes, off, C.byref(out)), "ncf_ipc_open")
                    ptrs.append(int(out.value))
                self._peers[kind].append(ptrs)
        self._mapped = [p for kind in ("w", "rows", "ids") for r in range(self.world) if r != self.rank for p in self._peers[kind][r]]

    def _close_peers(self):
        for p in getattr(self, "_mapped", []):
            self.lib.ncf_ipc_close(C.c_void_p(p))
        self._mapped = []

    def close(self):
        """Unmap the peers' buffers (CUDA IPC).  Every rank closes its engine at the same point of the program: a peer
        must not be inside a step that still reads this rank's tables."""
        if getattr(self, "_mapped", None):
            torch.cuda.synchronize(self.device)
            self._close_peers()
            self._peers = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def _peer_ptrs(self, r: int):
        if getattr(self, "_peer_engines", None) is not None:
            e = self._peer_engines[r]
            return e._own_pointers()
        return {k: self._peers[k][r] for k in ("w", "rows", "ids")}

    def _fill_plan(self, counts_host: torch.Tensor):
        """counts_host [world(requester), 2 * world + 1]: builds this step's ncf_shard_plan in pinned memory, uploads it on
        the stream, and returns (device plan pointer, (n distinct users, items), (rows received users, items), global N)."""
        W, me = self.world, self.rank
        begin, push_off, n_dist, n_recv, global_rows = plan_offsets_np(counts_host.numpy(), me)
        if self._plans is None:
            self._plans = [(torch.zeros(C.sizeof(_lib.ShardPlan), dtype=torch.uint8).pin_memory(),
                            torch.zeros(C.sizeof(_lib.ShardPlan), dtype=torch.uint8, device=self.device)) for _ in range(4)]
        host, dev = self._plans[self._plan_slot]
        self._plan_slot = (self._plan_slot + 1) % len(self._plans)
        # the struct as 163 int64 words (include/ncf_b200.h ncf_shard_plan): [world | rank] begin[2][17] tab[16][4]
        # push_rows[2][16] push_ids[2][16]; only begin and the two push pointer tables change from step to step
        MW = _lib.MAX_WORLD
        words = host.numpy().view(np.int64)
        static = self._plan_static if getattr(self, "_peer_engines", None) is None else None
        if static is None:
            ptrs = [self._peer_ptrs(o) for o in range(W)]
            static = {"tab": np.array([[p["w"][k] for k in range(4)] for p in ptrs], dtype=np.int64),
                      "rows": np.array([[p["rows"][side] for p in ptrs] for side in (0, 1)], dtype=np.int64),
                      "ids": np.array([[p["ids"][side] for p in ptrs] for side in (0, 1)], dtype=np.int64)}
            self._plan_static = static
        host.numpy().view(np.int32)[:2] = (W, me)
        words[1:1 + 2 * (MW + 1)].reshape(2, MW + 1)[:, :W + 1] = begin
        o_tab = 1 + 2 * (MW + 1)
        words[o_tab:o_tab + 4 * MW].reshape(MW, 4)[:W] = static["tab"]
        o_rows = o_tab + 4 * MW
        words[o_rows:o_rows + 2 * MW].reshape(2, MW)[:, :W] = static["rows"] + push_off * 512
        o_ids = o_rows + 2 * MW
        words[o_ids:o_ids + 2 * MW].reshape(2, MW)[:, :W] = static["ids"] + push_off * 8
        dev.copy_(host, non_blocking=True)
        return C.c_void_p(dev.data_ptr()), n_dist, n_recv, global_rows

    def phase_pull(self, plan_ptr, n_dist):
        """requester: LN'd [n,128] rows of its distinct ids, read from the owners' shards (one kernel per side)."""
        r = self._routed
        rows = []
        for side in (0, 1):
            out = self._bufs["rows"][side][:max(n_dist[side], 1)]
            _lib.check(self.lib.ncf_shard_pull_rows(plan_ptr, _lib.ptr(self.model._flat), side, _lib.ptr(r["local"][side]),
                                                    n_dist[side], _lib.ptr(out), self._s()), "ncf_shard_pull_rows")
            rows.append(out[:n_dist[side]])
        return rows

    def _early_loss(self):
        """train_step_host: the global loss needs the forward only.  Right behind the loss kernel (on the collectives'
        stream, at the same point of the step on every rank) this rank's share is exchanged and copied to pinned memory: the
        host has the step's result while the backward is still running, so preparing the next step (staging, routing,
        ~0.9 ms of host work per step) is off the critical path - read behind the tower-gradient all-reduce instead, every
        step started one host latency late (N = 2: 1.63 vs 1.49 ms per step end to end).  With look-ahead the share rides
        in the count all-gather of the NEXT batch (one collective, one event for the host: this step's loss and the next
        step's split sizes); the barrier that collective used to be becomes a one-word all-reduce (_train_step_p2p)."""
        main = torch.cuda.current_stream(self.device)
        multi = self.world > 1 and dist.is_initialized()
        nxt = self.__dict__.pop("_pending_next", None)
        if multi and self._aux is not None:
            if getattr(self, "_coll", None) is None:
                self._coll = torch.cuda.Stream(device=self.device)
            self._coll.wait_event(main.record_event())
            with torch.cuda.stream(self._coll):
                if nxt is not None:
                    self._coll.wait_stream(self._aux)            # the routing kernels of the next batch
                    self._prefetched = self._begin_count_gather(nxt, loss=self.loss)
                    self._loss_event, self._loss_from_counts = nxt["event"], nxt["counts_host"]
                else:
                    word = self.__dict__.setdefault("_loss_word", torch.zeros(1, device=self.device))
                    word.copy_(self.loss)
                    dist.all_reduce(word, group=self.group)
                    self._loss_host.copy_(word, non_blocking=True)
                    self._loss_event = self._coll.record_event()
        elif not multi:
            self._loss_host.copy_(self.loss, non_blocking=True)
            self._loss_event = main.record_event()

    def _barrier(self):
        """a collective every rank reaches in stream order (used where the step has no other one at that point)"""
        if self.world > 1 and dist.is_initialized():
            dist.all_reduce(self._barrier_word, group=self.group)

    def _adopt(self, routed):
        self._routed = routed
        self.N = routed["N"]
        self._plan = [(routed["counts"][:self.world], routed["local"][0], routed["pos"][0]),
                      (routed["counts"][self.world:2 * self.world], routed["local"][1], routed["pos"][1])]
        self._route_ws = routed["route_ws"]

    def train_step_host(self, user_ids: torch.Tensor, item_ids: torch.Tensor, targets: torch.Tensor, next_batch=None) -> float:
        """End-to-end step from HOST (ideally pinned) buffers (what trainer.py:253-254, 289 do per batch): H2D copies of
        this rank's ids and targets, the step, and the D2H read of the global loss.  next_batch = the host tensors of the
        NEXT call: copied on a copy stream into another staging set while this step computes, and routed a step ahead
        (train_step next_ids).  THREE staging sets: the set the next batch goes into was last read two steps ago, so its
        copy never waits for the step the device is still running - with two sets the copy, and with it the routing of
        the next batch, started 0.2 ms into the step, behind the towers (1.61 vs 1.49 ms per step at N = 2).  The host waits for the loss only, which is
        exchanged right behind the forward (_early_loss): the backward, the owner update and the dense Adam of this step
        overlap the host's work on the next one.  Every rank must drive the same sequence of calls (the exchanges are
        collectives)."""
        N = user_ids.numel()
        dev = self.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_dev_in", None) is None or self._dev_in[0][0].numel() < N:
            self._dev_in = [(torch.empty(N, dtype=torch.long, device=dev), torch.empty(N, dtype=torch.long, device=dev),
                             torch.empty(N, dtype=torch.float32, device=dev)) for _ in range(3)]
            self._stage_slot, self._staged_key, self._staged_event = 0, None, None
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slot_free = [None, None, None]
            self._loss_host = torch.zeros(1).pin_memory()
        cur = self._stage_slot
        nxt_slot = (cur + 1) % 3
        du, di, dt = (b[:N] for b in self._dev_in[cur])
        key = (user_ids.data_ptr(), item_ids.data_ptr(), targets.data_ptr(), N)
        if self._staged_key == key:
            main.wait_event(self._staged_event)
        else:
            du.copy_(user_ids.reshape(-1), non_blocking=True)
            di.copy_(item_ids.reshape(-1), non_blocking=True)
            dt.copy_(targets.reshape(-1), non_blocking=True)
        self._staged_key = None
        self._next_ids_event = None

        def stage_next():
            """copies the next batch into the other staging set on the copy stream; returns its device ids (or None).
            Called by the step once its own first kernels are enqueued (train_step next_ids as a callable)."""
            nu, ni, nt = next_batch
            M = nu.numel()
            if self._slot_free[nxt_slot] is not None:      # the step that last read that staging set may still be running
                self._copy_stream.wait_event(self._slot_free[nxt_slot])
            with torch.cuda.stream(self._copy_stream):
                for d, h in zip(self._dev_in[nxt_slot], (nu, ni, nt)):
                    d[:M].copy_(h.reshape(-1), non_blocking=True)
                self._staged_event = self._copy_stream.record_event()
            self._staged_key = (nu.data_ptr(), ni.data_ptr(), nt.data_ptr(), M)
            self._next_ids_event = self._staged_event
            return self._dev_in[nxt_slot][0][:M], self._dev_in[nxt_slot][1][:M]
        look = next_batch is not None and next_batch[0].numel() <= self._dev_in[nxt_slot][0].numel()
        self._want_loss_event = True
        try:
            loss = self.train_step(du, di, dt, next_ids=stage_next if look else None)
        finally:
            self._want_loss_event = False
            self._next_ids_event = None
        self._slot_free[cur] = main.record_event()
        self._stage_slot = nxt_slot
        ev = self.__dict__.pop("_loss_event", None)
        shares = self.__dict__.pop("_loss_from_counts", None)
        if ev is not None:
            if getattr(self, "_debug_no_loss_wait", False):      # tools/host_profile_sharded.py: is the wait the bottleneck?
                return float("nan")
            ev.synchronize()
            if shares is not None:      # every rank's share, summed in rank order (fp32, the same bits on every rank)
                value = float(shares[:, 2 * self.world + 1:].contiguous().view(torch.float32)[:, 0].sum(dtype=torch.float32))
            else:
                value = float(self._loss_host[0])
        else:
            value = float(loss.item())
        self.check_status()
        return value

    def _train_step_p2p(self, user_ids, item_ids, targets, next_ids):
        mark = self._mark
        mark(None)
        self._buffers(user_ids.numel())
        self._setup_p2p(self._bufs["n"])
        pre, self._prefetched = self._prefetched, None
        key = (user_ids.data_ptr(), item_ids.data_ptr(), user_ids.numel())
        if pre is None or pre["key"] != key:
            pre = self._begin_count_gather(self._route(user_ids, item_ids, self._bufs["slot"]))
            self._bufs["slot"] ^= 1
        self.step += 1
        self._adopt(pre)
        main = torch.cuda.current_stream(self.device)
        step_start = main.record_event() if next_ids is not None else None      # the previous step is complete here
        pre["event"].synchronize()                   # the split sizes: already there when the batch was routed a step ahead
        plan_ptr, n_dist, n_recv, global_rows = self._fill_plan(pre["counts_host"])
        mark("route + plan")
        rows = self.phase_pull(plan_ptr, n_dist)
        mark("pull rows (P2P)")
        # The pull kernels have written the ids this rank will push rows for into the owners' receive buffers.  Behind a
        # one-word all-reduce (every rank's ids have landed) each owner sorts what it received on the auxiliary stream, next
        # to the towers: the update behind the push starts at the segment sum (6 small kernels per side off the critical path).
        served = [self._recv_ids[0][:n_recv[0]], self._recv_ids[1][:n_recv[1]]]
        # Opt-in (NCF_SHARD_EARLY_SORT=1): measured at N = 2 it does not pay - 1.364 / 1.417 ms against 1.370 / 1.372 ms per step;
        # the extra collective and the sort's kernels compete with the routing of the next batch for the few free SMs.
        early_sort = self._aux is not None and os.environ.get("NCF_SHARD_EARLY_SORT", "0") == "1"
        ev_sorted = None
        if early_sort:
            if getattr(self, "_coll", None) is None:
                self._coll = torch.cuda.Stream(device=self.device)
            self._coll.wait_event(main.record_event())
            with torch.cuda.stream(self._coll):
                self._barrier()
                ev_ids = self._coll.record_event()
            self._aux.wait_event(ev_ids)
            with torch.cuda.stream(self._aux):
                self.phase_owner_sort(served)
                ev_sorted = self._aux.record_event()
        nxt_box = [None]

        def route_next():
            # the next batch's routing (sort, de-duplication) depends on its ids only: on the auxiliary stream it fills the
            # SMs this step's kernels leave idle (kernel boundaries, tails of the persistent tower kernels).  It is enqueued
            # BEFORE the towers: behind them it starts late on the device - they own every SM's shared memory - and the step
            # is 6 % slower (N = 2: 1.576 vs 1.484 ms); but AFTER the pull, so that a host that comes straight from waiting
            # for the previous loss (train_step_host) reaches this step's first kernel 0.2 ms sooner.
            ids = next_ids() if callable(next_ids) else next_ids      # train_step_host: stages the next batch on its copy stream here
            rs = self._aux if self._aux is not None else main
            if rs is not main:
                rs.wait_event(step_start)
            if getattr(self, "_next_ids_event", None) is not None:
                rs.wait_event(self._next_ids_event)          # the next ids arrive on a copy stream (train_step_host)
            with torch.cuda.stream(rs):
                nxt_box[0] = self._route(ids[0], ids[1], self._bufs["slot"])
            self._bufs["slot"] ^= 1
            self._pending_next = nxt_box[0] if getattr(self, "_want_loss_event", False) else None
        self._pending_next = None
        if next_ids is not None:
            # enqueued between the forward and the backward (NCF_SHARD_ROUTE_AT=pre: in front of the forward): the same device time
            # (1.366 vs 1.365 ms at N = 2), but a host that comes from waiting for the previous loss reaches this step's forward
            # sooner (end to end 1.375 vs 1.40 ms)
            if os.environ.get("NCF_SHARD_ROUTE_AT", "mid") == "mid":
                self._mid_hook = route_next
            else:
                route_next()
        self.phase_forward_backward(rows, targets, global_rows, push_plan=plan_ptr)
        self._pending_next = None
        nxt = nxt_box[0]
        mark("forward+backward+push (P2P)")
        # Collectives of the step, all issued on their own stream in the same order on every rank:
        #   A  count all-gather of the NEXT batch (look-ahead; else a one-word all-reduce): a barrier - every requester's
        #      pushed rows have landed before an owner reads its receive buffer;
        #   D  all-reduce of the tower gradients + the loss (336 KB): needs nothing from the owners, so it runs NEXT TO the
        #      owner update instead of after it;
        #   B  all-reduce of the 256 LayerNorm-affine gradients the owners produce: the barrier after which every shard is
        #      updated (the next pull may read it) and every receive buffer consumed (the next push may overwrite it).
        multi = self.world > 1 and dist.is_initialized()
        if self._aux is not None and getattr(self, "_coll", None) is None:
            self._coll = torch.cuda.Stream(device=self.device)
        coll = self._coll if self._aux is not None else main
        if coll is not main:
            coll.wait_stream(main)                   # after this rank's push and its tower gradients
            coll.wait_stream(self._aux)              # ... and after the routing kernels of the next batch
        with torch.cuda.stream(coll):
            if nxt is not None and self._prefetched is not nxt:
                self._prefetched = self._begin_count_gather(nxt)
            else:
                self._barrier()                      # (the counts travelled with the loss already: _early_loss)
            ev_a = coll.record_event() if coll is not main else None
            if multi:
                self._dense_and_loss[-1:].copy_(self.loss)
                dist.all_reduce(self._dense_and_loss, group=self.group)
            if getattr(self, "_want_loss_event", False) and getattr(self, "_loss_event", None) is None:
                # train_step_host without the early read (_early_loss): the global loss is final here
                self._loss_host.copy_(self._dense_and_loss[-1:] if multi else self.loss, non_blocking=True)
                self._loss_event = coll.record_event() if coll is not main else main.record_event()
        if ev_a is not None:
            main.wait_event(ev_a)
        mark("route next (barrier)")
        ln = self.__dict__.setdefault("_ln_grad", torch.zeros(256, device=self.device))
        ln.zero_()
        if ev_sorted is not None:
            main.wait_event(ev_sorted)
        self.phase_owner_update([self._recv_rows[0][:n_recv[0]], self._recv_rows[1][:n_recv[1]]], served=served, ln_grad=ln,
                                presorted=ev_sorted is not None)
        mark("owner update")
        if multi:
            dist.all_reduce(ln, group=self.group)
        if coll is not main:
            main.wait_stream(coll)
        self.dense_grad[:256].add_(ln)
        if multi:
            self.loss.copy_(self._dense_and_loss[-1:])
        self.phase_dense_adam()
        mark("dense allreduce+adam")
        return self.loss

    # ---- the real step ----------------------------------------------------------------------------
    def train_step(self, user_ids: torch.Tensor, item_ids: torch.Tensor, targets: torch.Tensor,
                   next_ids: Optional[Tuple[torch.Tensor, torch.Tensor]] = None) -> torch.Tensor:
        """Global ids int64 [N] and targets fp32 [N] of THIS rank's batch (device tensors; N may differ between ranks).
        Returns the global mean loss (device scalar, identical on all ranks).

        next_ids = (user_ids, item_ids) the NEXT call will be given (input-pipeline look-ahead; all ranks must pass it
        or none): that batch is routed (sort, de-duplication, count exchange) inside this step, so the one host
        synchronisation of a step - reading the counts - finds them already there instead of draining the queue.  It may
        be a callable returning that pair: the step calls it once its own first kernels are enqueued (train_step_host
        stages the next batch there)."""
        self.check_status()
        if self.exchange == "p2p":
            return self._train_step_p2p(user_ids, item_ids, targets, next_ids)
        mark = self._mark
        mark(None)
        pre, self._prefetched = self._prefetched, None
        key = (user_ids.data_ptr(), item_ids.data_ptr(), user_ids.numel())
        if pre is not None and pre["key"] == key:
            self.step += 1
            self._adopt(pre)
            mark("bucketize")
        else:
            self.phase_bucketize(user_ids, item_ids)
            pre = self._begin_count_gather(self._routed)
            mark("bucketize")
        pre["event"].synchronize()
        ch = pre["counts_host"]
        W = self.world
        global_rows = int(ch[:, 2 * W].sum())
        for side, r in enumerate(self.routers):
            r.send_counts = ch[self.rank, side * W:(side + 1) * W].tolist()
            r.recv_counts = ch[:, side * W + self.rank].tolist()
        ids_items, served = [], []
        for side, r in enumerate(self.routers):
            ids = self._plan[side][1]
            if r.world == 1:
                served.append(ids[:r.send_counts[0]])
                continue
            out = ids.new_empty(sum(r.recv_counts))
            ids_items.append((out, ids[:sum(r.send_counts)], r.recv_counts, r.send_counts))
            served.append(out)
        if ids_items:
            _alltoallv_many(ids_items, self.group)
        mark("a2a ids")
        rows_out = self.phase_owner_rows(served)
        mark("owner rows")
        rows = ShardRouter.exchange_rows_multi(self.routers, rows_out, to_owner=False)
        mark("a2a rows")
        grads = self.phase_forward_backward(rows, targets, global_rows)
        mark("forward+backward")
        recv = ShardRouter.exchange_rows_multi(self.routers, grads, to_owner=True)
        mark("a2a grads")
        if next_ids is not None:
            if callable(next_ids):
                next_ids = next_ids()
            keep = (self._routed, self._served)
            if getattr(self, "_next_ids_event", None) is not None:
                torch.cuda.current_stream(self.device).wait_event(self._next_ids_event)
            nxt = self._begin_count_gather(self._route(next_ids[0], next_ids[1], pre["slot"] ^ 1))
            self._prefetched = nxt
            self._routed, self._served = keep
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
        if not self.profile:
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
                    "w": [self.table(k).cpu() for k in range(4)], "m": [self.m[k][:self.table(k).shape[0]].cpu() for k in range(4)],
                    "v": [self.v[k][:self.table(k).shape[0]].cpu() for k in range(4)]},
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
                dst1 = dst0 + self.table(k).shape[0]
                lo, hi = max(src0, dst0), min(src1, dst1)
                if lo >= hi:
                    continue
                for mine, theirs in ((self.w, sh["w"]), (self.m, sh["m"]), (self.v, sh["v"])):
                    mine[k][lo - dst0:hi - dst0].copy_(theirs[k][lo - src0:hi - src0])

    def gather_tables(self) -> List[torch.Tensor]:
        """Reassemble the global tables on every rank (tests / checkpointing of small models)."""
        if self.world == 1:
            return [self.table(k).clone() for k in range(4)]
        out = []
        for k in range(4):
            block = shard_block(self.U if k % 2 == 0 else self.I, self.world)
            pad = torch.zeros(block, 64, device=self.device)
            pad[:self.table(k).shape[0]] = self.table(k)
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
        # large catalogues: operand images of the tensor-core pre-filter, as CatalogueScorer builds them
        self.img = None
        if e.I >= self.TC_MIN_ITEMS and os.environ.get("NCF_SCORE_TC", "1") != "0":
            self.img = torch.empty(int(self.lib.ncf_item_image_bytes(e.I)), dtype=torch.uint8, device=dev)
            _lib.check(self.lib.ncf_item_image(_lib.ptr(self.p_hat), _lib.ptr(self.g), e.I, _lib.ptr(self.img), e._s()),
                       "ncf_item_image")

    TC_MIN_ITEMS = 1 << 16
    TC_MIN_USERS = 64

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
        tabs = e._tables()
        if self.img is not None and n >= self.TC_MIN_USERS:      # same bits as the exact kernel (tests/test_gpu_parity.py)
            nbytes = int(self.lib.ncf_score_topk_tc_workspace_bytes(n, e.I, k_eff))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
            _lib.check(self.lib.ncf_score_topk_tc(C.byref(tabs), _lib.ptr(e.model._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                                  _lib.ptr(self.img), _lib.ptr(u), n, e.I, k_eff, _lib.ptr(idx), _lib.ptr(sc),
                                                  _lib.ptr(ws), nbytes, e._s()), "ncf_score_topk_tc")
            return idx, sc
        nbytes = int(self.lib.ncf_score_topk_workspace_bytes(n, e.I, k_eff))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=e.device)
        _lib.check(self.lib.ncf_score_topk(C.byref(tabs), _lib.ptr(e.model._flat), _lib.ptr(self.p_hat), _lib.ptr(self.g),
                                           _lib.ptr(u), n, e.I, k_eff, _lib.ptr(idx), _lib.ptr(sc), _lib.ptr(ws), nbytes,
                                           e._s()), "ncf_score_topk")
        return idx, sc
