"""Row-sharded path: the all-to-all routing on CPU with gloo (world_size 2), the bit-exact shard
arithmetic, and (GPU) an emulated 2-rank cluster against the single-GPU engine."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ncf_oracle as O
from tests.helpers import golden_params


def test_shard_arithmetic_bit_exact():
    from ncf_b200.sharding import shard_block, shard_owner_local, shard_rows
    g = torch.Generator().manual_seed(0)
    for rows, world in ((100, 8), (138493, 8), (26744, 4), (7, 4), (100_000_000, 8)):
        ids = torch.randint(0, rows, (1000,), generator=g)
        owner, local = shard_owner_local(ids, rows, world)
        ro, rl, block = O.row_shard(ids, rows, world)
        assert block == shard_block(rows, world)
        assert torch.equal(owner, ro) and torch.equal(local, rl)
        assert sum(shard_rows(rows, world, r) for r in range(world)) == rows
        assert int(owner.max()) < world


def test_plan_offsets_of_the_one_sided_step_match_a_brute_force_layout():
    """The host arithmetic behind the pull / push kernels: every requester's segment lands exactly once, in requester order,
    inside every owner's receive buffer, and the row counts add up (also with ragged batches and empty segments)."""
    import numpy as np
    from ncf_b200.sharding import plan_offsets, plan_offsets_np
    g = torch.Generator().manual_seed(0)
    for W in (1, 2, 3, 8):
        counts = torch.randint(0, 50, (W, 2 * W + 1), generator=g)
        counts[torch.rand(W, 2 * W + 1, generator=g) < 0.2] = 0
        counts = counts.tolist()
        for side in (0, 1):
            for o in range(W):                      # owner o's receive buffer: rank 0's rows, then rank 1's, ...
                layout, at = [], 0
                for r in range(W):
                    begin, push_off, n_dist, n_recv, total = plan_offsets(counts, r)
                    n = counts[r][side * W + o]
                    assert push_off[side][o] == at and begin[side][o + 1] - begin[side][o] == n
                    assert begin[side][0] == 0 and begin[side][W] == n_dist[side] == sum(counts[r][side * W:(side + 1) * W])
                    layout += [r] * n
                    at += n
                    assert total == sum(c[2 * W] for c in counts)
                assert plan_offsets(counts, o)[3][side] == len(layout)
        for r in range(W):                          # the array form the step runs is the same arithmetic
            a, b = plan_offsets(counts, r), plan_offsets_np(np.asarray(counts, dtype=np.int64), r)
            assert [row[:W + 1] for row in a[0]] == b[0].tolist() and a[1] == b[1].tolist() and a[2:] == b[2:]


def _router_worker(rank, world, init_file, rows, ret):
    from ncf_b200.sharding import ShardRouter, shard_block, shard_owner_local
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        block = shard_block(rows, world)
        table = torch.arange(rows, dtype=torch.float32).unsqueeze(1) * torch.tensor([1.0, -2.0])   # global [rows,2]
        my_table = table[rank * block:(rank + 1) * block]
        n = 57 + 13 * rank
        ids = torch.randint(0, rows, (n,), generator=g)
        owner, local = shard_owner_local(ids, rows, world)
        order = torch.argsort(owner, stable=True)              # what ncf_shard_bucketize produces on the GPU
        counts = torch.bincount(owner, minlength=world)
        router = ShardRouter()
        served = router.exchange_ids(local[order], counts)
        assert served.numel() == sum(router.recv_counts)
        got = router.return_rows(my_table[served])             # owner lookup, rows routed back
        pos = torch.empty_like(order)
        pos[order] = torch.arange(n)
        assert torch.equal(got[pos], table[ids])               # every sample received ITS row
        # gradient direction: every requester sends ones; owners segment-sum per local id
        recv = router.send_rows(torch.ones(n, 1))
        acc = torch.zeros(my_table.shape[0], 1).index_add_(0, served, recv)
        all_ids = [torch.empty(57 + 13 * r, dtype=torch.long) for r in range(world)]
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            all_ids[r] = torch.randint(0, rows, (57 + 13 * r,), generator=gr)
        want = torch.bincount(torch.cat(all_ids), minlength=rows)[rank * block:(rank + 1) * block].float()
        assert torch.equal(acc[:, 0], want)
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_router_all_to_all_gloo_world2():
    world = 2
    with tempfile.TemporaryDirectory() as d:
        ret = mp.get_context("spawn").Manager().dict()
        mp.spawn(_router_worker, args=(world, os.path.join(d, "rdzv"), 101, ret), nprocs=world, join=True)
        assert dict(ret) == {0: 1, 1: 1}


# ------------------------------------------------------------------------------------------------
def _emulated_exchange(per_rank_bufs, per_rank_counts, world):
    """all-to-all(v) between `world` in-process ranks: rank r's buffer is split by counts[r][o]."""
    out = []
    for o in range(world):
        parts = []
        for r in range(world):
            c = per_rank_counts[r]
            start = sum(c[:o])
            parts.append(per_rank_bufs[r][start:start + c[o]])
        out.append(torch.cat(parts))
    return out


def _p2p_emulated_step(engines, batches):
    """One step of the ONE-SIDED exchange with `world` in-process ranks on one GPU: the peers' pointers are plain device
    pointers, the count all-gather is a torch.stack, and the stream order of the phases stands in for the barriers."""
    world = len(engines)
    n = max(b[0].numel() for b in batches)
    for e in engines:
        e._buffers(n)
        e._setup_p2p(n, peers=engines)
    for e, b in zip(engines, batches):
        e.step += 1
        e._adopt(e._route(b[0], b[1], 0))
    counts = torch.stack([e._routed["counts"] for e in engines]).cpu()
    plans = [e._fill_plan(counts) for e in engines]
    rows = [e.phase_pull(pl[0], pl[1]) for e, pl in zip(engines, plans)]
    for e, b, pl, r in zip(engines, batches, plans, rows):
        assert e.phase_forward_backward(r, b[2], pl[3], push_plan=pl[0]) is None
    for e, pl in zip(engines, plans):
        n_recv = pl[2]
        e.phase_owner_update([e._recv_rows[0][:n_recv[0]], e._recv_rows[1][:n_recv[1]]],
                             served=[e._recv_ids[0][:n_recv[0]], e._recv_ids[1][:n_recv[1]]])
    assert plans[0][3] == sum(b[0].numel() for b in batches)


@pytest.mark.gpu
@pytest.mark.parametrize("world,precision,exchange", [(1, "fp32", "nccl"), (2, "fp32", "nccl"), (3, "fp32", "nccl"), (2, "bf16", "nccl"),
                                                      (1, "fp32", "p2p"), (2, "fp32", "p2p"), (3, "fp32", "p2p"), (4, "fp32", "p2p"), (5, "fp32", "p2p"),
                                                      (2, "bf16", "p2p")])
def test_emulated_cluster_matches_single_gpu_engine(world, precision, exchange):
    """`world` ShardedNCFEngine ranks driven phase by phase in one process must reproduce NCFTrainEngine on the
    concatenated batch: same loss, same weights.  exchange = "nccl": the all-to-alls emulated by copies; "p2p": the
    one-sided pull / push kernels on the other ranks' buffers (plain pointers on one GPU), with ragged batches (the ranks
    hold different numbers of rows) and, at world = 4, a 9-row item table whose last block is EMPTY (torchrec ROW_WISE
    handles empty shards)."""
    import ncf_b200
    from ncf_b200.sharding import ShardedNCFEngine
    U, I, B = 211, (9 if world == 4 else 97), 40
    pg, _ = golden_params()
    g = torch.Generator().manual_seed(world)
    p = {k: v.clone() for k, v in pg.items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * (1.0 / rows) ** 0.5

    def fresh():
        m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
        m.load_state_dict(p)
        m.compute_precision = precision
        return m.cuda().train()

    ref_model = fresh()
    ref = ncf_b200.NCFTrainEngine(ref_model, table_mode="fused_dense_equiv")
    models = [fresh() for _ in range(world)]
    tabs = [p[k] for k in O.TABLE_KEYS]
    engines = [ShardedNCFEngine(models[r], U, I, table_mode="fused_dense_equiv", init_tables=tabs, rank=r, world=world,
                                exchange=exchange) for r in range(world)]
    for step in range(3):
        batches = []
        for r in range(world):
            Br = B + (3 * r if exchange == "p2p" else 0)            # ragged: the global mean needs every rank's row count
            u = torch.randint(0, U, (Br,), generator=g).repeat_interleave(5)
            i = torch.randint(0, I, (Br * 5,), generator=g)
            t = torch.zeros(Br, 5)
            t[:, 0] = 1
            batches.append((u.cuda(), i.cuda(), t.reshape(-1).cuda()))
        ref_loss = ref.train_step(torch.cat([b[0] for b in batches]), torch.cat([b[1] for b in batches]),
                                  torch.cat([b[2] for b in batches])).clone()
        if exchange == "p2p":
            _p2p_emulated_step(engines, batches)
            dsum = sum(e.dense_grad for e in engines)
            lsum = sum(e.loss for e in engines)
            for e in engines:
                e.dense_grad.copy_(dsum)
                e.phase_dense_adam()
            assert abs(float(lsum) - float(ref_loss)) < (2e-6 if precision == "fp32" else 2e-3)
            continue
        plans = [engines[r].phase_bucketize(batches[r][0], batches[r][1]) for r in range(world)]
        served = [[None, None] for _ in range(world)]
        counts = [[plans[r][s][1].tolist() for r in range(world)] for s in (0, 1)]
        for s in (0, 1):
            ex = _emulated_exchange([plans[r][s][0] for r in range(world)], counts[s], world)
            for o in range(world):
                served[o][s] = ex[o]
        rows_out = [engines[o].phase_owner_rows(served[o]) for o in range(world)]
        rows = [[None, None] for _ in range(world)]
        for s in (0, 1):
            # reverse direction: owner o holds segments ordered by requester r with sizes counts[s][r][o]
            rev_counts = [[counts[s][r][o] for r in range(world)] for o in range(world)]
            back = _emulated_exchange([rows_out[o][s] for o in range(world)], rev_counts, world)
            for r in range(world):
                rows[r][s] = back[r]
        grads = [engines[r].phase_forward_backward(rows[r], batches[r][2], world * B * 5) for r in range(world)]
        recv = [[None, None] for _ in range(world)]
        for s in (0, 1):
            ex = _emulated_exchange([grads[r][s] for r in range(world)], counts[s], world)
            for o in range(world):
                recv[o][s] = ex[o]
        for o in range(world):
            engines[o].phase_owner_update(recv[o])
        dsum = sum(e.dense_grad for e in engines)
        lsum = sum(e.loss for e in engines)
        for e in engines:
            e.dense_grad.copy_(dsum)
            e.phase_dense_adam()
        assert abs(float(lsum) - float(ref_loss)) < (2e-6 if precision == "fp32" else 2e-3)
    block_u, block_i = (U + world - 1) // world, (I + world - 1) // world
    ref_tabs = ref_model._table_params()
    for k in range(4):
        full = torch.cat([e.table(k) for e in engines])
        d = (full - ref_tabs[k].detach()).abs()
        assert full.shape == ref_tabs[k].shape
        if precision == "fp32":
            assert float((d > 5e-6).float().mean()) < 3e-3 and d.max() < 1.5e-3, (k, float(d.max()))
        else:       # bf16 operands: Adam amplifies rounding noise on small gradients; most elements still agree
            assert float((d > 2e-4).float().mean()) < 0.15 and d.max() < 7e-3, (k, float(d.max()))
    if precision != "fp32":
        return
    d = (models[0]._flat - ref_model._flat).abs()
    d[ncf_b200._lib.dense_layout()[0][9][1]:ncf_b200._lib.dense_layout()[0][9][1] + 64] = 0   # k_proj.bias: noise
    assert float((d > 5e-6).float().mean()) < 3e-3 and d.max() < 1.5e-3
    assert all(torch.equal(models[0]._flat, m._flat) for m in models[1:])


@pytest.mark.gpu
def test_sharded_scorer_equals_single_gpu_scorer():
    """Users sharded / item side replicated: a world-size-1 ShardedCatalogueScorer on the engine's tables
    returns exactly the single-GPU CatalogueScorer result (indices bit-exact)."""
    import ncf_b200
    from ncf_b200.sharding import ShardedCatalogueScorer, ShardedNCFEngine
    U, I = 300, 1500
    pg, _ = golden_params()
    g = torch.Generator().manual_seed(8)
    p = {k: v.clone() for k, v in pg.items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * 0.05
    m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
    m.load_state_dict(p)
    m = m.cuda().eval()
    users = torch.arange(0, 40).cuda()
    ref_idx, ref_sc = ncf_b200.CatalogueScorer(m).topk(users, 100)
    eng = ShardedNCFEngine(m, U, I, init_tables=[p[k] for k in O.TABLE_KEYS], rank=0, world=1)
    idx, sc = ShardedCatalogueScorer(eng).topk_local_users(users, 100)
    assert torch.equal(idx, ref_idx) and torch.equal(sc, ref_sc)
    # and against the oracle's stable order on the oracle's own scores
    p_hat, gg = O.item_fold(p)
    um = O.layer_norm(p[O.K_UMF][:40], p["mf_norm.weight"], p["mf_norm.bias"])
    full = torch.sigmoid(um @ p_hat.t() + gg)
    want = O.topk_stable(full, 100)
    agree = (idx.cpu() == want).float().mean()
    assert agree > 0.98          # only near-ties (gap below fp32 resolution of the two summation orders) may swap


@pytest.mark.gpu
def test_sharded_checkpoint_reshards_on_load(tmp_path):
    """Save from 2 ranks (after a training step), load into 1 and 3 ranks: tables, moments, dense state equal."""
    import ncf_b200
    from ncf_b200.sharding import ShardedNCFEngine
    U, I = 101, 37
    pg, _ = golden_params()
    g = torch.Generator().manual_seed(1)
    p = {k: v.clone() for k, v in pg.items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = torch.rand(rows, 64, generator=g) * 0.1
    tabs = [p[k] for k in O.TABLE_KEYS]

    def model():
        m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
        m.load_state_dict(p)
        return m.cuda().train()

    src = [ShardedNCFEngine(model(), U, I, init_tables=tabs, rank=r, world=2) for r in range(2)]
    for e in src:                                   # make the state non-trivial
        for k in range(4):
            e.m[k].uniform_(-1, 1, generator=None)
            e.v[k].uniform_(0, 1)
        e.dense_m.normal_()
        e.step = 17
        e.save_checkpoint(str(tmp_path))
    full = lambda engines, attr, k: torch.cat([getattr(e, attr)[k] for e in engines])
    for world in (1, 3):
        dst = [ShardedNCFEngine(model(), U, I, rank=r, world=world, seed=99) for r in range(world)]
        for e in dst:
            e.load_checkpoint(str(tmp_path))
            assert e.step == 17 and torch.equal(e.dense_m, src[0].dense_m)
        for attr in ("w", "m", "v"):
            for k in range(4):
                assert torch.equal(full(dst, attr, k), full(src, attr, k)), (world, attr, k)


@pytest.mark.gpu
def test_lookahead_routing_does_not_change_results():
    """train_step(next_ids=...) routes the next batch early; the trajectory is identical to routing it on demand."""
    import ncf_b200
    from ncf_b200.sharding import ShardedNCFEngine
    U, I, B = 300, 120, 64
    pg, _ = golden_params()
    g = torch.Generator().manual_seed(3)
    p = {k: v.clone() for k, v in pg.items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * (1.0 / rows) ** 0.5
    batches = []
    for _ in range(4):
        u = torch.randint(0, U, (B,), generator=g).repeat_interleave(5)
        i = torch.randint(0, I, (B * 5,), generator=g)
        t = torch.zeros(B, 5)
        t[:, 0] = 1
        batches.append((u.cuda(), i.cuda(), t.reshape(-1).cuda()))
    out = []
    host = [tuple(x.cpu().pin_memory() for x in b) for b in batches]
    for look, exchange in ((False, "nccl"), (True, "nccl"), (False, "p2p"), (True, "p2p"), ("host", "p2p"), ("host", "nccl"),
                           ("early_sort", "p2p")):
        # "early_sort": the owner's id sort ahead of the push (ncf_shard_owner_sort + ncf_shard_owner_update_sorted)
        os.environ["NCF_SHARD_EARLY_SORT"] = "1" if look == "early_sort" else "0"
        m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
        m.load_state_dict(p)
        m = m.cuda().train()
        eng = ShardedNCFEngine(m, U, I, table_mode="fused_sparse", init_tables=[p[k] for k in O.TABLE_KEYS], rank=0, world=1,
                               exchange=exchange)
        losses = []
        for s, b in enumerate(batches):
            if look == "host":        # host buffers in, loss out; the next batch staged on the copy stream (every other step)
                nb = host[s + 1] if (s + 1 < len(host) and s != 1) else None
                losses.append(eng.train_step_host(*host[s], next_batch=nb))
                continue
            nxt = batches[s + 1][:2] if (look and s + 1 < len(batches)) else None
            losses.append(float(eng.train_step(*b, next_ids=nxt)))
        torch.cuda.synchronize()
        os.environ.pop("NCF_SHARD_EARLY_SORT", None)
        out.append((losses, [t.clone() for t in eng.w]))
    for other in out[1:]:
        assert max(abs(a - b) for a, b in zip(out[0][0], other[0])) < 1e-5
        for a, b in zip(out[0][1], other[1]):
            assert float((a - b).abs().mean()) < 1e-6
