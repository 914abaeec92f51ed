#!/usr/bin/env python3
"""Shake the sharded step for races: the sharded engine against TWO single-GPU engines on rank 0 (one with the auxiliary
stream, one serial) on the same batches, tables compared after EVERY step, with random host-side delays injected in front
of the phases of ShardedNCFEngine (different on every rank) so that the ranks drift against each other.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/stress_sharded.py [repeats] [max_sleep_ms] [steps]"""
import os
import random
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402

KEYS = ("mf_embedding_collection.embedding_bags.user_id.weight", "mf_embedding_collection.embedding_bags.product_id.weight",
        "mlp_embedding_collection.embedding_bags.user_id.weight", "mlp_embedding_collection.embedding_bags.product_id.weight")
NAMES = ("user_mf", "item_mf", "user_mlp", "item_mlp")


def main():
    import torch
    import torch.distributed as dist
    import ncf_b200
    from ncf_b200.sharding import ShardedNCFEngine
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    max_ms = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    rng = random.Random(1000 + rank)

    def delayed(fn):
        def wrap(*a, **k):
            if max_ms > 0 and rng.random() < 0.5:
                time.sleep(rng.random() * max_ms / 1e3)
            return fn(*a, **k)
        return wrap
    for name in ("_fill_plan", "phase_pull", "phase_forward_backward", "phase_owner_update", "_route", "_begin_count_gather"):
        setattr(ShardedNCFEngine, name, delayed(getattr(ShardedNCFEngine, name)))
    U, I, Bv = 20011, 5003, 2048
    eps = float(os.environ.get("EPS", "1e-4"))      # 1e-4: rounding noise is not amplified by Adam, tables agree to 1e-7
    thr = 1e-7 if eps >= 1e-5 else 1e-4
    bad = 0
    for rep in range(repeats):
        torch.manual_seed(99)
        tables = [(torch.rand(r, 64) * 2 - 1) * (1.0 / r) ** 0.5 for r in (U, I, U, I)]
        model = bench.build_model(1, 1, dev, "fp32")
        model.dropout = 0.0
        eng = ShardedNCFEngine(model, U, I, lr=1e-3, eps=eps, weight_decay=1e-5, table_mode="fused_dense_equiv", init_tables=tables)
        batches = bench.make_batches(U, I, Bv, steps, 555 + rank, device=dev)
        singles = []
        if rank == 0:
            for aux in ("1", "0"):
                os.environ["NCF_AUX_STREAM"] = aux
                ref = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
                sd = ref.state_dict()
                for k, v in model.state_dict().items():
                    if "embedding_collection" not in k:
                        sd[k] = v.detach().cpu().clone()
                for k, t in zip(KEYS, tables):
                    sd[k] = t.clone()
                ref.load_state_dict(sd)
                ref = ref.to(dev).train()
                singles.append(ncf_b200.NCFTrainEngine(ref, lr=1e-3, eps=eps, weight_decay=1e-5, table_mode="fused_dense_equiv"))
            os.environ["NCF_AUX_STREAM"] = "1"
        first_bad = None
        for s in range(steps):
            u, i, t = batches[s]
            eng.train_step(u, i, t)
            parts = [[torch.empty_like(x) for _ in range(world)] for x in (u, i, t)]
            for p, x in zip(parts, (u, i, t)):
                dist.all_gather(p, x)
            got = eng.gather_tables()
            if rank == 0:
                cu, ci, ct = (torch.cat(p) for p in parts)
                for e in singles:
                    e.train_step(cu, ci, ct)
                torch.cuda.synchronize()
                rep_lines = []
                for k in range(4):
                    wa, wb = (e.model._table_params()[k].detach() for e in singles)
                    dsa = (got[k] - wa).abs().max(dim=1).values
                    dsb = (got[k] - wb).abs().max(dim=1).values
                    dab = (wa - wb).abs().max(dim=1).values
                    if max(float(dsa.max()), float(dsb.max()), float(dab.max())) > thr:
                        ids = (cu if k % 2 == 0 else ci)
                        named = torch.zeros(got[k].shape[0], dtype=torch.bool, device=dev)
                        named[ids] = True
                        block = (got[k].shape[0] + world - 1) // world
                        for nm, d in (("sharded-vs-single(aux)", dsa), ("sharded-vs-single(serial)", dsb), ("single(aux)-vs-single(serial)", dab)):
                            rows = torch.nonzero(d > thr).flatten()
                            if rows.numel():
                                rep_lines.append(f"    {NAMES[k]} {nm}: {rows.numel()} rows; first {rows[:8].tolist()} owners "
                                                 f"{(rows[:8] // block).tolist()} named-this-step {named[rows[:8]].tolist()} "
                                                 f"({int(named[rows].sum())} of them named) max {float(d.max()):.2e}")
                if rep_lines and (first_bad is None or os.environ.get("ALL_STEPS")):
                    first_bad = s if first_bad is None else first_bad
                    print(f"repeat {rep}: tables differ after step {s}", flush=True)
                    print("\n".join(rep_lines), flush=True)
        if rank == 0:
            bad += first_bad is not None
            for e in singles:
                e.close()
        del eng, singles
        torch.cuda.empty_cache()
    if rank == 0:
        print(f"stress: {bad} of {repeats} repeats with differing tables")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
