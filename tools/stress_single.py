#!/usr/bin/env python3
"""Run-to-run reproducibility of NCFTrainEngine (fp32, no dropout, Adam eps = 1e-4 so that rounding noise is not amplified):
the same 3 steps from the same state, repeated; every repeat's tables are compared with the first one's.
usage: python tools/stress_single.py [repeats] [table_mode] [precision]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import ncf_b200
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    mode = sys.argv[2] if len(sys.argv) > 2 else "fused_dense_equiv"
    prec = sys.argv[3] if len(sys.argv) > 3 else "fp32"
    steps = int(os.environ.get("STEPS", "3"))
    dev = torch.device("cuda", 0)
    U, I, Bv = 20011, 5003, 4096
    torch.manual_seed(99)
    tables = [(torch.rand(r, 64) * 2 - 1) * (1.0 / r) ** 0.5 for r in (U, I, U, I)]
    batches = bench.make_batches(U, I, Bv, steps, 555, device=dev)
    base = bench.build_model(1, 1, dev, prec)
    keys = ("mf_embedding_collection.embedding_bags.user_id.weight", "mf_embedding_collection.embedding_bags.product_id.weight",
            "mlp_embedding_collection.embedding_bags.user_id.weight", "mlp_embedding_collection.embedding_bags.product_id.weight")
    first = None
    bad = 0
    for rep in range(repeats):
        m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0, compute_precision=prec) if prec != "fp32" else ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
        sd = m.state_dict()
        for k, v in base.state_dict().items():
            if "embedding_collection" not in k:
                sd[k] = v.detach().cpu().clone()
        for k, t in zip(keys, tables):
            sd[k] = t.clone()
        m.load_state_dict(sd)
        m = m.to(dev).train()
        eng = ncf_b200.NCFTrainEngine(m, lr=1e-3, eps=1e-4, weight_decay=1e-5, table_mode=mode)
        per_step = []
        for s in range(steps):
            eng.train_step(*batches[s])
            per_step.append([p.detach().clone() for p in m._table_params()])
        torch.cuda.synchronize()
        if first is None:
            first = per_step
        else:
            msgs = []
            for s in range(steps):
                for k in range(4):
                    d = (per_step[s][k] - first[s][k]).abs().max(dim=1).values
                    rows = torch.nonzero(d > 1e-7).flatten()
                    if rows.numel():
                        msgs.append(f"step {s} table {k}: {rows.numel()} rows, max {float(d.max()):.1e}, first {rows[:6].tolist()}")
            if msgs:
                bad += 1
                print(f"repeat {rep}: " + "; ".join(msgs[:6]), flush=True)
        eng.close()
    print(f"stress_single[{mode},{prec}]: {bad} of {repeats - 1} repeats differ from the first")


if __name__ == "__main__":
    main()
