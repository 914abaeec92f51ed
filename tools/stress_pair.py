#!/usr/bin/env python3
"""Two NCFTrainEngines on one GPU (argument "1,0": one with the auxiliary stream, one serial), stepped alternately on the same
batches from the same state: their tables must agree to 1e-7 after every step (fp32, no dropout, Adam eps 1e-4, so that
rounding noise is not amplified).
usage: python tools/stress_pair.py [repeats] [aux flags, e.g. 1,0]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import ncf_b200
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    modes = (sys.argv[2] if len(sys.argv) > 2 else "1,0").split(",")
    steps = 3
    dev = torch.device("cuda", 0)
    U, I, Bv = 20011, 5003, 4096
    torch.manual_seed(99)
    tables = [(torch.rand(r, 64) * 2 - 1) * (1.0 / r) ** 0.5 for r in (U, I, U, I)]
    batches = bench.make_batches(U, I, Bv, steps, 555, device=dev)
    if os.environ.get("GATHERED"):      # the batch rank 0 of tools/stress_sharded.py trains its single engines on (world 2)
        b0, b1 = bench.make_batches(U, I, Bv // 2, steps, 555, device=dev), bench.make_batches(U, I, Bv // 2, steps, 556, device=dev)
        batches = [tuple(torch.cat([x, y]) for x, y in zip(a, b)) for a, b in zip(b0, b1)]
    perturb = float(os.environ.get("PERTURB", "0"))      # relative noise on the second engine's dense weights
    base = bench.build_model(1, 1, dev, "fp32")
    keys = ("mf_embedding_collection.embedding_bags.user_id.weight", "mf_embedding_collection.embedding_bags.product_id.weight",
            "mlp_embedding_collection.embedding_bags.user_id.weight", "mlp_embedding_collection.embedding_bags.product_id.weight")
    bad = 0
    for rep in range(repeats):
        engs = []
        for aux in modes:
            os.environ["NCF_AUX_STREAM"] = aux
            m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
            sd = m.state_dict()
            for k, v in base.state_dict().items():
                if "embedding_collection" not in k:
                    sd[k] = v.detach().cpu().clone()
            for k, t in zip(keys, tables):
                sd[k] = t.clone()
            if perturb and len(engs) == 1:
                g = torch.Generator().manual_seed(rep)
                for k in sd:
                    if "embedding_collection" not in k and sd[k].dtype == torch.float32:
                        sd[k] = sd[k] * (1.0 + perturb * torch.randn(sd[k].shape, generator=g))
            m.load_state_dict(sd)
            m = m.to(dev).train()
            engs.append(ncf_b200.NCFTrainEngine(m, lr=1e-3, eps=1e-4, weight_decay=1e-5, table_mode="fused_dense_equiv"))
        msgs = []
        for s in range(steps):
            for e in engs:
                e.train_step(*batches[s])
            torch.cuda.synchronize()
            for k in range(4):
                d = (engs[0].model._table_params()[k].detach() - engs[1].model._table_params()[k].detach()).abs().max(dim=1).values
                rows = torch.nonzero(d > float(os.environ.get('THR', '1e-7'))).flatten()
                if rows.numel():
                    msgs.append(f"step {s} table {k}: {rows.numel()} rows max {float(d.max()):.1e} first {rows[:6].tolist()}")
        if msgs:
            bad += 1
            print(f"repeat {rep}: " + "; ".join(msgs[:4]), flush=True)
        for e in engs:
            e.close()
    print(f"stress_pair[{modes}]: {bad} of {repeats} repeats differ")


if __name__ == "__main__":
    main()
