"""Times the UNMODIFIED reference training loop on the host cores (TEST / BASELINE INFRASTRUCTURE ONLY).

SURVEY 8d "CPU baseline beside it": the shimmed reference code itself - `ModelTrainer.train_epoch`
(/root/reference/src/model/trainer.py:216-337) driving `AdvancedNCF.forward` (src/model/architecture.py:258-381),
`nn.BCELoss`, `torch.optim.Adam` - imported through `oracle/shims/` exactly like `oracle/make_golden.py` does, timed
  (i)  model-only: on pre-collated batches (the Python DataLoader excluded), and
  (ii) end to end with the reference's own input pipeline: `SheetzDataset` + `ConsistentBatchSampler` +
       `collate_recommender_batch` (src/model/data_prep.py:13-320, 397-444) in a `DataLoader(num_workers=4)`.
The reference tree exists in the build container only (/root/reference); on the GPU box `available()` is False and
bench.py falls back to the oracle port (`cpu_baseline.kind = "port"`).
"""
import logging
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NCF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "src", "model", "trainer.py"))


def _import_reference():
    for p in (os.path.join(HERE, "shims"), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    logging.disable(logging.INFO)
    os.environ.setdefault("TQDM_DISABLE", "1")
    from src.model.architecture import AdvancedNCF          # noqa: E402  (reference, unmodified)
    from src.model.trainer import ModelTrainer              # noqa: E402
    from src.model import data_prep                         # noqa: E402
    from torchrec.sparse.jagged_tensor import KeyedJaggedTensor   # noqa: E402  (shim)
    return AdvancedNCF, ModelTrainer, data_prep, KeyedJaggedTensor


def _trainer(users, items, batch_size, threads):
    import torch
    AdvancedNCF, ModelTrainer, data_prep, KJT = _import_reference()
    torch.set_num_threads(threads)
    torch.manual_seed(1234)
    model = AdvancedNCF(users, items, 5, 24, 64, 64, 32, [256, 128, 64], 4, 0.2, 4)
    cfg = {"num_users": users, "num_products": items, "batch_size": batch_size, "learning_rate": 1e-3, "weight_decay": 1e-5,
           "project_id": "none", "dataset_id": "none", "negative_samples": 4}
    return ModelTrainer(model, cfg, num_gpus=1), data_prep, KJT


def train_rate(users, items, batches, steps, warmup, threads):
    """(i) model-only steps of the reference trainer on pre-collated (user_ids, item_ids, targets) triples.
    Returns (sample rows per second, ms per step)."""
    import torch
    n = batches[0][0].numel()
    trainer, _, KJT = _trainer(users, items, n // 5, threads)
    loader = []
    for s in range(warmup + steps):
        u, i, t = batches[s % len(batches)]
        values = torch.cat([u, i]).long()
        loader.append((KJT.from_lengths_sync(keys=["user_id", "product_id"], values=values,
                                             lengths=torch.ones(values.numel(), dtype=torch.long)), t.view(-1, 1)))
    if warmup:
        trainer.train_epoch(loader[:warmup])
    t0 = time.perf_counter()
    trainer.train_epoch(loader[warmup:])
    dt = time.perf_counter() - t0
    return n * steps / dt, 1e3 * dt / steps


def loader_rate(users, items, interactions, batch_size, steps, threads, num_workers=4):
    """(ii) the same loop fed by the reference's own Dataset / sampler / collate in a DataLoader: `interactions` is a dict
    of numpy columns user / item / day (ncf_b200.synthetic).  Returns (sample rows per second, ms per step, steps run)."""
    import numpy as np
    import pandas as pd
    from torch.utils.data import DataLoader
    trainer, dp, _ = _trainer(users, items, batch_size, threads)
    cards = np.arange(users)
    prods = np.arange(items)
    inter = pd.DataFrame({"user_id": interactions["user"], "product_id": interactions["item"], "amount": 1.0,
                          "transaction_timestamp": pd.Timestamp("2024-01-01") + pd.to_timedelta(interactions["day"], unit="D")})
    uf = pd.DataFrame({"cardnumber": cards, "recent_interactions": 0, "preferred_categories": 0})
    pf = pd.DataFrame({"product_id": prods, "total_purchases": 0, "total_revenue": 0.0})
    ds = dp.SheetzDataset(inter, uf, pf, mode="train", validation_days=10, negative_samples=4)
    sampler = dp.ConsistentBatchSampler(len(ds), batch_size, shuffle=True)
    loader = DataLoader(ds, batch_sampler=sampler, collate_fn=dp.collate_recommender_batch, num_workers=num_workers)

    class Bounded:
        def __iter__(self):
            for k, b in enumerate(loader):
                if k >= steps:
                    return
                yield b

        def __len__(self):
            return steps
    t0 = time.perf_counter()
    trainer.train_epoch(Bounded())
    dt = time.perf_counter() - t0
    done = min(steps, len(sampler))
    return batch_size * 5 * done / dt, 1e3 * dt / done, done


if __name__ == "__main__":      # python oracle/reference_runner.py : the two timings at the shipped shape
    import json
    import torch
    sys.path.insert(0, os.path.dirname(HERE))
    import importlib.util
    spec = importlib.util.spec_from_file_location("syn", os.path.join(os.path.dirname(HERE), "neural-collaborative-filtering-demo_b200",
                                                                      "synthetic.py"))
    syn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(syn)
    threads = os.cpu_count() or 1
    g = torch.Generator().manual_seed(0)
    batches = []
    for _ in range(4):
        u = torch.randint(0, 8031, (256,), generator=g).repeat_interleave(5)
        i = torch.randint(0, 366, (1280,), generator=g)
        t = torch.zeros(256, 5)
        t[:, 0] = 1
        batches.append((u, i, t.reshape(-1)))
    r1 = train_rate(8031, 366, batches, 20, 3, threads)
    inter = syn.c0_interactions(days=20, tx_per_day=200)
    r2 = loader_rate(8031, 366, inter, 256, 12, threads)
    print(json.dumps({"shape": "config[0] 8,031 x 366, batch 256", "cores": threads,
                      "model_only": {"samples_per_s": r1[0], "ms_per_step": r1[1]},
                      "with_reference_dataloader": {"samples_per_s": r2[0], "ms_per_step": r2[1], "steps": r2[2]}}))
