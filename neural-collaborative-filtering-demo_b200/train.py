"""Driver equivalent of the reference's `python -m src.train` (src/train.py:152-250) on B200.

The reference entry point loads two BigQuery feature views, builds `AdvancedNCF` + `ModelTrainer`, trains, and
uploads `state_dict()` to GCS.  BigQuery / GCS are out of scope, so the interactions come
from the seeded restatement of the repo's own data generator (`synthetic.c0_interactions`) or a MovieLens-shaped
generator, and the artefacts stay on the local disk; everything between those two ends is the reference's flow:

    initialize_model (train.py:41-69)  ->  ModelTrainer (:205-209)  ->  data loaders (:214-220)
    ->  trainer.train(train_loader, val_loader, num_epochs) (:225-229)  ->  torch.save(state_dict) (:90)

followed by what config[0] of BASELINE.json adds: ranking metrics on a real `[users, 1+99]` layout
(`calculate_metrics(batch_size=users, negative_samples=99)`, utils/metrics.py:9-108), batch scoring of the
validation pairs (`local_inference.py:120-129`) and full-catalogue top-10 for 100 users (`app.py:43-77`).

    python -m ncf_b200.train --shape c0 --epochs 1 [--engine fused|module] [--precision bf16|fp32]
"""
from __future__ import annotations

import argparse
import json
import logging
import os
import time
from typing import Any, Dict, Optional

import numpy as np
import torch

from . import synthetic
from .architecture import AdvancedNCF
from .data_prep import InteractionSampler
from .kjt import make_kjt
from .metrics import calculate_metrics
from .scoring import CatalogueScorer
from .trainer import ModelTrainer, NCFTrainEngine

# the `model.ncf` block of the reference's config/config.yaml:53-69 (hyper-parameter source of truth)
DEFAULT_NCF_CONFIG: Dict[str, Any] = {
    "embedding_dim": 64, "layers": [256, 128, 64], "dropout": 0.2, "num_heads": 4, "temporal_dim": 32,
    "learning_rate": 0.001, "weight_decay": 1e-5, "batch_size": 256, "epochs": 50, "validation_days": 10,
    "negative_samples": 4, "num_workers": 4, "early_stopping_patience": 5,
}

SHAPES = {
    # name: (users, items, generator)
    "c0": (8031, 366, lambda seed: synthetic.c0_interactions(seed=seed)),
    "c1": (6040, 3706, lambda seed: synthetic.zipf_interactions(6040, 3706, 1000209, seed=seed)),
    "c2": (138493, 26744, lambda seed: synthetic.zipf_interactions(138493, 26744, 20000263, seed=seed)),
}


def load_model_config(path: Optional[str] = None) -> Dict[str, Any]:
    """ConfigLoader.get_model_config()['ncf'] (utils/config.py:53-63) incl. its float/int coercions; without a
    file the reference's shipped values."""
    cfg = dict(DEFAULT_NCF_CONFIG)
    if path:
        import yaml
        with open(path) as f:
            cfg.update(yaml.safe_load(f)["model"]["ncf"])
    cfg["weight_decay"] = float(cfg.get("weight_decay", 1e-5))
    cfg["learning_rate"] = float(cfg.get("learning_rate", 0.001))
    cfg["batch_size"] = int(cfg.get("batch_size", 256))
    cfg["epochs"] = int(cfg.get("epochs", 50))
    return cfg


def initialize_model(num_users: int, num_products: int, model_config: Dict[str, Any], num_departments: int = 5,
                     num_categories: int = 24) -> AdvancedNCF:
    """train.py:41-69 (the counts the reference reads off the feature frames are passed in)."""
    for param in ("embedding_dim", "layers", "num_heads", "dropout"):
        if param not in model_config:
            raise ValueError(f"Missing required model parameter: {param}")
    return AdvancedNCF(num_users=num_users, num_products=num_products, num_departments=num_departments,
                       num_categories=num_categories, mf_embedding_dim=model_config["embedding_dim"],
                       mlp_embedding_dim=model_config["embedding_dim"], temporal_dim=model_config.get("temporal_dim", 32),
                       mlp_hidden_dims=model_config["layers"], num_heads=model_config["num_heads"],
                       dropout=model_config["dropout"], negative_samples=model_config.get("negative_samples", 4))


def build_trainer_config(training_config: Dict[str, Any], num_users: int, num_products: int) -> Dict[str, Any]:
    """train.py:109-150 without the GCP ids."""
    if num_users == 0 or num_products == 0:
        raise ValueError(f"Invalid feature counts: users={num_users}, products={num_products}.")
    return {"num_users": num_users, "num_products": num_products, "batch_size": training_config["batch_size"],
            "learning_rate": training_config["learning_rate"], "weight_decay": training_config.get("weight_decay", 1e-5),
            "epochs": training_config["epochs"], "validation_days": training_config.get("validation_days", 10),
            "embedding_dim": training_config["embedding_dim"], "temporal_dim": training_config.get("temporal_dim", 32),
            "num_heads": training_config["num_heads"], "dropout": training_config["dropout"],
            "negative_samples": training_config.get("negative_samples", 4)}


class ValidationPairs:
    """The reference's validation loader (data_prep.py:214-228 + collate): positives only, S = 1."""

    def __init__(self, users: torch.Tensor, items: torch.Tensor, batch_size: int):
        self.users, self.items, self.batch_size = users, items, batch_size

    def __iter__(self):
        for s in range(0, self.users.numel(), self.batch_size):
            u, i = self.users[s:s + self.batch_size], self.items[s:s + self.batch_size]
            yield make_kjt(u, i), torch.ones(u.numel(), 1, device=u.device)


@torch.no_grad()
def ranking_eval(model: AdvancedNCF, cand, k_values=(1, 5, 10), chunk: int = 1 << 20) -> Dict[str, float]:
    """Scores the `[users, 1+neg]` candidate layout with the eval-mode forward and ranks it with calculate_metrics."""
    dev = next(model.parameters()).device
    u, i, t = (torch.as_tensor(x).to(dev) for x in cand)
    was = model.training
    model.eval()
    out = torch.cat([model(make_kjt(u[s:s + chunk], i[s:s + chunk])).reshape(-1) for s in range(0, u.numel(), chunk)])
    model.train(was)
    M = int((t.numel() // max(1, int(t.sum().item()))))
    return calculate_metrics(out, t, list(k_values), batch_size=t.numel() // M, negative_samples=M - 1)


def run(shape: str = "c0", epochs: int = 1, engine: str = "fused", precision: str = "bf16", seed: int = 42,
        config_path: Optional[str] = None, out_dir: Optional[str] = None, max_interactions: int = 0,
        eval_users: int = 2000, score_users: int = 100, top_k: int = 10) -> Dict[str, Any]:
    """One end-to-end job; returns a summary dict (also printed as one JSON line by main())."""
    cfg = load_model_config(config_path)
    users, items, gen = SHAPES[shape]
    torch.manual_seed(seed)
    np.random.seed(seed)
    inter = gen(seed)
    if max_interactions and inter["user"].size > max_interactions:
        inter = {k: v[:max_interactions] for k, v in inter.items()}
    train, val = synthetic.time_split(inter, cfg["validation_days"])
    tcfg = build_trainer_config(cfg, users, items)
    model = initialize_model(users, items, cfg)
    model.compute_precision = precision
    dev = torch.device("cuda", torch.cuda.current_device())
    t0 = time.perf_counter()
    loader = InteractionSampler(torch.from_numpy(train["user"]), torch.from_numpy(train["item"]), users, items,
                                negative_samples=cfg["negative_samples"], batch_size=cfg["batch_size"], seed=seed, device=dev)
    val_loader = ValidationPairs(torch.from_numpy(val["user"]).to(dev), torch.from_numpy(val["item"]).to(dev), 512)
    history: Dict[str, Any]
    if engine == "module":           # the reference loop body, our module as the drop-in (ModelTrainer.train)
        trainer = ModelTrainer(model, tcfg, num_gpus=1)
        history = trainer.train(loader, val_loader, num_epochs=epochs,
                                early_stopping_patience=cfg.get("early_stopping_patience", 5))
    else:                            # the fused engine: one C call per step
        model = model.to(dev).train()
        eng = NCFTrainEngine(model, lr=tcfg["learning_rate"], weight_decay=tcfg["weight_decay"])
        history = {"train_loss": []}
        for _ in range(epochs):
            tot = torch.zeros(1, device=dev)
            nb = 0
            for kjt, tg in loader:
                v = kjt.values()
                n = v.numel() // 2
                tot += eng.train_step(v[:n], v[n:], tg.reshape(-1))
                nb += 1
            history["train_loss"].append(float(tot.item()) / max(nb, 1))
    train_s = time.perf_counter() - t0
    steps = epochs * len(loader)
    cand = synthetic.eval_candidates(val, train, items, 99, max_users=eval_users, seed=seed + 1)
    metrics = ranking_eval(model, cand)
    # batch scoring of the validation pairs (local_inference.py:120-129) + full-catalogue top-k (app.py:43-77)
    model.eval()
    with torch.no_grad():
        vu, vi = torch.from_numpy(val["user"]).to(dev), torch.from_numpy(val["item"]).to(dev)
        pair_scores = torch.cat([model(make_kjt(vu[s:s + 65536], vi[s:s + 65536])).reshape(-1)
                                 for s in range(0, vu.numel(), 65536)])
        idx, sc = CatalogueScorer(model).topk(torch.arange(min(score_users, users), device=dev), top_k)
    summary = {"shape": shape, "users": users, "items": items, "train_interactions": int(train["user"].size),
               "val_interactions": int(val["user"].size), "epochs": epochs, "steps": steps, "engine": engine,
               "precision": precision, "train_seconds": train_s,
               "train_samples_per_s": steps * cfg["batch_size"] * (1 + cfg["negative_samples"]) / train_s,
               "train_loss": history["train_loss"], "metrics": metrics, "mean_val_pair_score": float(pair_scores.mean()),
               "top1_items_first_users": idx[:5, 0].tolist()}
    if out_dir:                       # train.py:71-107 without the GCS upload
        os.makedirs(out_dir, exist_ok=True)
        torch.save(model.state_dict(), os.path.join(out_dir, f"train_{shape}_model.pt"))
        torch.save(history, os.path.join(out_dir, f"train_{shape}_history.pt"))
    return summary


def main():
    logging.basicConfig(level=logging.INFO, format="%(asctime)s - %(levelname)s - %(message)s")
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--shape", default="c0", choices=sorted(SHAPES))
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--engine", default="fused", choices=["fused", "module"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--config", default=None, help="a config.yaml with the reference's model.ncf block")
    ap.add_argument("--out-dir", default=os.environ.get("AIP_MODEL_DIR") or None)
    ap.add_argument("--max-interactions", type=int, default=0)
    a = ap.parse_args()
    print(json.dumps(run(a.shape, a.epochs, a.engine, a.precision, a.seed, a.config, a.out_dir, a.max_interactions)))


if __name__ == "__main__":
    main()
