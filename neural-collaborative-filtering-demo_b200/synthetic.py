"""Seeded synthetic interaction data of the shapes BASELINE.json names (no BigQuery, no network).

`c0_interactions` restates the reference's own data generator as array code (SURVEY 8d, config[0]):
  * customers: 10,000 generated, ~80 % enrolled -> 8,031 active (scripts `02b_generate_customers.py:88`,
    `loyalty_customer_generator.py:24-29`); a transaction picks one uniformly
    (`transaction_generator.py:223` `random.choice(self.customers)`);
  * items per transaction = clip(Poisson(2.5), 1, 8) (`transaction_generator.py:185-186`);
  * primary product uniform over the catalogue (`:101`); every follow-up is, with probability 0.7, a uniform
    product of one of the primary category's two affinity categories and otherwise uniform (`:105-115`).  The
    affinity table is keyed by three-letter category CODES ('MTO', 'BEV', ...) while the catalogue's
    `category_id`s are 'C0001'-style (`product_generator.py:96-103`), so `category_id[:3]` never matches and
    the branch is dead with the reference's own catalogue: `affinity=None` (default) reproduces that,
    `affinity=` a [num_categories, 2] array switches the branch on;
  * 90 days x 1,000 transactions per day from 2024-01-01 (`02c_generate_transactions.py:76-78`), hour of day
    ~ HOURLY_WEIGHTS (`transaction_generator.py:27-34, 73-80`);
  * split: the last `validation_days` = 10 days are the validation set (`data_prep.py:78-88`).

`zipf_interactions` is the MovieLens-shaped generator of SURVEY 8d C1/C2: users uniform, items Zipf(1.0)
over a seeded permutation, timestamps uniform over 90 days.

Everything is numpy on the host and seeded; the result feeds `InteractionSampler` (device-side negative
sampling) or the reference-style DataLoader of the CPU arm.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

# transaction_generator.py:27-34
HOURLY_WEIGHTS = np.array([0.2, 0.1, 0.1, 0.1, 0.3, 0.8, 1.5, 2.0, 1.8, 1.2, 1.0, 1.5,
                           2.0, 1.5, 1.0, 1.2, 1.8, 2.0, 1.8, 1.5, 1.2, 0.8, 0.5, 0.3])


def c0_interactions(num_users: int = 8031, num_products: int = 366, days: int = 90, tx_per_day: int = 1000,
                    seed: int = 42, product_category: Optional[np.ndarray] = None,
                    affinity: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """-> {"user", "item" (dense ids), "day", "hour", "tx"} one entry per purchased line, in generation order."""
    rng = np.random.default_rng(seed)
    n_tx = days * tx_per_day
    cust = rng.integers(0, num_users, n_tx)
    n_items = np.clip(rng.poisson(2.5, n_tx), 1, 8)
    day = np.repeat(np.arange(days), tx_per_day)
    hour = rng.choice(24, size=n_tx, p=HOURLY_WEIGHTS / HOURLY_WEIGHTS.sum())
    total = int(n_items.sum())
    tx = np.repeat(np.arange(n_tx), n_items)
    first = np.ones(total, dtype=bool)
    first[1:] = tx[1:] != tx[:-1]
    item = rng.integers(0, num_products, total)                   # primary + the "otherwise uniform" follow-ups
    if affinity is not None and product_category is not None:
        # follow-ups: 70 % from one of the two related categories of the PRIMARY product (:105-115)
        primary = item[np.flatnonzero(first)][np.cumsum(first) - 1]
        rel = affinity[product_category[primary], rng.integers(0, 2, total)]
        use = (~first) & (rng.random(total) < 0.7) & (rel >= 0)
        by_cat = [np.flatnonzero(product_category == c) for c in range(int(product_category.max()) + 1)]
        for c, members in enumerate(by_cat):
            sel = np.flatnonzero(use & (rel == c))
            if sel.size and members.size:
                item[sel] = members[rng.integers(0, members.size, sel.size)]
    return {"user": cust[tx].astype(np.int64), "item": item.astype(np.int64), "day": day[tx].astype(np.int64),
            "hour": hour[tx].astype(np.int64), "tx": tx.astype(np.int64)}


def zipf_interactions(num_users: int, num_products: int, n: int, seed: int = 1234, exponent: float = 1.0,
                      days: int = 90) -> Dict[str, np.ndarray]:
    """user ~ uniform, item ~ Zipf(exponent) over a seeded permutation, day ~ uniform."""
    rng = np.random.default_rng(seed)
    perm = rng.permutation(num_products)
    w = 1.0 / np.arange(1, num_products + 1, dtype=np.float64) ** exponent
    cdf = np.cumsum(w / w.sum())
    item = perm[np.minimum(np.searchsorted(cdf, rng.random(n)), num_products - 1)]
    return {"user": rng.integers(0, num_users, n).astype(np.int64), "item": item.astype(np.int64),
            "day": rng.integers(0, days, n).astype(np.int64), "hour": rng.integers(0, 24, n).astype(np.int64),
            "tx": np.arange(n, dtype=np.int64)}


def time_split(inter: Dict[str, np.ndarray], validation_days: int = 10):
    """SheetzDataset's split (data_prep.py:78-88): rows with timestamp >= latest - validation_days are validation.
    Timestamps here are whole days, so `latest - 10 days` keeps the last 10 whole days + the latest day's rows."""
    split = int(inter["day"].max()) - validation_days
    tr = inter["day"] < split
    return ({k: v[tr] for k, v in inter.items()}, {k: v[~tr] for k, v in inter.items()})


def eval_candidates(val: Dict[str, np.ndarray], train: Dict[str, np.ndarray], num_products: int, negatives: int = 99,
                    max_users: int = 0, seed: int = 7):
    """The `[users, 1+negatives]` layout `calculate_metrics(batch_size=users, negative_samples=99)` ranks
    (utils/metrics.py:9-108; SURVEY 3.2: the reference's own validate() is degenerate, so the harness builds it):
    one held-out positive per validation user + `negatives` uniform items the user never touched.
    -> (user_ids [G*(1+neg)], item_ids, targets), group-major, positive first."""
    rng = np.random.default_rng(seed)
    users, first = np.unique(val["user"], return_index=True)
    if max_users and users.size > max_users:
        pick = np.sort(rng.choice(users.size, max_users, replace=False))
        users, first = users[pick], first[pick]
    pos = val["item"][first]
    seen = np.unique(np.concatenate([train["user"], val["user"]]) * num_products
                     + np.concatenate([train["item"], val["item"]]))
    G, M = users.size, 1 + negatives
    items = np.empty((G, M), dtype=np.int64)
    items[:, 0] = pos
    neg = rng.integers(0, num_products, (G, negatives))
    for _ in range(20):                                              # redraw the (rare) clashes
        bad = np.isin(users[:, None] * num_products + neg, seen)
        if not bad.any():
            break
        neg[bad] = rng.integers(0, num_products, int(bad.sum()))
    items[:, 1:] = neg
    targets = np.zeros((G, M), dtype=np.float32)
    targets[:, 0] = 1.0
    return np.repeat(users, M).astype(np.int64), items.reshape(-1), targets.reshape(-1)
