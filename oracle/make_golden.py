#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE ONLY).

Runs only in the build container, where the reference tree is mounted read-only at
/root/reference; the resulting fixtures are committed so that nothing on the GPU box needs the
reference.  The reference modules are imported unchanged through `oracle/shims/` (torchrec and
google.* stand-ins, see their docstrings).

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Fixtures
  golden_ckpt_compact.npz  the shipped checkpoint src/inference/demo/train_20241225_002713_model/
                           (58 non-table tensors, both 366-row item tables, the user rows the
                           golden predictions touch + users 0..255) and the 1000 rows of
                           src/inference/demo/data/predictions.csv.
  train_step.npz           AdvancedNCF(97, 53, ...) with the shipped dense weights and seeded
                           tables; dropout 0; two iterations of the ModelTrainer.train_epoch body
                           (trainer.py:253-285): outputs, loss, every gradient, weights after 2 Adam
                           steps.
  train_dropout.npz        same model, dropout 0.2 with nn.Dropout instrumented to record its
                           keep masks: outputs, loss, selected gradients.
  forward_simple.npz       forward_simple with and without the hour path (fresh Linear recorded),
                           eval forward, get_user/product_embeddings on the shipped checkpoint.
  topk.npz                 app.py:get_recommendations order (pandas nlargest) for 16 users.
  metrics.npz              utils/metrics.calculate_metrics on seeded score matrices.
  sampler.npz              SheetzDataset negative sampling (np.random seeded), ConsistentBatchSampler, collate.
"""
import io
import json
import os
import sys
import zipfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
REF = os.environ.get("NCF_REFERENCE", "/root/reference")
OUT = os.path.join(REPO, "tests", "golden")
sys.path.insert(0, os.path.join(HERE, "shims"))
sys.path.insert(0, REF)

import pandas as pd  # noqa: E402
import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

from src.model.architecture import AdvancedNCF  # noqa: E402  (reference, unmodified)
from src.utils.metrics import calculate_metrics  # noqa: E402
from torchrec.sparse.jagged_tensor import KeyedJaggedTensor  # noqa: E402  (shim)

TABLES = ("mf_embedding_collection.embedding_bags.user_id.weight",
          "mf_embedding_collection.embedding_bags.product_id.weight",
          "mlp_embedding_collection.embedding_bags.user_id.weight",
          "mlp_embedding_collection.embedding_bags.product_id.weight")


def load_shipped_state_dict():
    """Re-zip the unzipped torch archive and torch.load it."""
    src = os.path.join(REF, "src/inference/demo/train_20241225_002713_model")
    buf = io.BytesIO()
    with zipfile.ZipFile(buf, "w", zipfile.ZIP_STORED) as z:
        for root, _, files in os.walk(src):
            for f in files:
                p = os.path.join(root, f)
                z.write(p, "archive/" + os.path.relpath(p, src))
    buf.seek(0)
    return torch.load(buf, map_location="cpu", weights_only=True)


def kjt(users, items):
    v = torch.cat([users, items])
    return KeyedJaggedTensor.from_lengths_sync(["user_id", "product_id"], v,
                                               torch.ones(v.numel(), dtype=torch.long))


def npd(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def golden_ckpt(sd):
    df = pd.read_csv(os.path.join(REF, "src/inference/demo/data/predictions.csv"))
    keep = np.union1d(np.arange(256), df.user_id.unique()).astype(np.int64)
    out = {}
    for k, v in sd.items():
        a = v.numpy()
        if k in (TABLES[0], TABLES[2]):
            a = a[keep]
        out["sd/" + k] = a
    out["user_rows_kept"] = keep
    out["num_users"] = np.int64(sd[TABLES[0]].shape[0])
    out["pred_user_id"] = df.user_id.values.astype(np.int64)
    out["pred_product_id"] = df.product_id.values.astype(np.int64)
    out["pred_label"] = df.label.values.astype(np.int64)
    out["pred_prediction"] = df.prediction.values.astype(np.float32)
    out["pred_original_product_id"] = df.original_product_id.values.astype("U16")
    m = AdvancedNCF(8031, 366, 5, 24, 64, 64, 32, [256, 128, 64], 4, 0.2, 4)
    m.load_state_dict(sd, strict=True)
    m.eval()
    # the reference, run here, on its own golden rows (batch 32, local_inference.py:120-129)
    preds = []
    with torch.no_grad():
        for i in range(0, len(df), 32):
            b = df.iloc[i:i + 32]
            preds.append(m(kjt(torch.tensor(b.user_id.values), torch.tensor(b.product_id.values))).flatten())
    out["pred_reference_cpu"] = torch.cat(preds).numpy()
    np.savez_compressed(os.path.join(OUT, "golden_ckpt_compact.npz"), **out)
    return m, df


def small_model(sd, dropout, seed):
    torch.manual_seed(seed)
    m = AdvancedNCF(97, 53, 5, 24, 64, 64, 32, [256, 128, 64], 4, dropout, 4)
    dense = {k: v for k, v in sd.items() if k not in TABLES}
    missing, unexpected = m.load_state_dict(dense, strict=False)
    assert set(missing) == set(TABLES) and not unexpected
    return m


def make_batch(seed, B, S, U, I):
    g = torch.Generator().manual_seed(seed)
    users = torch.randint(0, U, (B,), generator=g).repeat_interleave(S)
    items = torch.randint(0, I, (B * S,), generator=g)
    # force some duplicate ids across interactions (exercise the segmented gradient reduce)
    users[S:2 * S] = users[0]
    items[7] = items[1]
    items[11] = items[1]
    targets = torch.zeros(B, S)
    targets[:, 0] = 1.0
    return users, items, targets.reshape(-1, 1)


def train_step_fixture(sd):
    m = small_model(sd, 0.0, 7)
    m.train()
    init = {k: v.clone() for k, v in m.state_dict().items()}
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)    # trainer.py:71-75
    crit = nn.BCELoss()                                                   # trainer.py:78
    out = {}
    for k in TABLES:
        out["init/" + k] = init[k].numpy()
    for step in (1, 2):
        users, items, targets = make_batch(100 + step, 6, 5, 97, 53)
        o = m(kjt(users, items))                                          # trainer.py:258
        loss = crit(o, targets)                                           # trainer.py:271
        opt.zero_grad()
        loss.backward()                                                   # trainer.py:276
        out[f"s{step}/users"] = users.numpy()
        out[f"s{step}/items"] = items.numpy()
        out[f"s{step}/targets"] = targets.numpy()
        out[f"s{step}/outputs"] = o.detach().numpy()
        out[f"s{step}/loss"] = loss.detach().numpy()
        if step == 1:
            for k, p in m.named_parameters():
                if p.grad is not None:
                    out["s1/grad/" + k] = p.grad.detach().numpy().copy()
            out["s1/nograd"] = np.array([k for k, p in m.named_parameters() if p.grad is None])
        opt.step()                                                        # trainer.py:285
    for k, v in m.state_dict().items():
        if k in TABLES or not torch.equal(v, init[k]):
            out["after2/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, "train_step.npz"), **out)


class RecordingDropout(nn.Dropout):
    """nn.Dropout with the same semantics whose keep masks are recorded in call order."""
    log = []

    def forward(self, x):
        if not self.training or self.p == 0.0:
            return x
        keep = torch.rand_like(x) >= self.p
        RecordingDropout.log.append(keep)
        return x * keep.to(x.dtype) / (1.0 - self.p)


def train_dropout_fixture(sd):
    orig = nn.Dropout
    nn.Dropout = RecordingDropout          # instrument torch, not the reference
    try:
        m = small_model(sd, 0.2, 11)
    finally:
        nn.Dropout = orig
    m.train()
    torch.manual_seed(5)
    RecordingDropout.log = []
    users, items, targets = make_batch(211, 6, 5, 97, 53)
    o = m(kjt(users, items))
    loss = nn.BCELoss()(o, targets)
    loss.backward()
    masks = RecordingDropout.log
    assert [tuple(x.shape) for x in masks] == [(6, 4, 5, 5), (30, 256), (30, 128), (30, 64)]
    out = {"users": users.numpy(), "items": items.numpy(), "targets": targets.numpy(),
           "outputs": o.detach().numpy(), "loss": loss.detach().numpy(),
           "mask/attn": masks[0].numpy(), "mask/mlp0": masks[1].numpy(),
           "mask/mlp1": masks[2].numpy(), "mask/mlp2": masks[3].numpy()}
    for k in TABLES:
        out["init/" + k] = m.state_dict()[k].numpy()
    for k in TABLES + ("mlp.0.weight", "mlp.10.weight", "user_product_attention.q_proj.weight",
                       "user_product_attention.out_proj.bias", "mlp_norm.weight", "final.0.weight"):
        out["grad/" + k] = dict(m.named_parameters())[k].grad.numpy()
    np.savez_compressed(os.path.join(OUT, "train_dropout.npz"), **out)


def forward_simple_fixture(m):
    g = torch.Generator().manual_seed(3)
    keep = np.load(os.path.join(OUT, "golden_ckpt_compact.npz"))["user_rows_kept"]
    users = torch.from_numpy(keep)[torch.randint(0, len(keep), (64,), generator=g)]
    items = torch.randint(0, 366, (64,), generator=g)
    hour = torch.randint(0, 24, (64,), generator=g)
    out = {"users": users.numpy(), "items": items.numpy(), "hour": hour.numpy()}
    with torch.no_grad():
        out["no_hour"] = m.forward_simple(users, items).numpy()
        out["eval_forward"] = m(kjt(users, items)).numpy()
        torch.manual_seed(99)
        lin = nn.Linear(32, 64)              # the draw forward_simple makes first (architecture.py:437-441)
        out["temporal_proj_weight"] = lin.weight.detach().numpy()
        out["temporal_proj_bias"] = lin.bias.detach().numpy()
        torch.manual_seed(99)
        out["with_hour"] = m.forward_simple(users, items, hour).numpy()
        ue = m.get_user_embeddings({"user_features": kjt(users, items)})
        out["user_emb_mf"] = ue["mf"].numpy()
        out["user_emb_mlp"] = ue["mlp"].numpy()
        dept = torch.randint(0, 5, (64,), generator=g)
        cat = torch.randint(0, 24, (64,), generator=g)
        pe = m.get_product_embeddings({"product_features": kjt(users, items),
                                       "category_features": {"department_ids": dept, "category_ids": cat}})
        out["dept"] = dept.numpy()
        out["cat"] = cat.numpy()
        out["prod_emb_mf"] = pe["mf"].numpy()
        out["prod_emb_mlp"] = pe["mlp"].numpy()
        out["prod_emb_category"] = pe["category"].numpy()
        hh = torch.arange(24)
        te = m.temporal_encoding(hh, hh % 7, hh % 12, hh * 37 + 400)
        out["temporal_encoding"] = te.numpy()
    np.savez_compressed(os.path.join(OUT, "forward_simple.npz"), **out)


def topk_fixture(m):
    keep = np.load(os.path.join(OUT, "golden_ckpt_compact.npz"))["user_rows_kept"]
    users = keep[::max(1, len(keep) // 16)][:16]
    products_df = pd.DataFrame({"product_id": [f"P{i:08X}" for i in range(366)]})
    scores, order = [], []
    for u in users.tolist():
        allp = torch.arange(len(products_df)) % 366                        # app.py:48
        cust = torch.full((len(allp),), u)                                 # app.py:49
        with torch.no_grad():
            s = m.forward_simple(cust, allp)                               # app.py:67
        rec = products_df.copy()
        rec["score"] = s.numpy()
        rec = rec.nlargest(10, "score")                                    # app.py:75
        scores.append(s.numpy())
        order.append(rec.index.values.astype(np.int64))
    np.savez_compressed(os.path.join(OUT, "topk.npz"), users=users, scores=np.stack(scores),
                        top10=np.stack(order))


def metrics_fixture():
    g = torch.Generator().manual_seed(17)
    res = {}
    arrays = {}
    for name, (users, neg) in {"a": (64, 99), "b": (200, 4), "c": (33, 19)}.items():
        P = torch.rand(users, 1 + neg, generator=g)
        P = (P * 50).round() / 50 if name == "c" else P          # 'c' has score ties
        T = torch.zeros(users, 1 + neg)
        T[torch.arange(users), torch.randint(0, 1 + neg, (users,), generator=g)] = 1.0
        if name == "b":
            T[::7, 0] = 1.0                                      # some rows with two positives
        r = calculate_metrics(P.reshape(-1, 1), T.reshape(-1, 1), [1, 5, 10], users, neg)
        res[name] = {k: float(v) for k, v in r.items()}
        arrays[f"{name}/pred"] = P.numpy()
        arrays[f"{name}/target"] = T.numpy()
        arrays[f"{name}/neg"] = np.int64(neg)
    arrays["expected_json"] = np.array(json.dumps(res))
    np.savez_compressed(os.path.join(OUT, "metrics.npz"), **arrays)


def sampler_fixture():
    """SheetzDataset (train mode) on small synthetic frames: negative sampling with np.random.seed, the padded
    batch sampler and collate_recommender_batch, all from the unmodified reference."""
    from src.model.data_prep import ConsistentBatchSampler, SheetzDataset, collate_recommender_batch
    rng = np.random.RandomState(3)
    U, I, M = 20, 15, 240
    users = [f"{6011000000000000 + u}" for u in range(U)]
    prods = [f"P{(i * 2654435761) % (1 << 32):08X}" for i in range(I)]
    ts = pd.Timestamp("2024-01-01") + pd.to_timedelta(rng.randint(0, 90 * 24 * 3600, M), unit="s")
    pop = rng.zipf(1.5, M) % I
    inter = pd.DataFrame({"user_id": [users[u] for u in rng.randint(0, U, M)], "product_id": [prods[i] for i in pop],
                          "amount": rng.rand(M), "transaction_timestamp": ts})
    uf = pd.DataFrame({"cardnumber": users, "recent_interactions": ["{}"] * U, "preferred_categories": ["[]"] * U})
    pf = pd.DataFrame({"product_id": prods, "total_purchases": [1] * I, "total_revenue": [1.0] * I})
    ds = SheetzDataset(inter, uf, pf, mode="train", validation_days=10, negative_samples=4)
    il = np.array([(u, p) for u, p, _ in ds.interaction_list], dtype=np.int64)
    np.random.seed(123)
    samples = [ds[i] for i in range(40)]
    prod_ids = np.stack([s[0]["product_id"].numpy() for s in samples])
    user_ids = np.stack([s[0]["user_id"].numpy() for s in samples])
    targets = np.stack([s[1].numpy() for s in samples])
    np.random.seed(5)
    batches = np.array(list(ConsistentBatchSampler(23, 8, shuffle=True)), dtype=np.int64)
    kjt_b, tgt_b = collate_recommender_batch(samples[:6])
    np.savez_compressed(os.path.join(OUT, "sampler.npz"), interactions=il, num_users=np.int64(ds.num_users),
                        num_products=np.int64(ds.num_products), weights=ds.product_weights, sample_user_ids=user_ids,
                        sample_product_ids=prod_ids, sample_targets=targets, batches=batches,
                        collate_values=kjt_b.values().numpy(), collate_lengths=kjt_b.lengths().numpy(),
                        collate_targets=tgt_b.numpy())


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    sd = load_shipped_state_dict()
    m, _ = golden_ckpt(sd)
    train_step_fixture(sd)
    train_dropout_fixture(sd)
    forward_simple_fixture(m)
    topk_fixture(m)
    metrics_fixture()
    sampler_fixture()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
