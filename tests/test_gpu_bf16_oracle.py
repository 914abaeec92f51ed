"""The BENCHMARKED path - bf16 tcgen05 towers in TRAIN mode (fused S = 5 attention `attn_tc_fwd/bwd`, `mlp_tc_fwd2/bwd/
wgrad`) - checked directly against the oracle (oracle/ncf_oracle.py = the reference's arithmetic in fp32), not against
this repo's own fp32 kernels.

Dropout: the kernels draw Philox masks; `ncf_dropout_mask` dumps exactly those masks and the oracle applies them, so
both sides drop the same elements and the comparison isolates the arithmetic.

Tolerances (BASELINE.json north_star: "bf16 attention/MLP <= 2e-2 relative on logits"):
  LOGIT_RTOL  max |logit - logit_ref| / max |logit_ref| <= 2e-2                       (forward)
  gradients   every tensor: max |g - g_ref| / max |g_ref| <= GRAD_RTOL = 0.10 and cosine(g, g_ref) >= 0.999
              (bf16 rounds every GEMM operand to 8 bits: ~0.4 % per operand, compounding over the 5-GEMM chain of
              the backward; k_proj.bias is excluded: its gradient is pure rounding noise)
  trained     AUC within 0.01 and HR@10 within 0.02 of the fp32 oracle trained on the same batches and masks
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ncf_oracle as O
from tests.helpers import golden_params

pytestmark = pytest.mark.gpu

LOGIT_RTOL = 2e-2
GRAD_RTOL = 0.10
GRAD_COS = 0.999
S = 5


def _logit(p):
    p = p.double().clamp(1e-300, 1 - 1e-16)
    return torch.log(p) - torch.log1p(-p)


def _dump_masks(cfg, N):
    """keep masks of the four dropout sites for the forward that ran with `cfg` (device bool tensors)."""
    from ncf_b200 import _lib
    lib = _lib.load()
    B = N // S
    shapes = {"attn": (0, (B, 4, S, S)), "mlp0": (1, (N, 256)), "mlp1": (2, (N, 128)), "mlp2": (3, (N, 64))}
    masks = {}
    for name, (site, shape) in shapes.items():
        n = int(np.prod(shape))
        buf = torch.empty(n, dtype=torch.uint8, device="cuda")
        _lib.check(lib.ncf_dropout_mask(C.byref(cfg), site, n, _lib.ptr(buf), None))
        masks[name] = buf.bool().view(shape)
    return masks


def _params(U, I, seed, dense="golden"):
    """oracle parameter dict: trained dense weights of the shipped checkpoint (or torch default init) + random tables."""
    g = torch.Generator().manual_seed(seed)
    if dense == "golden":
        pg, _ = golden_params()
        p = {k: v.clone() for k, v in pg.items()}
    else:
        import ncf_b200
        torch.manual_seed(seed)
        p = {k: v.detach().clone() for k, v in ncf_b200.AdvancedNCF(8, 8, 5, 24).state_dict().items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * (1.0 / rows) ** 0.5
    return p, g


def _batch(U, I, B, g):
    u = torch.randint(0, U, (B,), generator=g).repeat_interleave(S)
    w = 1.0 / torch.arange(1, I + 1).float()
    perm = torch.randperm(I, generator=g)
    pos = perm[torch.multinomial(w, B, replacement=True, generator=g)]
    i = torch.randint(0, I, (B, S), generator=g)
    i[:, 0] = pos
    t = torch.zeros(B, S)
    t[:, 0] = 1
    return u, i.reshape(-1), t.reshape(-1, 1)


# The gradient that reaches the USER rows of the MLP tower (and q_proj) flows only through the attention scores:
# d score = p * (dp - sum_k p_k dp_k), a difference of nearly equal numbers while the attention is still close to
# uniform (a freshly initialised model), so the 0.4 % operand rounding of bf16 is amplified there; with the trained
# weights of the shipped checkpoint it stays inside the common bound.
Q_PATH = ("mlp_embedding_collection.embedding_bags.user_id.weight", "user_product_attention.q_proj.weight",
          "user_product_attention.q_proj.bias", "user_product_attention.k_proj.weight")
Q_PATH_RTOL, Q_PATH_COS = 0.2, 0.99
FRESH_COS = 0.995        # every other tensor of the freshly initialised model (its gradients are smaller and noisier)


def _check_grads(named, leaves, what, fresh_init=False):
    rows, bad = [], []
    for k, leaf in leaves.items():
        if k.endswith("k_proj.bias"):
            continue
        a, b = leaf.grad.detach().double().reshape(-1), named[k].grad.detach().double().reshape(-1)
        rel = float((a - b).abs().max() / a.abs().max().clamp_min(1e-30))
        cos = float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))
        rtol, cmin = (Q_PATH_RTOL, Q_PATH_COS) if (fresh_init and k in Q_PATH) else (GRAD_RTOL, FRESH_COS if fresh_init else GRAD_COS)
        rows.append((rel, cos, k))
        if not (rel <= rtol and cos >= cmin):
            bad.append((k, rel, cos))
    rows.sort(reverse=True)
    print(f"gradients {what}: " + "; ".join(f"{k.split('.')[-3] if k.count('.') > 2 else k}:{r:.3f}/{c:.4f}" for r, c, k in rows[:8]))
    assert not bad, (what, bad)
    return rows[0]


@pytest.mark.parametrize("U,I,B,dense,dropout", [(6040, 3706, 4096, "golden", 0.2),       # config[1] shape
                                                 (6040, 3706, 4096, "init", 0.2),         # bench.py's random-init model
                                                 (6040, 3706, 1001, "golden", 0.0),       # ragged tiles, no dropout
                                                 (138493, 26744, 65536, "golden", 0.2)])   # config[2], full bench size
def test_bf16_train_forward_backward_vs_oracle(U, I, B, dense, dropout):
    """module path: model(kjt) in train mode, BCELoss, loss.backward() with precision bf16 and S = 5 (the fused
    attn_tc_fwd/bwd + mlp_tc_* kernels); reference = the oracle's forward + BCELoss + autograd on the same masks."""
    import ncf_b200
    from ncf_b200 import _lib
    p, g = _params(U, I, 7 + B, dense)
    u, i, t = _batch(U, I, B, g)
    N = B * S
    m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=dropout)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    m = m.cuda().train()
    m.compute_precision = "bf16"
    m._dropout_seed = 4242
    du, di, dt = u.cuda(), i.cuda(), t.cuda()
    out = m(ncf_b200.make_kjt(du, di))
    cfg = m._run_cfg(S, True)
    m._fwd_calls -= 1
    cfg.step = m._fwd_calls                          # the step value the forward above used
    assert cfg.precision == _lib.NCF_BF16_TC and cfg.S == 5
    masks = _dump_masks(cfg, N) if dropout else None
    loss = nn.BCELoss()(out, dt)
    loss.backward()

    keys = list(O.TABLE_KEYS) + list(O.ACTIVE_DENSE_KEYS)
    q = {k: v.cuda() for k, v in p.items()}
    leaves = {k: q[k].clone().requires_grad_(True) for k in keys}
    q.update(leaves)
    ref = O.forward(q, du, di, training=True, dropout_p=dropout, masks=masks)
    lref = O.bce_loss(ref, dt)
    lref.backward()

    lg, lr = _logit(out.detach().reshape(-1)), _logit(ref.detach().reshape(-1))
    rel = float((lg - lr).abs().max() / lr.abs().max())
    assert rel <= LOGIT_RTOL, f"logits: {rel:.4f} > {LOGIT_RTOL}"
    assert abs(float(loss) - float(lref)) <= 2e-3 * max(1.0, abs(float(lref)))
    worst = _check_grads(dict(m.named_parameters()), leaves, (U, I, B, dense), fresh_init=dense == "init")
    print(f"bf16 vs oracle {U}x{I} B={B} {dense} p={dropout}: logit rel {rel:.4f}, worst grad {worst}")


def test_bf16_engine_step_vs_oracle_train_step():
    """ncf_train_step (what bench.py times) for 3 steps in bf16 with dropout vs the oracle's train_step (fp32, dense
    torch-Adam semantics) on the same batches and masks: loss, probabilities and updated tables."""
    import ncf_b200
    U, I, B = 6040, 3706, 4096
    p, g = _params(U, I, 3)
    m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.2)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    m = m.cuda().train()
    m.compute_precision = "bf16"
    m._dropout_seed = 99
    eng = ncf_b200.NCFTrainEngine(m, lr=1e-3, weight_decay=1e-5, table_mode="fused_dense_equiv")
    po = {k: v.cuda() for k, v in p.items()}
    state = {}
    w0 = {k: po[k].clone() for k in O.TABLE_KEYS}
    for step in (1, 2, 3):
        u, i, t = _batch(U, I, B, g)
        du, di, dt = u.cuda(), i.cuda(), t.cuda()
        loss = eng.train_step(du, di, dt.reshape(-1))
        masks = _dump_masks(eng._cfg(), B * S)
        lo, out_o, _ = O.train_step(po, state, step, du, di, dt, dropout_p=0.2, masks=masks)
        assert abs(float(loss) - float(lo)) <= 2e-3, (step, float(loss), float(lo))
        lg, lr = _logit(eng.outputs), _logit(out_o.reshape(-1))
        rel = float((lg - lr).abs().max() / lr.abs().max())
        # step 1 compares arithmetic on identical weights; afterwards the two trajectories have taken their own Adam steps
        # (Adam turns rounding-level gradient differences into lr-sized weight differences), so the bound is looser
        assert rel <= (LOGIT_RTOL if step == 1 else 1.5 * LOGIT_RTOL), (step, rel)
        print(f"engine step {step}: loss {float(loss):.6f} vs oracle {float(lo):.6f}, logit rel {rel:.4f}")
    sd = m.state_dict()
    for k in O.TABLE_KEYS:
        # Adam turns every gradient into a step of about lr, so the tables are compared as displacement vectors:
        # the update direction must agree (an occasional sign flip of a near-zero gradient moves one element by 2 lr)
        a, b = (sd[k] - w0[k]).double().reshape(-1), (po[k] - w0[k]).double().reshape(-1)
        cos = float((a @ b) / (a.norm() * b.norm()))
        assert cos > 0.99, (k, cos)
        assert float((sd[k] - po[k]).abs().max()) <= 2.5e-3 * 3, k


def test_trained_auc_and_hr10_match_the_fp32_oracle():
    """J3 'matched AUC/HR@10': one epoch over a config[1]-shaped Zipf set (where item popularity is learnable) and over
    the config[0] restatement of the repo's own datagen, ours-bf16 (fused engine) vs the oracle-fp32 training on the
    SAME batches (device sampler) and the SAME dropout masks; both evaluated with calculate_metrics(batch_size=users,
    negative_samples=99) (utils/metrics.py:9-108) on one held-out positive + 99 unseen negatives per validation user."""
    import ncf_b200
    from ncf_b200 import synthetic
    from ncf_b200.train import ranking_eval
    results = {}
    for name, U, I, inter in (("c1-zipf", 6040, 3706, synthetic.zipf_interactions(6040, 3706, 120000, seed=1234)),
                              ("c0", 8031, 366, synthetic.c0_interactions(days=45, seed=42))):
        train, val = synthetic.time_split(inter, 10)
        cand = synthetic.eval_candidates(val, train, I, 99, max_users=1500, seed=5)
        p, _ = _params(U, I, 21, dense="init")
        m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.2)
        m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
        m = m.cuda().train()
        m.compute_precision = "bf16"
        m._dropout_seed = 77
        eng = ncf_b200.NCFTrainEngine(m, lr=1e-3, weight_decay=1e-5, table_mode="fused_dense_equiv")
        po = {k: v.cuda() for k, v in p.items()}
        state = {}
        loader = ncf_b200.InteractionSampler(torch.from_numpy(train["user"]), torch.from_numpy(train["item"]), U, I,
                                             negative_samples=4, batch_size=256, seed=42)
        step = 0
        for kjt, tg in loader:
            v = kjt.values()
            n = v.numel() // 2
            step += 1
            eng.train_step(v[:n], v[n:], tg.reshape(-1))
            masks = _dump_masks(eng._cfg(), n)
            O.train_step(po, state, step, v[:n], v[n:], tg, dropout_p=0.2, masks=masks)
        ours = ranking_eval(m, cand)
        # the oracle's trained weights, scored by the oracle's own eval-mode forward
        cu, ci, ct = (torch.as_tensor(x).cuda() for x in cand)
        with torch.no_grad():
            ref_scores = O.forward(po, cu, ci, training=False).reshape(-1)
        ref = ncf_b200.calculate_metrics(ref_scores, ct, [1, 5, 10], batch_size=ct.numel() // 100, negative_samples=99)
        results[name] = (ours, ref, step)
        print(f"{name}: {step} steps; ours auc {ours['auc']:.4f} hr@10 {ours['hit_rate@10']:.4f} ndcg@10 {ours['ndcg@10']:.4f}"
              f" | oracle auc {ref['auc']:.4f} hr@10 {ref['hit_rate@10']:.4f} ndcg@10 {ref['ndcg@10']:.4f}")
        assert abs(ours["auc"] - ref["auc"]) <= 0.01, (name, ours["auc"], ref["auc"])
        assert abs(ours["hit_rate@10"] - ref["hit_rate@10"]) <= 0.02, (name, ours["hit_rate@10"], ref["hit_rate@10"])
    # the Zipf set has a learnable signal: both models must have learnt it (otherwise "matched" would be vacuous)
    assert results["c1-zipf"][1]["auc"] > 0.6 and results["c1-zipf"][0]["auc"] > 0.6
