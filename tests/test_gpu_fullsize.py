"""BASELINE-size checks (config[2]: 138,493 x 26,744, 65,536 interactions x 5 rows per step) that do not depend on a
CPU run of the oracle: the oracle's own torch code executed on the device as the reference for one fp32 step, and
size-independent properties of the table update."""
import pytest
import torch
import torch.nn as nn

from oracle import ncf_oracle as O
from tests.helpers import golden_params

pytestmark = pytest.mark.gpu

U, I, B, S = 138493, 26744, 65536, 5


def _params(seed=7):
    pg, _ = golden_params()
    g = torch.Generator().manual_seed(seed)
    p = {k: v.clone() for k, v in pg.items()}
    for k, rows in zip(O.TABLE_KEYS, (U, I, U, I)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * (1.0 / rows) ** 0.5
    return p, g


def _batch(g):
    u = torch.randint(0, U, (B,), generator=g).repeat_interleave(S)
    w = 1.0 / torch.arange(1, I + 1).float()
    perm = torch.randperm(I, generator=g)
    pos = perm[torch.multinomial(w, B, replacement=True, generator=g)]
    i = torch.randint(0, I, (B, S), generator=g)
    i[:, 0] = pos
    t = torch.zeros(B, S)
    t[:, 0] = 1
    return u, i.reshape(-1), t.reshape(-1, 1)


def _rel(a, b):
    a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_full_size_fp32_step_matches_the_oracle_code_run_on_the_device():
    """One training-mode forward + backward at the config[2] shape, fp32 kernels, against oracle/ncf_oracle.py's
    forward + BCELoss + autograd executed by torch ON THE GPU (same code as the CPU oracle, 327,680 rows)."""
    import ncf_b200
    p, g = _params()
    u, i, t = _batch(g)
    m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    m = m.cuda().train()
    out = m(ncf_b200.make_kjt(u.cuda(), i.cuda()))
    loss = nn.BCELoss()(out, t.cuda())
    loss.backward()
    keys = list(O.TABLE_KEYS) + list(O.ACTIVE_DENSE_KEYS)
    q = {k: v.cuda() for k, v in p.items()}
    leaves = {k: q[k].clone().requires_grad_(True) for k in keys}
    q.update(leaves)
    ref = O.forward(q, u.cuda(), i.cuda(), training=True)
    lref = O.bce_loss(ref, t.cuda())
    lref.backward()
    assert _rel(out, ref) < 1e-5
    assert abs(float(loss) - float(lref)) < 2e-6
    named = dict(m.named_parameters())
    for k, leaf in leaves.items():
        if k.endswith("k_proj.bias"):          # pure rounding noise (softmax shift invariance), see DESIGN.md
            continue
        # sums over 327,680 rows in a different order: 1e-4 of the tensor's largest entry
        assert _rel(named[k].grad, leaf.grad) < 1e-4, k
    # rows the batch never touched have exactly zero gradient
    for k, ids in ((O.K_UMF, u), (O.K_PMLP, i)):
        mask = torch.ones(named[k].shape[0], dtype=torch.bool)
        mask[torch.unique(ids)] = False
        assert float(named[k].grad[mask.cuda()].abs().max()) == 0.0


def test_full_size_bf16_engine_step_properties():
    """bf16 tcgen05 engine at the config[2] shape: (1) rows the batch never touched receive exactly the reference's
    dense-Adam step for a zero data gradient (g = wd * w), bit for bit against the formula evaluated by torch;
    (2) touched rows move; (3) the loss is finite and goes down over a few steps on a repeated batch; (4) the
    probabilities stay inside (0, 1)."""
    import ncf_b200
    p, g = _params(11)
    u, i, t = _batch(g)
    m = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.2)
    m.load_state_dict({k: v.clone() for k, v in p.items()}, strict=True)
    m = m.cuda().train()
    m.compute_precision = "bf16"
    lr, b1, b2, eps, wd = 1e-3, 0.9, 0.999, 1e-8, 1e-5
    eng = ncf_b200.NCFTrainEngine(m, lr=lr, betas=(b1, b2), eps=eps, weight_decay=wd, table_mode="fused_dense_equiv")
    w0 = p[O.K_UMF].cuda()
    du, di, dt = u.cuda(), i.cuda(), t.reshape(-1).cuda()
    losses = [float(eng.train_step(du, di, dt))]
    w1 = m.mf_embedding_collection.embedding_bags["user_id"].weight.detach()
    untouched = torch.ones(U, dtype=torch.bool, device="cuda")
    untouched[torch.unique(du)] = False
    # torch.optim.Adam single-tensor order with g = wd * w, m = v = 0, t = 1
    gg = wd * w0
    mm = gg * (1 - b1)
    vv = (1 - b2) * gg * gg
    denom = vv.sqrt() * (1.0 / (1 - b2) ** 0.5) + eps
    expect = w0 - (lr / (1 - b1)) * (mm / denom)
    d = (w1[untouched] - expect[untouched]).abs()
    assert float(d.max()) <= 2e-9, float(d.max())                     # one ulp of |w| ~ 3e-3 is 2.3e-10 .. fma contraction
    assert float((w1[~untouched] - w0[~untouched]).abs().max()) > 1e-4  # touched rows took a data step
    for _ in range(4):
        losses.append(float(eng.train_step(du, di, dt)))
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < losses[0]
    assert float(eng.outputs.min()) > 0.0 and float(eng.outputs.max()) < 1.0
