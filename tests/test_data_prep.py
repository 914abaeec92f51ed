"""Input pipeline: the oracle restatement of the reference sampler/collate is pinned to a
fixture produced by the unmodified reference; the device sampler is checked against the oracle's
distribution and the reference's constraints."""
import numpy as np
import pytest
import torch

from oracle import ncf_oracle as O
from tests.helpers import load_npz


def _history(inter):
    h = {}
    for u, p in inter.tolist():
        h.setdefault(u, set()).add(p)
    return h


def test_oracle_sampler_matches_reference_fixture():
    z = load_npz("sampler.npz")
    inter = z["interactions"]
    I = int(z["num_products"])
    w = O.product_weights(inter[:, 1], I)
    assert np.allclose(w, z["weights"], rtol=0, atol=1e-15)
    hist = _history(inter)
    np.random.seed(123)
    for k in range(40):                                  # same numpy stream as the reference run
        u, p = int(inter[k, 0]), int(inter[k, 1])
        negs = [O.sample_negative(u, p, w, hist, I) for _ in range(4)]
        assert [p] + negs == z["sample_product_ids"][k].tolist()
        assert z["sample_user_ids"][k].tolist() == [u] * 5
        assert z["sample_targets"][k].tolist() == [1, 0, 0, 0, 0]
    np.random.seed(5)
    assert np.array_equal(np.array(list(O.consistent_batches(23, 8))), z["batches"])
    samples = [(z["sample_user_ids"][k], z["sample_product_ids"][k], z["sample_targets"][k]) for k in range(6)]
    values, lengths, targets = O.collate(samples)
    assert np.array_equal(values.numpy(), z["collate_values"]) and np.array_equal(lengths.numpy(), z["collate_lengths"])
    assert np.array_equal(targets.numpy(), z["collate_targets"])


def test_id_remap_helpers_match_oracle():
    import ncf_b200
    assert ncf_b200.first_appearance_index(["b", "a", "b", "c"]) == O.first_appearance_index(["b", "a", "b", "c"])
    assert ncf_b200.remap_product_id("P76681787", 366) == O.remap_product_id("P76681787", 366) == 179
    assert ncf_b200.remap_cardnumber("6011000000005021", 8031) == O.remap_cardnumber("6011000000005021", 8031)


@pytest.mark.gpu
def test_device_sampler_layout_constraints_and_distribution():
    import ncf_b200
    z = load_npz("sampler.npz")
    inter = torch.from_numpy(z["interactions"])
    U, I = int(z["num_users"]), int(z["num_products"])
    s = ncf_b200.InteractionSampler(inter[:, 0], inter[:, 1], U, I, negative_samples=4, batch_size=64, seed=9)
    assert torch.allclose(s.weights.cpu(), torch.from_numpy(z["weights"]), atol=1e-15)
    hist = _history(z["interactions"])
    n_batches, counts = 0, torch.zeros(I, dtype=torch.long)
    seen = []
    M = inter.shape[0]
    last = M - (len(s) - 1) * 64                     # reference padding: batch + batch[:64 - len(batch)]
    sizes = [64] * (len(s) - 1) + [min(64, 2 * last)]
    for epoch in range(40):
        for bi, (kjt, t) in enumerate(s):
            n_batches += 1
            v = kjt.values().cpu()
            N = v.numel() // 2
            nb = sizes[bi]
            assert N == nb * 5 and t.shape == (N, 1)
            u, it = v[:N].view(-1, 5), v[N:].view(-1, 5)
            assert torch.equal(t.cpu().view(-1, 5), torch.tensor([1., 0, 0, 0, 0]).expand(nb, 5))
            assert torch.all(u == u[:, :1])                                   # user repeated S times
            for b in range(nb):
                uu, pos = int(u[b, 0]), int(it[b, 0])
                assert pos in hist[uu]
                for neg in it[b, 1:].tolist():
                    assert neg != pos and neg not in hist[uu]                 # data_prep.py:150
            counts += torch.bincount(it[:, 1:].reshape(-1), minlength=I)
            if epoch == 0:
                seen.append(torch.stack([u[:, 0], it[:, 0]], 1))
    assert n_batches == 40 * len(s)
    # every interaction appears once per epoch, the last batch is padded with its own first rows
    e0 = torch.cat(seen)
    key = lambda x: (x[:, 0] * I + x[:, 1]).sort().values
    assert torch.equal(key(e0[:M]), key(inter))
    assert torch.equal(e0[M:], e0[(len(s) - 1) * 64:(len(s) - 1) * 64 + (e0.shape[0] - M)])
    # negatives follow the inverse-popularity law, restricted per user by the rejection rule: compare with
    # the oracle's expected frequencies (weights renormalised over each user's allowed items)
    w = z["weights"]
    expect = np.zeros(I)
    for uu, pos in z["interactions"].tolist():
        allowed = np.ones(I, bool)
        allowed[list(hist[uu])] = False
        p = w * allowed
        expect += p / p.sum()
    expect = expect / expect.sum()
    freq = (counts.double() / counts.sum()).numpy()
    assert np.abs(freq - expect).max() < 0.01, np.abs(freq - expect).max()
