"""Shared helpers for the parity tests (fixtures -> oracle parameter dicts)."""
import json
import os

import numpy as np
import torch

from oracle import ncf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_npz(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_params():
    """Shipped checkpoint as an oracle parameter dict; user tables are rebuilt at full height
    (8031 rows) with the kept rows in place and zeros elsewhere."""
    z = load_npz("golden_ckpt_compact.npz")
    keep = torch.from_numpy(z["user_rows_kept"])
    U = int(z["num_users"])
    p = {}
    for k in z.files:
        if not k.startswith("sd/"):
            continue
        name = k[3:]
        t = torch.from_numpy(z[k])
        if name in (O.K_UMF, O.K_UMLP):
            full = torch.zeros(U, t.shape[1])
            full[keep] = t
            t = full
        p[name] = t
    return p, z


def small_params(z, prefix="init/"):
    """Parameters of the AdvancedNCF(97, 53) fixtures: shipped dense weights + the fixture's tables."""
    p, _ = golden_params()
    for k in O.TABLE_KEYS:
        p[k] = torch.from_numpy(z[prefix + k]).clone()
    return {k: v.clone() for k, v in p.items()}


def rel_err(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def expected_metrics(z):
    return json.loads(str(z["expected_json"]))
