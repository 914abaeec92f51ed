"""CPU-side checks (no GPU compute): the C-ABI library loads and exports every symbol the header
declares, the host mirror keeps the reference's state_dict contract, metrics match the reference."""
import os
import re

import pytest
import torch

from tests.helpers import expected_metrics, golden_params, load_npz

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import ctypes
    from ncf_b200 import _lib
    hdr = open(os.path.join(REPO, "include", "ncf_b200.h")).read()
    declared = set(re.findall(r"NCF_API\s+[\w\s\*]+?\b(ncf_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    version = int(re.search(r"#define NCF_ABI_VERSION (\d+)", hdr).group(1))
    assert _lib.load().ncf_version() == version


def test_dense_layout_matches_module_parameters():
    import ncf_b200
    from ncf_b200 import _lib
    layout, numel = _lib.dense_layout()
    m = ncf_b200.AdvancedNCF(10, 10, 5, 24)
    named = dict(m.named_parameters())
    total = 0
    prev_end = 0
    for key, off, size in layout:
        assert named[key].numel() == size, key
        assert off % 4 == 0 and off >= prev_end
        prev_end = off + size
        total += size
    assert total == 83909 and numel >= prev_end          # 83,909 active dense values


def test_state_dict_contract_and_golden_checkpoint_loads_strict():
    import ncf_b200
    p, _ = golden_params()
    m = ncf_b200.AdvancedNCF(8031, 366, 5, 24, 64, 64, 32, [256, 128, 64], 4, 0.2, 4)
    assert list(m.state_dict().keys()) == list(p.keys())     # same 62 keys, same order as the checkpoint
    m.load_state_dict(p, strict=True)
    for attr in ("num_users", "num_products", "num_departments", "num_categories", "mf_embedding_dim",
                 "temporal_dim", "num_heads", "negative_samples"):
        assert hasattr(m, attr)
    assert m.user_product_attention.scale == 4.0 and m.final[0].weight.shape == (1, 2)


def test_module_refuses_cpu_and_bad_shapes():
    import ncf_b200
    m = ncf_b200.AdvancedNCF(10, 10, 5, 24).eval()
    with pytest.raises(ncf_b200.NcfError):
        m(ncf_b200.make_kjt(torch.arange(4), torch.arange(4)))
    m.train()
    with pytest.raises(ValueError):
        m(ncf_b200.make_kjt(torch.arange(4), torch.arange(4)))
    with pytest.raises(NotImplementedError):
        ncf_b200.AdvancedNCF(10, 10, 5, 24, mf_embedding_dim=32)._check_geometry()


def test_kjt_layout():
    import ncf_b200
    k = ncf_b200.make_kjt(torch.tensor([5, 6, 7]), torch.tensor([1, 2, 3]))
    assert k.keys() == ["user_id", "product_id"]
    assert k.values().tolist() == [5, 6, 7, 1, 2, 3] and k.lengths().tolist() == [1] * 6
    k2 = ncf_b200.KeyedJaggedTensor(keys=["user_id", "product_id"], values=torch.tensor([0, 9]),
                                    offsets=torch.tensor([0, 1, 2]))
    assert k2.lengths().tolist() == [1, 1]


def test_metrics_match_reference_fixture():
    import ncf_b200
    z = load_npz("metrics.npz")
    exp = expected_metrics(z)
    for name in ("a", "b", "c"):
        P = torch.from_numpy(z[f"{name}/pred"])
        T = torch.from_numpy(z[f"{name}/target"])
        neg = int(z[f"{name}/neg"])
        got = ncf_b200.calculate_metrics(P.reshape(-1, 1), T.reshape(-1, 1), [1, 5, 10], P.shape[0], neg)
        assert set(got) == set(exp[name])
        for k, v in got.items():
            if name == "c" and not k.startswith(("auc", "acc", "pos_", "neg_")):
                continue      # tie order of torch.sort is unspecified in the reference itself
            assert abs(v - exp[name][k]) < 1e-6, (name, k, v, exp[name][k])
    with pytest.raises(ValueError):
        ncf_b200.calculate_metrics(torch.rand(10), torch.rand(10), [1], 3, 4)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs on the host cores only (the oracle port of the reference step) and prints ONE
    JSON line with the keys the driver reads; nothing of it may touch CUDA."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(repo, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-batch", "64"], capture_output=True, text=True, timeout=300,
                         env={**os.environ, "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("c2") and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
