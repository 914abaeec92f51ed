#!/usr/bin/env python3
"""Benchmark of the AdvancedNCF training hot path on B200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c0|c3shard] [--impl reference]

A "step" = one iteration of the reference's ModelTrainer.train_epoch loop body (forward, BCELoss,
backward, Adam on dense parameters and on the four embedding tables) over one synthetic batch of B
interactions x (1 + 4 negatives) sample rows.  metric = train sample rows / s.
  value     inputs already resident in HBM (NCFTrainEngine.train_step)
  e2e       the same step from pinned HOST buffers: H2D copies of ids+targets and the D2H read of
            the loss inside the timed region (NCFTrainEngine.train_step_host)
  roofline  the embedding-path kernels (HBM-bound) and the towers, timed with CUDA events
  cpu_baseline / --impl reference: the CPU oracle port of the reference step on the host cores
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (users, items, default interactions per step, description)
    "c0": (8031, 366, 256, "config[0] shipped shape 8,031 x 366, batch 256 (config.yaml)"),
    "c1": (6040, 3706, 65536, "config[1] MovieLens-1M shape 6,040 x 3,706"),
    "c2": (138493, 26744, 65536, "config[2] MovieLens-20M shape 138,493 x 26,744"),
    "c3shard": (12500000, 1250000, 65536, "1/8 shard of config[3] (100M x 10M over 8 GPUs): 12.5M x 1.25M per GPU"),
    "c3": (100000000, 10000000, 65536, "config[3] 100M x 10M row-sharded (needs >= 2 GPUs)"),
}
S = 5                      # 1 + negative_samples (config.yaml:65)
BYTES_FWD_PER_INTERACTION = 3072 + 80          # (2 user + 2*S item rows) * 256 B + ids
BYTES_BWD_PER_INTERACTION = 18432              # 12 unique rows * (read w,m,v + write w,m,v)
FLOP_FWD_PER_ROW = 165e3                       # SURVEY 8d dense towers, forward
FLOP_TRAIN_PER_ROW = 495e3
FLOP_ATTN_FWD = 32.8e3 + 1.3e3                 # projections + scores/AV per row; backward = 2x, + 1x recompute
FLOP_MLP_FWD = 131.1e3 + 0.5e3                 # MLP + heads/LN per row; backward = 2x
# bytes per sample row a fused implementation has to move: fp32 row vectors in and out,
# bf16 saved activations between forward and backward
BYTES_ATTN_FWD = 2 * 128 + 128                 # xu, xp (bf16 rows) in; a (bf16) out
BYTES_MLP_FWD = 128 + 2 * (512 + 256) + 128 + 24 + 16              # a in; r1,y1,r2,y2,r3 (bf16), LN stats, scalars out
BYTES_MLP_BWD = 28 + (896 + 24 + 896 + 128) + (128 + 768 + 896)   # head (scalars); chain (r, stats, dz, da); wgrad
BYTES_ATTN_BWD = 3 * 128 + 2 * 128             # xu, xp, da (bf16 rows) in; dxu, dxp (bf16 rows) out


def peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_batches(users, items, B, nbatch, seed, device=None, pin=False):
    """Seeded synthetic interactions: users uniform, positive item ~ Zipf(1.0) over a seeded
    permutation, 4 uniform negatives per positive; rows interaction-major (positive first), targets
    [1,0,0,0,0] (data_prep.py:210-212, 286-303)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(items, generator=g)
    w = 1.0 / torch.arange(1, items + 1, dtype=torch.float64)
    cdf = torch.cumsum(w / w.sum(), 0)
    out = []
    for _ in range(nbatch):
        u = torch.randint(0, users, (B,), generator=g).repeat_interleave(S)
        pos = perm[torch.searchsorted(cdf, torch.rand(B, generator=g, dtype=torch.float64)).clamp_max(items - 1)]
        it = torch.randint(0, items, (B, S), generator=g)
        it[:, 0] = pos
        t = torch.zeros(B, S)
        t[:, 0] = 1.0
        trip = (u.contiguous(), it.reshape(-1).contiguous(), t.reshape(-1).contiguous())
        if device is not None:
            trip = tuple(x.to(device) for x in trip)
        elif pin:
            trip = tuple(x.pin_memory() for x in trip)
        out.append(trip)
    return out


def build_model(users, items, device, precision):
    import torch
    import ncf_b200
    torch.manual_seed(1234)
    m = ncf_b200.AdvancedNCF(users, items, 5, 24, 64, 64, 32, [256, 128, 64], 4, 0.2, 4)
    m.compute_precision = precision
    return m.to(device).train()


def cpu_reference_rate(users, items, B_sample, steps, warmup, threads):
    """The reference training step on the host cores: (rows/s, ms/step, kind).
    kind "reference": the UNMODIFIED reference `ModelTrainer.train_epoch` imported through oracle/shims (only where the
    reference tree is mounted, i.e. the build container); kind "port": the oracle port of the same step
    (oracle/ncf_oracle.py: forward, BCELoss, dense table gradients, torch-Adam update of every row)."""
    import torch
    from oracle import ncf_oracle as O
    from oracle import reference_runner as R
    if R.available() and not os.environ.get("NCF_CPU_PORT"):
        batches = make_batches(users, items, B_sample, 2, 99)
        rate, ms = R.train_rate(users, items, batches, steps, max(warmup, 1), threads)
        return rate, ms, "reference"
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    p = cpu_params(users, items, g)
    rate, ms = _cpu_train_rate(p, users, items, B_sample, steps, warmup, g)
    return rate, ms, "port"


def cpu_params(users, items, g):
    import torch
    from oracle import ncf_oracle as O
    p = {}
    for k, rows in zip(O.TABLE_KEYS, (users, items, users, items)):
        p[k] = (torch.rand(rows, 64, generator=g) * 2 - 1) * (1.0 / rows) ** 0.5
    shapes = {"mf_norm.weight": (64,), "mf_norm.bias": (64,), "mlp_norm.weight": (64,), "mlp_norm.bias": (64,),
              "mlp.0.weight": (256, 96), "mlp.0.bias": (256,), "mlp.2.weight": (256,), "mlp.2.bias": (256,),
              "mlp.4.weight": (128, 256), "mlp.4.bias": (128,), "mlp.6.weight": (128,), "mlp.6.bias": (128,),
              "mlp.8.weight": (64, 128), "mlp.8.bias": (64,), "mlp.10.weight": (64,), "mlp.10.bias": (64,),
              "mf_output.weight": (1, 64), "mf_output.bias": (1,), "mlp_output.weight": (1, 64),
              "mlp_output.bias": (1,), "final.0.weight": (1, 2), "final.0.bias": (1,)}
    for n in ("q", "k", "v", "out"):
        shapes[f"user_product_attention.{n}_proj.weight"] = (64, 64)
        shapes[f"user_product_attention.{n}_proj.bias"] = (64,)
    for k, shp in shapes.items():
        p[k] = torch.ones(shp) if (k.endswith("weight") and len(shp) == 1) else (torch.rand(shp, generator=g) - 0.5) * 0.2
    p["temporal_encoding.hour_embed.weight"] = torch.zeros(24, 32)
    return p


def _cpu_train_rate(p, users, items, B_sample, steps, warmup, g):
    import torch
    from oracle import ncf_oracle as O
    batches = make_batches(users, items, B_sample, 2, 99)
    state = {}
    times = []
    masks = None
    for s in range(warmup + steps):
        u, i, t = batches[s % len(batches)]
        N = u.numel()
        masks = {"attn": torch.rand(N // S, 4, S, S, generator=g) >= 0.2, "mlp0": torch.rand(N, 256, generator=g) >= 0.2,
                 "mlp1": torch.rand(N, 128, generator=g) >= 0.2, "mlp2": torch.rand(N, 64, generator=g) >= 0.2}
        t0 = time.perf_counter()
        O.train_step(p, state, s + 1, u, i, t.view(-1, 1), dropout_p=0.2, masks=masks)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    return B_sample * S / (ms / 1e3), ms


def time_kernel(fn, iters, stream_sync):
    import torch
    for _ in range(3):
        fn()
    stream_sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def timed_blocks(step_fn, K, barrier, world, dev, min_seconds=1.0, max_blocks=200):
    """Times blocks of EXACTLY K steps (CUDA events on the launch stream, barrier + synchronize on both sides, max over
    ranks per block) until at least `min_seconds` of timed work have run, so that clocks are sampled under load; the
    reported number is the MEDIAN block.  Returns the list of block times in ms."""
    import torch
    import torch.distributed as dist
    blocks, total = [], 0.0
    while True:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for s in range(K):
            step_fn(s)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        blocks.append(ms)
        total += ms
        if total >= min_seconds * 1e3 or len(blocks) >= max_blocks:
            return blocks


def wall_blocks(step_fn, K, barrier, world, dev, min_seconds=0.5, max_blocks=100):
    """Same for the end-to-end leg: wall clock around K steps that each end with a device->host read."""
    import torch
    import torch.distributed as dist
    blocks, total = [], 0.0
    while True:
        barrier()
        t0 = time.perf_counter()
        for s in range(K):
            step_fn(s)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0])
        blocks.append(ms)
        total += ms
        if total >= min_seconds * 1e3 or len(blocks) >= max_blocks:
            return blocks


def build_model_on_device(users, items, device, precision):
    """Like build_model but the (large) tables are drawn on the device: no multi-GB host init + copy."""
    import torch
    import ncf_b200
    torch.manual_seed(1234)
    with torch.device(device):
        m = ncf_b200.AdvancedNCF(users, items, 5, 24, 64, 64, 32, [256, 128, 64], 4, 0.2, 4)
    m.compute_precision = precision
    return m.train()


SCORE_SHAPES = {"score": (1000000, 10000000, "config[4] full-catalogue scoring + top-100: 1M users x 10M items"),
                "score_small": (138493, 26744, "full-catalogue scoring + top-100 at the config[2] shape")}


def score_measure(users, items, n, k, steps, warmup, world, rank, dev, min_seconds=0.5, verify_users=0, verify_where="cuda"):
    """Full-catalogue scoring + top-k (app.py:43-77 at scale) on this rank: a step scores `n` users against ALL `items`.
    Users are sharded over the ranks, the folded item side is replicated.  Returns per-rank timing dict
    (block times already max-reduced over ranks)."""
    import torch
    import torch.distributed as dist
    import ncf_b200
    from ncf_b200 import _lib
    lib = _lib.load()
    users_local = (users + world - 1) // world
    model = build_model_on_device(users_local, items, dev, "fp32").eval()      # same seed on every rank: replicated item side
    scorer = ncf_b200.CatalogueScorer(model)
    g = torch.Generator().manual_seed(77 + rank)
    dev_b = [torch.randint(0, users_local, (n,), generator=g).to(dev) for _ in range(4)]
    host_b = [torch.randint(0, users_local, (n,), generator=g).pin_memory() for _ in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for s in range(warmup):
        scorer.topk(dev_b[s % 4], k)
    barrier()
    l0 = int(lib.ncf_launch_count())
    blocks = timed_blocks(lambda s: scorer.topk(dev_b[s % 4], k), steps, barrier, world, dev, min_seconds)
    launches = (int(lib.ncf_launch_count()) - l0) // len(blocks)

    def e2e_step(s):
        idx, sc = scorer.topk(host_b[s % 4].to(dev, non_blocking=True), k)
        idx.cpu(), sc.cpu()
    e2e = wall_blocks(e2e_step, steps, barrier, world, dev, min_seconds)
    out = {"ms_block": statistics.median(blocks), "blocks": len(blocks), "e2e_ms_block": statistics.median(e2e),
           "launches": launches, "tc": scorer.img is not None and n >= scorer.TC_MIN_USERS}
    if verify_users:
        out["verify"] = score_verify(model, scorer, verify_users, k, rank, verify_where)
    del scorer, model
    torch.cuda.empty_cache()
    return out


def score_verify(model, scorer, n_users, k, rank, where="cuda"):
    """seeded users scored by the oracle's forward_simple (oracle/ncf_oracle.py, the reference's per-pair
    arithmetic: full attention + MLP tower for every (user, item)) over ALL items in 1M-item chunks + stable top-k (score
    desc, index asc), compared position by position with the product's result; a differing position must be an fp32
    near-tie (reference gap below 2 ulp of the score).  where = "cpu": the oracle on the host cores (about 15 s per user
    at 10M items); "cuda": the same torch code executed on the device (256 users in seconds)."""
    import torch
    from oracle import ncf_oracle as O
    g = torch.Generator().manual_seed(4242 + rank)
    users = torch.randint(0, model.num_users, (n_users,), generator=g)
    dev = next(model.parameters()).device
    idx, sc = scorer.topk(users.to(dev), k)
    odev = dev if where == "cuda" else torch.device("cpu")
    idx = idx.to(odev)
    sd = {kk: v.detach().to(odev) for kk, v in model.state_dict().items()}
    I = model.num_products
    exact = near = bad = 0
    t0 = time.perf_counter()
    for j, u in enumerate(users.tolist()):
        parts = []
        for s0 in range(0, I, 1 << 20):
            it = torch.arange(s0, min(I, s0 + (1 << 20)), device=odev)
            parts.append(O.forward_simple(sd, torch.full((it.numel(),), u, dtype=torch.long, device=odev), it))
        ref = torch.cat(parts)
        want = O.topk_stable(ref, k)
        d = idx[j] != want
        exact += int((~d).sum())
        if bool(d.any()):
            gap = (ref[idx[j]] - ref[want]).abs()[d]
            ulp2 = 2.4e-7 * ref[want][d].abs().clamp_min(1e-3)
            near += int((gap <= ulp2).sum())
            bad += int((gap > ulp2).sum())
    return {"users": n_users, "items": I, "k": k, "positions": n_users * k, "index_equal": exact, "near_tie_swaps": near,
            "mismatches": bad, "reference_seconds": time.perf_counter() - t0,
            "reference": f"oracle forward_simple in 1M-item chunks + stable top-k (app.py:48-75), torch fp32 on {where}"}


def score_line(args, world, rank, local):
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", local)
    users, items, desc = SCORE_SHAPES[args.workload]
    n = args.batch or 2048          # users per GPU per step: 16 tiles of 128 users for the tensor-core pre-filter
    k = 100
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    r = score_measure(users, items, n, k, args.steps, args.warmup, world, rank, dev, 1.0,
                      verify_users=args.verify_users if args.verify else 0, verify_where=args.verify_where)
    clocks = sampler.stop() if rank == 0 else None
    ver = r.get("verify")
    if world > 1 and ver is not None:
        t = torch.tensor([ver["index_equal"], ver["near_tie_swaps"], ver["mismatches"], ver["positions"]], device=dev)
        dist.all_reduce(t)
        ver.update(index_equal=int(t[0]), near_tie_swaps=int(t[1]), mismatches=int(t[2]), positions=int(t[3]),
                   users=ver["users"] * world)
    if rank != 0:
        return
    pk = peaks()
    pairs = world * n * items * args.steps
    value = pairs / (r["ms_block"] / 1e3)
    tflops = value * 128 / 1e12
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu = cpu_score_rate(items, os.cpu_count() or 1)
    line = {"metric": "scoring_samples_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_block"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "timed_blocks": r["blocks"],
            "config": {"workload": f"{args.workload}: {desc}", "users_per_step_per_gpu": n, "items": items, "top_k": k,
                       "l2": f"folded item table {items * 260 / 1e6:.0f} MB streamed per user tile"},
            "users_per_s": world * n * args.steps / (r["ms_block"] / 1e3),
            "e2e": {"value": pairs / (r["e2e_ms_block"] / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": n * 8,
                    "d2h_bytes_per_step": n * k * 12, "ms_per_step": r["e2e_ms_block"] / args.steps},
            "gpu_launches": r["launches"], "clocks": clocks,
            "roofline": {"kernel": "score_tc_kernel (tcgen05 bf16 upper-bound GEMM + exact fp32 re-scoring of survivors)"
                                   if r["tc"] else "score_topk_kernel (exact fp32)",
                         "bound": "tensor", "achieved": tflops, "peak": world * pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": tflops / (world * pk["bf16_tflops"]), "traffic": None,
                         "peak_source": pk["source"] + " (bf16 burst) x n_gpus; 128 flop per (user, item) pair; the kernel is "
                                        "bound by its per-tile epilogue (threshold scan of the TMEM accumulators), not by the tensor pipe"},
            "cpu_baseline": cpu}
    if ver is not None:
        line["topk_verify"] = ver
    print(json.dumps(line))


def run_score(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    score_line(args, world, rank, local)
    if world > 1:
        dist.destroy_process_group()


def cpu_score_rate(items, threads):
    """Reference scoring on the host: forward_simple over the catalogue for one user (app.py:48-67) + stable
    top-100, through the oracle port, on a bounded sample of the catalogue."""
    import torch
    from oracle import ncf_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(5)
    sample = min(items, 200000)
    p = cpu_params(1000, sample, g)
    t0 = time.perf_counter()
    s = O.forward_simple(p, torch.zeros(sample, dtype=torch.long), torch.arange(sample))
    O.topk_stable(s, 100)
    dt = time.perf_counter() - t0
    return {"value": sample / dt, "unit": "pairs/s", "cores": threads, "kind": "port",
            "sample": f"1 user x {sample} of the {items} items (forward_simple + stable top-100)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    users, items, B, desc = WORKLOADS[args.workload]
    B = args.batch or B
    threads = os.cpu_count() or 1
    B_sample = min(B, args.ref_batch)
    rate, ms, kind = cpu_reference_rate(users, items, B_sample, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": rate, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "interactions_per_step_per_gpu": B, "rows_per_interaction": S,
                       "table_update": "dense torch.optim.Adam on every row (the reference's own semantics)", "towers": "fp32",
                       "l2": "host DRAM; bounded sample of the step"},
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": kind,
                             "sample": f"{B_sample} of the {B} interactions of a step, {args.steps} steps; "
                                       + ("the unmodified reference ModelTrainer.train_epoch through oracle/shims" if kind == "reference"
                                          else "oracle port of the reference step (the reference tree is not on this box)")},
            "e2e": {"value": rate, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + sorted(SCORE_SHAPES))
    ap.add_argument("--batch", type=int, default=0, help="interactions per step per GPU (default per workload)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16"])
    ap.add_argument("--table-mode", default="auto", choices=["auto", "fused_dense_equiv", "fused_sparse"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-batch", type=int, default=4096, help="interactions per CPU-reference step (bounded sample)")
    ap.add_argument("--cpu-baseline-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true",
                    help="score workloads: check the top-k of seeded users against the CPU oracle (SURVEY 8d C4); "
                         "train workloads at N > 1: always on (3 sharded steps vs the single-GPU engine)")
    ap.add_argument("--verify-users", type=int, default=32, help="users per rank checked by --verify (score workloads)")
    ap.add_argument("--verify-where", default="cuda", choices=["cuda", "cpu"],
                    help="run the oracle of --verify on the device (fast) or on the host cores (about 15 s per user at 10M items)")
    ap.add_argument("--min-seconds", type=float, default=1.0, help="timed blocks of --steps steps are repeated until this much "
                                                                    "timed work has run; the median block is reported")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary blocks (scoring, c3shard / c3)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.workload in SCORE_SHAPES:
        if args.impl == "reference":
            if int(os.environ.get("RANK", "0")) == 0:
                users, items, desc = SCORE_SHAPES[args.workload]
                cpu = cpu_score_rate(items, os.cpu_count() or 1)
                print(json.dumps({"impl": "reference", "metric": "scoring_samples_per_s", "value": cpu["value"],
                                  "unit": "pairs/s", "n_gpus": args.gpus, "steps": 1, "warmup": 0, "higher_is_better": True,
                                  "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                                  "config": {"workload": f"{args.workload}: {desc}"}, "cpu_baseline": cpu,
                                  "e2e": {"value": cpu["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0,
                                          "d2h_bytes_per_step": 0}}))
            return
        run_score(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import ncf_b200
    from ncf_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    users, items, B, desc = WORKLOADS[args.workload]
    B = args.batch or B
    N = B * S
    precision = args.precision
    if precision == "auto":
        precision = "bf16" if getattr(_lib, "HAS_BF16_TC", False) else "fp32"
    table_mode = args.table_mode
    if table_mode == "auto":
        table_mode = "fused_sparse" if args.workload in ("c3shard", "c3") else "fused_dense_equiv"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    if world > 1:
        r = sharded_measure(args, users, items, B, precision, table_mode, world, rank, dev, barrier)
    else:
        if args.workload == "c3":
            raise SystemExit("workload c3 (169 GB of tables + Adam state) needs --gpus >= 2")
        r = single_measure(args, users, items, B, precision, table_mode, dev, barrier)
    clocks = sampler.stop() if rank == 0 else None
    ms_step = r["ms_block"] / args.steps
    value = world * N * args.steps / (r["ms_block"] / 1e3)
    e2e_value = world * N * args.steps / (r["e2e_ms_block"] / 1e3)
    pk = peaks()
    hbm = pk["hbm_gbs"]
    sus = pk["bf16_tflops_sustained"]
    # SURVEY 8d basis for the WHOLE step: 4,316.8 algorithmic bytes and 495 kFLOP per sample row
    survey = {"bytes_per_sample": 4316.8, "flop_per_sample": FLOP_TRAIN_PER_ROW,
              "hbm": {"achieved": N * 4316.8 / ms_step / 1e6, "unit": "GB/s per GPU", "peak": hbm,
                      "frac": N * 4316.8 / ms_step / 1e6 / hbm},
              "tensor": {"achieved": N * FLOP_TRAIN_PER_ROW / ms_step / 1e9, "unit": "TFLOP/s per GPU", "peak": sus,
                         "frac": N * FLOP_TRAIN_PER_ROW / ms_step / 1e9 / sus}}

    extras = {}
    if not args.no_extras and args.workload == "c2":
        extras = extra_blocks(args, world, rank, dev, precision, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "fp32", "data": "synthetic", "timed_blocks": r["blocks"],
        "config": {"workload": f"{args.workload}: {desc}", "interactions_per_step_per_gpu": B, "rows_per_interaction": S,
                   "table_update": table_mode, "towers": precision,
                   "timing": f"median of {r['blocks']} blocks of {args.steps} steps (>= 1 s of timed work), CUDA events, max over ranks",
                   "l2": f"inputs larger than L2: activations+ids of one step {N * 4316.8 / 1e6:.0f} MB algorithmic, 4 rotating batches"},
        "interactions_per_s": value / S,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": N * (8 + 8 + 4), "d2h_bytes_per_step": 4,
                "ms_per_step": r["e2e_ms_block"] / args.steps},
        "gpu_launches": r["launches"], "clocks": clocks, "roofline": r.get("roofline"), "survey_basis_whole_step": survey,
        "cpu_baseline": r.get("cpu_baseline"), "last_loss": r["last_loss"],
    }
    if world > 1:
        line["config"]["parallelism"] = (f"tables row-sharded over {world} GPUs (all-to-all ids/rows/grads), towers "
                                         f"data-parallel (dense all-reduce)")
        line["parity_vs_single_gpu"] = r.get("parity_vs_single_gpu")
        line["parity_detail"] = r.get("parity_detail")
        if r.get("phases"):
            line["sharded_phases_ms"] = r["phases"]
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def single_measure(args, users, items, B, precision, table_mode, dev, barrier):
    """N = 1: NCFTrainEngine on the device-resident batches (value), from pinned host buffers (e2e), and every stage of
    the step on its own through the C ABI (roofline)."""
    import torch
    import ncf_b200
    from ncf_b200 import _lib
    lib = _lib.load()
    N = B * S
    big = users * 64 * 4 * 6 > (8 << 30)
    model = build_model_on_device(users, items, dev, precision) if big else build_model(users, items, dev, precision)
    eng = ncf_b200.NCFTrainEngine(model, lr=1e-3, weight_decay=1e-5, table_mode=table_mode, max_rows=N)
    nb = 4
    dev_batches = make_batches(users, items, B, nb, 1234, device=dev)
    host_batches = make_batches(users, items, B, nb, 4321, pin=True)
    for s in range(args.warmup):
        eng.train_step(*dev_batches[s % nb])
    barrier()
    l0 = int(lib.ncf_launch_count())
    blocks = timed_blocks(lambda s: eng.train_step(*dev_batches[s % nb]), args.steps, barrier, 1, dev, args.min_seconds)
    launches = (int(lib.ncf_launch_count()) - l0) // len(blocks)
    last_loss = float(eng.loss.item())
    for s in range(3):
        eng.train_step_host(*host_batches[s % nb])

    def e2e_step(s):
        eng.train_step_host(*host_batches[s % nb], next_batch=host_batches[(s + 1) % nb] if s + 1 < args.steps else None)
    e2e = wall_blocks(e2e_step, args.steps, barrier, 1, dev, args.min_seconds / 2)
    ms_step = statistics.median(blocks) / args.steps
    roofline = stage_roofline(args, model, dev_batches[0], users, items, B, precision, table_mode, dev, ms_step)
    cpu_baseline = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        Bs = min(B, args.ref_batch)
        rate, ms, kind = cpu_reference_rate(users, items, Bs, args.cpu_baseline_steps, 1, threads)
        cpu_baseline = {"value": rate, "unit": "samples/s", "cores": threads, "kind": kind, "ms_per_step": ms,
                        "sample": f"{Bs} of the {B} interactions of a step, {args.cpu_baseline_steps} steps"}
    eng.close()
    del eng, model
    torch.cuda.empty_cache()
    return {"ms_block": statistics.median(blocks), "blocks": len(blocks), "e2e_ms_block": statistics.median(e2e),
            "launches": launches, "last_loss": last_loss, "roofline": roofline, "cpu_baseline": cpu_baseline}


def stage_roofline(args, model, batch, users, items, B, precision, table_mode, dev, ms_step, brief=False):
    """Per-kernel roofline (same inputs, CUDA events on the launch stream): the step is timed piecewise through the C ABI:
    K1 (ncf_gather_ln_gmf_fwd), the tower stages in place on the workspace of a full forward, K6 (ncf_emb_bwd_adam_both)
    and the dense-equivalent sweep.  brief: K1 / K6 / sweep only (the embedding path of another table size)."""
    import torch
    from ncf_b200 import _lib
    lib = _lib.load()
    N = B * S
    pk = peaks()
    u, it, tg = batch
    tabs = model._tables_struct()
    st = torch.cuda.current_stream(dev)
    sync = torch.cuda.synchronize
    sptr = C.c_void_p(st.cuda_stream)
    flat = model._flat
    cfg = _lib.RunCfg()
    cfg.S, cfg.training, cfg.dropout_p, cfg.seed, cfg.step = S, 1, 0.2, 7, 3
    cfg.precision = _lib.NCF_BF16_TC if precision == "bf16" else _lib.NCF_FP32
    wsb = int(lib.ncf_workspace_bytes(N, C.byref(cfg)))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    out = torch.empty(N, device=dev)
    gout = torch.randn(N, device=dev) * 1e-6
    dgrad = torch.zeros(flat.numel(), device=dev)
    mf = torch.empty(N, device=dev)
    xu = torch.empty(N, 64, device=dev)
    xp = torch.empty(N, 64, device=dev)
    row_dt = torch.bfloat16 if precision == "bf16" else torch.float32     # per-sample row arrays: the format the step uses
    ypm = torch.empty(N, 64, device=dev, dtype=row_dt)
    yum = torch.empty(N, 64, device=dev, dtype=row_dt)
    adam = _lib.AdamCfg()
    adam.lr, adam.beta1, adam.beta2, adam.eps, adam.weight_decay, adam.step = 1e-3, 0.9, 0.999, 1e-8, 1e-5, 7
    ews_bytes = int(lib.ncf_emb_bwd_workspace_bytes(N))
    ews = torch.empty(ews_bytes, dtype=torch.uint8, device=dev)
    dmf = torch.randn(N, device=dev) * 1e-6
    dx = (torch.randn(N, 64, device=dev) * 1e-6).to(row_dt)

    k1_fn = lib.ncf_gather_ln_gmf_fwd_bf16 if precision == "bf16" else lib.ncf_gather_ln_gmf_fwd    # the variant the step runs
    k6_fn = lib.ncf_emb_bwd_adam_both_bf16 if precision == "bf16" else lib.ncf_emb_bwd_adam_both   # (bf16 row arrays: the fp32-sized buffers are large enough)

    def k1():
        _lib.check(k1_fn(C.byref(tabs), _lib.ptr(flat), _lib.ptr(u), _lib.ptr(it), N, None, None,
                         _lib.ptr(mf), _lib.ptr(xu), _lib.ptr(xp), _lib.ptr(ypm), _lib.ptr(yum), sptr))

    def fwd():
        _lib.check(lib.ncf_forward(C.byref(cfg), C.byref(tabs), _lib.ptr(flat), _lib.ptr(u), _lib.ptr(it), N, None, None,
                                   None, _lib.ptr(out), _lib.ptr(ws), wsb, sptr))

    def attn_fwd():
        _lib.check(lib.ncf_attn_fwd(C.byref(cfg), _lib.ptr(flat), N, _lib.ptr(ws), wsb, sptr))

    def mlp_fwd():
        _lib.check(lib.ncf_mlp_fwd(C.byref(cfg), _lib.ptr(flat), N, _lib.ptr(out), _lib.ptr(ws), wsb, sptr))

    def mlp_bwd():
        _lib.check(lib.ncf_mlp_bwd(C.byref(cfg), _lib.ptr(flat), _lib.ptr(dgrad), N, _lib.ptr(gout), _lib.ptr(ws), wsb, sptr))

    def attn_bwd():
        _lib.check(lib.ncf_attn_bwd(C.byref(cfg), _lib.ptr(flat), _lib.ptr(dgrad), N, _lib.ptr(ws), wsb, sptr))

    def k6():
        adam.emb_mode = _lib.EMB_ADAM_SPARSE
        _lib.check(k6_fn(C.byref(adam), C.byref(tabs), _lib.ptr(flat), _lib.ptr(dgrad), _lib.ptr(u),
                                             _lib.ptr(it), N, _lib.ptr(dmf), _lib.ptr(dx), _lib.ptr(dx), _lib.ptr(ypm),
                                             _lib.ptr(yum), _lib.ptr(ews), ews_bytes, sptr))

    def sweep():
        adam.emb_mode = _lib.EMB_ADAM_DENSE_EQUIV
        _lib.check(lib.ncf_emb_adam_sweep(C.byref(adam), C.byref(tabs), sptr))

    hbm = pk["hbm_gbs"]
    sus = pk["bf16_tflops_sustained"]
    table_mb = 2 * (users + items) * 256 / 1e6
    resident = "L2-resident" if table_mb < 120 else "HBM-resident"
    residency = (f"{resident}: the four tables are {table_mb:.0f} MB against the 126 MB L2"
                 + (" - the fraction below is an L2-assisted figure, NOT an HBM measurement (see c3shard_embedding_path for "
                    "the HBM-resident one)" if resident == "L2-resident" else ""))

    def hbm_piece(ms, nbytes, **extra):
        d = {"ms": ms, "bound": "hbm", "achieved": nbytes / ms / 1e6, "unit": "GB/s", "peak": hbm,
             "frac": nbytes / ms / 1e6 / hbm, "algorithmic_bytes": nbytes}
        d.update(extra)
        return d

    k1_bytes = B * BYTES_FWD_PER_INTERACTION
    k6_bytes = B * BYTES_BWD_PER_INTERACTION
    k1()
    k1_ms = time_kernel(k1, 20, sync)
    k6_ms = time_kernel(k6, 10, sync)
    emb_ms, emb_bytes = k1_ms + k6_ms, k1_bytes + k6_bytes
    emb = {"ms": emb_ms, "achieved": emb_bytes / emb_ms / 1e6, "unit": "GB/s", "peak": hbm, "frac": emb_bytes / emb_ms / 1e6 / hbm,
           "tables": residency,
           "K1": hbm_piece(k1_ms, k1_bytes, gather_effective_gbs=4 * N * 256 / k1_ms / 1e6, gather_lookups=4 * N,
                           gather_note="SURVEY 8d gather GB/s: every lookup counted (4 rows x 256 B per sample, no de-duplication); " + resident),
           "K6": hbm_piece(k6_ms, k6_bytes)}
    if brief:
        return emb

    def tensor_piece(ms, flop_per_row, bytes_per_row):
        # SURVEY 8d counts the towers against the tensor pipe; their arithmetic intensity (flop / byte of
        # activations a fused kernel has to move) is far below the bf16 ridge, so the HBM view is given as well
        tf = N * flop_per_row / ms / 1e9
        gb = N * bytes_per_row / ms / 1e6
        return {"ms": ms, "bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "peak": sus, "frac": tf / sus,
                "algorithmic_flop": N * flop_per_row, "hbm_view": {"algorithmic_bytes": N * bytes_per_row, "achieved": gb,
                                                                   "unit": "GB/s", "peak": hbm, "frac": gb / hbm,
                                                                   "flop_per_byte": flop_per_row / bytes_per_row}}

    # Stage calls run in place on the workspace a full forward left behind (include/ncf_b200.h); every tower stage moves
    # more bytes per call than the 126 MB L2 holds.
    fwd()
    afwd_ms = time_kernel(attn_fwd, 10, sync)
    mfwd_ms = time_kernel(mlp_fwd, 10, sync)
    mbwd_ms = time_kernel(mlp_bwd, 10, sync)
    abwd_ms = time_kernel(attn_bwd, 10, sync)
    sweep_ms = time_kernel(sweep, 10, sync) if table_mode == "fused_dense_equiv" else 0.0
    sweep_bytes = 2 * (users + items) * 1536
    # Both backward towers the way the step schedules them (ncf_backward without the embedding half, auxiliary stream set):
    # the MLP weight-gradient kernel runs on a side stream over 28 SMs NEXT TO the attention backward, so the two stages
    # above, timed one after the other, sum to more than the step spends on them.
    sched = None
    if tcp_sched := (precision == "bf16"):
        aux_s = torch.cuda.Stream(device=dev)

        def bwd_towers():
            adam.emb_mode = _lib.EMB_NONE
            _lib.check(lib.ncf_backward(C.byref(cfg), C.byref(adam), C.byref(tabs), _lib.ptr(flat), _lib.ptr(dgrad), _lib.ptr(u),
                                        _lib.ptr(it), N, _lib.ptr(gout), _lib.ptr(ws), wsb, sptr))
        lib.ncf_set_aux_stream(C.c_void_p(aux_s.cuda_stream))
        try:
            both_ms = time_kernel(bwd_towers, 10, sync)
        finally:
            lib.ncf_set_aux_stream(None)
        tf = N * (2 * FLOP_MLP_FWD + 3 * FLOP_ATTN_FWD) / both_ms / 1e9
        sched = {"ms": both_ms, "one_after_the_other_ms": mbwd_ms + abwd_ms, "achieved": tf, "unit": "TFLOP/s", "peak": sus,
                 "frac": tf / sus, "what": "head_bwd + mlp_tc_bwd2, then attn_tc_bwd on 120 SMs next to mlp_tc_wgrad on 28 SMs "
                                           "(side stream), as ncf_train_step runs them"}
    tcp = precision == "bf16"
    kernels = {
        "K1 gather_ln_gmf_fwd": emb["K1"],
        "attention forward" + (" (attn_tc_fwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(afwd_ms, FLOP_ATTN_FWD, BYTES_ATTN_FWD),
        "MLP forward" + (" (mlp_tc_fwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(mfwd_ms, FLOP_MLP_FWD, BYTES_MLP_FWD),
        "MLP backward" + (" (head_bwd + mlp_tc_bwd + mlp_tc_wgrad kernels)" if tcp else " (fp32 kernels)"):
            tensor_piece(mbwd_ms, 2 * FLOP_MLP_FWD, BYTES_MLP_BWD),
        "attention backward" + (" (attn_tc_bwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(abwd_ms, 3 * FLOP_ATTN_FWD, BYTES_ATTN_BWD),
        "K6 emb_bwd_adam_both (1 sort + segment-sum + apply, both sides)": emb["K6"],
    }
    if sweep_ms:
        kernels["dense-equivalent Adam sweep"] = hbm_piece(sweep_ms, sweep_bytes, note=residency)
    top = max(kernels, key=lambda k: kernels[k]["ms"])
    kt = kernels[top]
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")     # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(top)
    pieces_ms = k1_ms + afwd_ms + mfwd_ms + mbwd_ms + abwd_ms + k6_ms + sweep_ms
    step_bytes = N * (BYTES_ATTN_FWD + BYTES_MLP_FWD + BYTES_MLP_BWD + BYTES_ATTN_BWD) + emb_bytes + (sweep_bytes if sweep_ms else 0)
    roofline = {"kernel": top, "bound": kt["bound"], "achieved": kt["achieved"], "peak": kt["peak"], "unit": kt["unit"],
                "frac": kt["frac"], "traffic": traffic,
                "peak_source": pk["source"] + (" (sustained bf16)" if kt["bound"] == "tensor" else " (copy)"),
                "embedding_path": {k: v for k, v in emb.items() if k not in ("K1", "K6")},
                "whole_step_hbm_view": {"algorithmic_bytes": step_bytes, "ms": ms_step, "achieved": step_bytes / ms_step / 1e6,
                                        "unit": "GB/s", "frac": step_bytes / ms_step / 1e6 / hbm,
                                        "note": "implementation view: counts the activation bytes this implementation moves "
                                                "between its kernels; the SURVEY-basis fractions are in survey_basis_whole_step"},
                "pieces_sum_ms": pieces_ms, "kernels": kernels}
    if sched:
        roofline["backward_towers_as_scheduled"] = sched
    if "hbm_view" in kt:          # tower stages: SURVEY 8d counts them against the tensor pipe; their intensity says HBM
        roofline["hbm_view"] = kt["hbm_view"]
    return roofline


def sharded_measure(args, users, items, B, precision, table_mode, world, rank, dev, barrier):
    """N > 1: ShardedNCFEngine (tables row-sharded) on this rank's batches; also verifies the NCCL path against
    the single-GPU engine on the gathered batch (parity_vs_single_gpu)."""
    import torch
    import torch.distributed as dist
    from ncf_b200 import _lib
    from ncf_b200.sharding import ShardedNCFEngine
    lib = _lib.load()
    N = B * S
    parity, detail = verify_sharded(world, rank, dev)
    model = build_model(1, 1, dev, precision)
    eng = ShardedNCFEngine(model, users, items, lr=1e-3, weight_decay=1e-5, table_mode=table_mode)
    nb = 4
    dev_batches = make_batches(users, items, B, nb, 1234 + rank, device=dev)
    host_batches = make_batches(users, items, B, nb, 4321 + rank, pin=True)
    def step_host(u, i, t, nxt=None):
        return eng.train_step_host(u, i, t, next_batch=nxt)

    def dev_step(s):
        eng.train_step(*dev_batches[s % nb], next_ids=dev_batches[(s + 1) % nb][:2] if s + 1 < args.steps else None)
    for s in range(args.warmup):
        eng.train_step(*dev_batches[s % nb])
    barrier()
    l0 = int(lib.ncf_launch_count())
    blocks = timed_blocks(dev_step, args.steps, barrier, world, dev, args.min_seconds)
    launches = (int(lib.ncf_launch_count()) - l0) // len(blocks)
    last_loss = float(eng.loss.item())
    for s in range(3):
        step_host(*host_batches[s % nb])

    def e2e_step(s):
        step_host(*host_batches[s % nb], nxt=host_batches[(s + 1) % nb] if s + 1 < args.steps else None)
    e2e = wall_blocks(e2e_step, args.steps, barrier, world, dev, args.min_seconds / 2)
    # the loss train_step_host hands back early (its own all-reduce behind the forward) is the step's global loss
    early = step_host(*host_batches[0])
    final = float(eng.loss.item())
    if abs(early - final) > 1e-6 * max(1.0, abs(final)):
        raise RuntimeError(f"train_step_host returned {early}, the step's loss is {final}")
    phases = profile_phases(eng, dev_batches, nb, 6)
    ms_step = statistics.median(blocks) / args.steps
    pk = peaks()
    gbs = N * 4316.8 / ms_step / 1e6
    roofline = {"kernel": "whole step per GPU on the SURVEY 8d basis (4,316.8 algorithmic B per sample; NVLink exchange not counted; "
                          "the per-kernel pieces are measured by the N = 1 run)", "bound": "hbm", "achieved": gbs,
                "peak": pk["hbm_gbs"], "unit": "GB/s per GPU", "frac": gbs / pk["hbm_gbs"], "traffic": None,
                "peak_source": pk["source"] + " (copy)"}
    del eng
    torch.cuda.empty_cache()
    return {"ms_block": statistics.median(blocks), "blocks": len(blocks), "e2e_ms_block": statistics.median(e2e),
            "launches": launches, "last_loss": last_loss, "roofline": roofline, "parity_vs_single_gpu": parity,
            "parity_detail": detail, "phases": phases}


def profile_phases(eng, dev_batches, nb, steps):
    """CUDA-event time of every phase of the sharded step (ShardedNCFEngine.profile), a few extra steps after the timed region;
    the per-phase synchronisation removes overlap, so the phases sum to more than a timed step."""
    eng.profile = True
    try:
        if getattr(eng, "phase_ms", None):
            eng.phase_ms.clear()
        for s in range(steps):
            eng.train_step(*dev_batches[s % nb])
        prof = dict(getattr(eng, "phase_ms", {}) or {})
    finally:
        eng.profile = False
    n = prof.pop("steps", 0)
    return {k: v / n for k, v in prof.items()} if n else None


def verify_sharded(world, rank, dev, steps=3):
    """The REAL NCCL / peer-memory path against the single-GPU engine: `steps` steps of ShardedNCFEngine, fp32, no dropout
    (per-rank Philox streams differ from one global stream); rank 0 runs NCFTrainEngine on the all-gathered batch with the
    SAME tables.  Two passes:
      sharp   Adam eps = 1e-4, far above every gradient of this problem (1e-7 .. 1e-5): the update is lr * m / (sqrt(v) + eps)
              ~ linear in the gradient, so rounding noise is NOT amplified and a lost, doubled or stale row (any logic error of
              the exchange: the MF and MLP halves of a row travel together) is off by >= 1e-5.  Asserted: loss 2e-6 every step;
              the two MF tables equal to 1e-6 EVERYWHERE; the two MLP tables to 5e-4 (the share of their rows that is off by more
              than 1e-7 is reported).  The MLP tables get the loose bound because their gradient passes three ReLUs: about once per
              step some pre-activation of the 9 M in a batch lies within rounding noise of 0 (fp32 atomics order the dense-gradient
              sums differently from run to run, in BOTH engines), its relu' flips, and that ONE interaction's user row and five
              item rows move by a fraction of an update - then whatever those rows touch in the next steps (5,003 items, every one
              in ~4 interactions of a step: after 6 steps half of the rows have moved).  Measured with
              tools/stress_pair.py: two single-GPU engines on one GPU differ the same way on the same interactions, and a 1e-7
              relative perturbation of the dense weights reproduces it; the forward value (ReLU is continuous) and with it the
              loss and the MF-side gradients do not move;
      stock   Adam eps = 1e-8 (the reference's): loss within 2e-6 every step.  Its tables are reported, not asserted: Adam turns
              such a flip into whole +-lr steps."""
    import torch
    import torch.distributed as dist
    import ncf_b200
    from ncf_b200.sharding import ShardedNCFEngine
    U, I, Bv = 20011, 5003, 2048
    KEYS = ("mf_embedding_collection.embedding_bags.user_id.weight", "mf_embedding_collection.embedding_bags.product_id.weight",
            "mlp_embedding_collection.embedding_bags.user_id.weight", "mlp_embedding_collection.embedding_bags.product_id.weight")

    def one_pass(eps):
        torch.manual_seed(99)
        tables = [(torch.rand(r, 64) * 2 - 1) * (1.0 / r) ** 0.5 for r in (U, I, U, I)]
        model = build_model(1, 1, dev, "fp32")
        model.dropout = 0.0
        eng = ShardedNCFEngine(model, U, I, lr=1e-3, eps=eps, weight_decay=1e-5, table_mode="fused_dense_equiv", init_tables=tables)
        batches = make_batches(U, I, Bv, steps, 555 + rank, device=dev)
        single = None
        if rank == 0:
            ref_model = ncf_b200.AdvancedNCF(U, I, 5, 24, dropout=0.0)
            sd = ref_model.state_dict()
            for k, v in model.state_dict().items():
                if "embedding_collection" not in k:
                    sd[k] = v.detach().cpu().clone()
            for k, t in zip(KEYS, tables):
                sd[k] = t.clone()
            ref_model.load_state_dict(sd)
            ref_model = ref_model.to(dev).train()
            single = ncf_b200.NCFTrainEngine(ref_model, lr=1e-3, eps=eps, weight_decay=1e-5, table_mode="fused_dense_equiv")
        worst_loss = 0.0
        for s in range(steps):
            u, i, t = batches[s]
            loss = float(eng.train_step(u, i, t).item())
            parts = [[torch.empty_like(x) for _ in range(world)] for x in (u, i, t)]
            for p, x in zip(parts, (u, i, t)):
                dist.all_gather(p, x)
            if rank == 0:
                ref = float(single.train_step(torch.cat(parts[0]), torch.cat(parts[1]), torch.cat(parts[2])).item())
                worst_loss = max(worst_loss, abs(loss - ref))
        got = eng.gather_tables()
        r = {"loss": worst_loss, "mf": 0.0, "mlp": 0.0, "mlp_rows_off": 0.0, "frac6": 0.0}
        if rank == 0:
            for k, (g, w) in enumerate(zip(got, single.model._table_params())):
                d = (g - w.detach()).abs()
                fam = "mf" if k < 2 else "mlp"
                r[fam] = max(r[fam], float(d.max()))
                r["frac6"] = max(r["frac6"], float((d > 8e-6).float().mean()))
                if k >= 2:
                    r["mlp_rows_off"] = max(r["mlp_rows_off"], float((d.max(dim=1).values > 1e-7).float().mean()))
            single.close()
        eng.close()
        del eng
        torch.cuda.empty_cache()
        return r

    sharp = one_pass(1e-4)
    stock = one_pass(1e-8)
    ok, detail = True, None
    if rank == 0:
        ok = sharp["loss"] <= 2e-6 and sharp["mf"] <= 1e-6 and sharp["mlp"] <= 5e-4 and stock["loss"] <= 2e-6
        detail = {"steps": steps, "world": world, "shape": f"{U} x {I}, {Bv} interactions per rank, fp32, dropout 0",
                  "adam_eps_1e-4": {"max_abs_loss_diff": sharp["loss"], "max_abs_diff_mf_tables": sharp["mf"],
                                    "max_abs_diff_mlp_tables": sharp["mlp"], "frac_mlp_rows_off_by_more_than_1e-7": sharp["mlp_rows_off"],
                                    "asserted": "loss 2e-6; MF tables 1e-6 everywhere; MLP tables 5e-4 (relu' flips of single "
                                                "interactions under rounding noise, see bench.verify_sharded)"},
                  "adam_eps_1e-8": {"max_abs_loss_diff": stock["loss"], "max_abs_table_diff": max(stock["mf"], stock["mlp"]),
                                    "frac_table_elements_off_by_more_than_8e-6": stock["frac6"], "asserted": "loss 2e-6"}}
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    return bool(flag.item()), detail


def extra_blocks(args, world, rank, dev, precision, barrier):
    """Secondary measurements the BASELINE metric names next to the c2 headline:
      N = 1 : "scoring"  config[4] shape, 2,048 users x 10M items, top-100 (pairs/s, tensor fraction, e2e, top-k check)
              "c3shard_embedding_path"  K1 / K6 on one GPU's 1/8 share of config[3] (tables HBM-resident)
      N > 1 : "c3"  config[3] 100M x 10M row-sharded over the N GPUs: ms/step, samples/s, embedding-path fraction per GPU"""
    import torch
    import torch.distributed as dist
    out = {}
    pk = peaks()
    if world == 1:
        users, items, desc = SCORE_SHAPES["score"]
        n, k, steps = 2048, 100, 5
        r = score_measure(users, items, n, k, steps, 3, 1, 0, dev, 0.3, verify_users=16)
        pairs_s = n * items * steps / (r["ms_block"] / 1e3)
        tf = pairs_s * 128 / 1e12
        out["scoring"] = {"metric": "scoring_samples_per_s", "value": pairs_s, "unit": "pairs/s", "users_per_s": pairs_s / items,
                          "config": {"workload": "score: " + desc, "users_per_step": n, "items": items, "top_k": k},
                          "ms_per_step": r["ms_block"] / steps,
                          "e2e": {"value": n * items * steps / (r["e2e_ms_block"] / 1e3), "unit": "pairs/s",
                                  "h2d_bytes_per_step": n * 8, "d2h_bytes_per_step": n * k * 12},
                          "roofline": {"kernel": "score_tc_kernel", "bound": "tensor", "achieved": tf, "peak": pk["bf16_tflops"],
                                       "unit": "TFLOP/s", "frac": tf / pk["bf16_tflops"], "traffic": None,
                                       "peak_source": pk["source"] + " (bf16 burst); 128 flop per (user, item) pair"},
                          "topk_verify": r.get("verify"), "gpu_launches": r["launches"],
                          "cpu_baseline": None if args.no_cpu_baseline else cpu_score_rate(items, os.cpu_count() or 1)}
        users3, items3, B3, desc3 = WORKLOADS["c3shard"]
        m3 = build_model_on_device(users3, items3, dev, precision)
        m3.configure_table_optimizer("fused_sparse")
        m3._ensure_flat()
        b3 = make_batches(users3, items3, B3, 1, 1234, device=dev)[0]
        out["c3shard_embedding_path"] = stage_roofline(args, m3, b3, users3, items3, B3, precision, "fused_sparse", dev, 1.0, brief=True)
        out["c3shard_embedding_path"]["workload"] = "c3shard: " + desc3
        del m3
        torch.cuda.empty_cache()
    else:
        from ncf_b200.sharding import ShardedNCFEngine
        users3, items3, B3, desc3 = WORKLOADS["c3"]
        N3 = B3 * S
        per_gpu_gb = 2 * ((users3 + items3) // world) * 256 * 3 / 1e9
        ft = torch.tensor([torch.cuda.mem_get_info(dev)[0] / 1e9], device=dev)
        dist.all_reduce(ft, op=dist.ReduceOp.MIN)               # the same decision on every rank
        free = float(ft[0])
        if per_gpu_gb + 12 > free:
            if rank == 0:
                out["c3"] = {"unavailable": f"needs {per_gpu_gb:.0f} GB per GPU for tables + Adam state, {free:.0f} GB free"}
            out.update(sharded_score_block(args, world, rank, dev, barrier))
            return out
        model = build_model(1, 1, dev, precision)
        eng = ShardedNCFEngine(model, users3, items3, lr=1e-3, weight_decay=1e-5, table_mode="fused_sparse")
        nb = 4
        batches = make_batches(users3, items3, B3, nb, 777 + rank, device=dev)
        for s in range(3):
            eng.train_step(*batches[s % nb])
        barrier()
        steps = max(5, args.steps // 2)
        blocks = timed_blocks(lambda s: eng.train_step(*batches[s % nb], next_ids=batches[(s + 1) % nb][:2] if s + 1 < steps else None),
                              steps, barrier, world, dev, 0.5)
        phases = profile_phases(eng, batches, nb, 6)
        ms = statistics.median(blocks) / steps
        emb_keys = ("owner rows", "pull rows (P2P)", "requester segment sum + push (inside backward)", "owner update")
        t = torch.tensor([sum(phases.get(k, 0.0) for k in emb_keys)] if phases else [0.0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        emb_ms = float(t[0])
        emb_bytes = B3 * (BYTES_FWD_PER_INTERACTION + BYTES_BWD_PER_INTERACTION)      # per GPU: it owns 1/world of every batch's rows
        if rank == 0:
            out["c3"] = {"workload": "c3: " + desc3, "n_gpus": world, "ms_per_step": ms, "samples_per_s": world * N3 / (ms / 1e3),
                         "interactions_per_step_per_gpu": B3, "table_update": "fused_sparse", "tables_gb_per_gpu": per_gpu_gb,
                         "embedding_path": {"what": "requester pull gather + LayerNorm over peer memory (ncf_shard_pull_rows), requester "
                                                    "segment sum + push of the gradient rows (inside ncf_shard_backward_push: that call "
                                                    "minus the towers' stage calls) and owner segment-sum + LayerNorm backward + Adam "
                                                    "(ncf_shard_owner_update); max over ranks, timed phase by phase (no overlap)",
                                            "ms": emb_ms, "algorithmic_bytes_per_gpu": emb_bytes,
                                            "achieved": emb_bytes / emb_ms / 1e6 if emb_ms else None, "unit": "GB/s per GPU",
                                            "peak": pk["hbm_gbs"], "frac": emb_bytes / emb_ms / 1e6 / pk["hbm_gbs"] if emb_ms else None,
                                            "tables": "HBM-resident"},
                         "whole_step_survey_basis": {"achieved": N3 * 4316.8 / ms / 1e6, "unit": "GB/s per GPU", "peak": pk["hbm_gbs"],
                                                     "frac": N3 * 4316.8 / ms / 1e6 / pk["hbm_gbs"]},
                         "sharded_phases_ms": phases}
        del eng
        torch.cuda.empty_cache()
        out.update(sharded_score_block(args, world, rank, dev, barrier))
    return out


def sharded_score_block(args, world, rank, dev, barrier):
    """config[4] through the SHARDED scorer: tables row-sharded over the ranks, the folded item side all-gathered
    once, every rank ranks users it owns against all 10M items (tensor-core pre-filter + exact re-scoring); a few users per
    rank are checked against the oracle's forward_simple + stable top-k run on the gathered tables."""
    import torch
    import torch.distributed as dist
    from ncf_b200.sharding import ShardedNCFEngine, ShardedCatalogueScorer
    from oracle import ncf_oracle as O
    users, items, desc = SCORE_SHAPES["score"]
    n, k, steps = 2048, 100, 5
    model = build_model(1, 1, dev, "fp32").eval()
    eng = ShardedNCFEngine(model, users, items, table_mode="fused_sparse", exchange="nccl")
    t0 = time.perf_counter()
    scorer = ShardedCatalogueScorer(eng)
    torch.cuda.synchronize()
    fold_s = time.perf_counter() - t0
    g = torch.Generator().manual_seed(99 + rank)
    batches = [torch.randint(0, eng.rows_u, (n,), generator=g).to(dev) for _ in range(4)]
    for s in range(3):
        scorer.topk_local_users(batches[s % 4], k)
    barrier()
    blocks = timed_blocks(lambda s: scorer.topk_local_users(batches[s % 4], k), steps, barrier, world, dev, 0.3)
    ms = statistics.median(blocks) / steps
    # verification: 4 users per rank against the reference arithmetic on the gathered tables
    nv = 4
    vu = torch.randint(0, eng.rows_u, (nv,), generator=g).to(dev)
    idx, _ = scorer.topk_local_users(vu, k)
    tabs = eng.gather_tables()
    sd = {kk: v.detach() for kk, v in model.state_dict().items()}
    for key, t in zip(O.TABLE_KEYS, tabs):
        sd[key] = t
    block_u = (users + world - 1) // world
    exact = near = bad = 0
    for j, lu in enumerate(vu.tolist()):
        gu = rank * block_u + lu
        parts = []
        for s0 in range(0, items, 1 << 20):
            it = torch.arange(s0, min(items, s0 + (1 << 20)), device=dev)
            parts.append(O.forward_simple(sd, torch.full((it.numel(),), gu, dtype=torch.long, device=dev), it))
        ref = torch.cat(parts)
        want = O.topk_stable(ref, k)
        d = idx[j] != want
        exact += int((~d).sum())
        if bool(d.any()):
            gap = (ref[idx[j]] - ref[want]).abs()[d]
            ulp2 = 2.4e-7 * ref[want][d].abs().clamp_min(1e-3)
            near += int((gap <= ulp2).sum())
            bad += int((gap > ulp2).sum())
    t = torch.tensor([exact, near, bad], device=dev)
    dist.all_reduce(t)
    del scorer, eng, tabs, sd
    torch.cuda.empty_cache()
    if rank != 0:
        return {}
    pk = peaks()
    pairs_s = world * n * items / (ms / 1e3)
    tf = pairs_s * 128 / 1e12
    return {"scoring": {"metric": "scoring_samples_per_s", "value": pairs_s, "unit": "pairs/s", "users_per_s": pairs_s / items,
                        "n_gpus": world, "ms_per_step": ms,
                        "config": {"workload": "score: " + desc, "users_per_step_per_gpu": n, "items": items, "top_k": k,
                                   "parallelism": "users sharded with the tables (ShardedCatalogueScorer), folded item side all-gathered "
                                                  f"once ({fold_s:.2f} s incl. the item fold): no merge step"},
                        "roofline": {"kernel": "score_tc_kernel", "bound": "tensor", "achieved": tf, "peak": world * pk["bf16_tflops"],
                                     "unit": "TFLOP/s", "frac": tf / (world * pk["bf16_tflops"]), "traffic": None,
                                     "peak_source": pk["source"] + " (bf16 burst) x n_gpus; 128 flop per (user, item) pair"},
                        "topk_verify": {"users": nv * world, "items": items, "k": k, "positions": nv * world * k,
                                        "index_equal": int(t[0]), "near_tie_swaps": int(t[1]), "mismatches": int(t[2]),
                                        "reference": "oracle forward_simple over the gathered tables + stable top-k, torch fp32 on cuda"}}}


if __name__ == "__main__":
    main()
