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
BYTES_ATTN_BWD = 3 * 128 + 2 * 256             # xu, xp, da (bf16 rows) in; dxu, dxp (fp32) out


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
    """The CPU oracle port of the reference step (oracle/ncf_oracle.py: forward, BCELoss, dense table
    gradients, torch-Adam update of every row) on the host cores.  Returns (rows/s, ms/step)."""
    import torch
    from oracle import ncf_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(1234)
    p = cpu_params(users, items, g)
    return _cpu_train_rate(p, users, items, B_sample, steps, warmup, g)


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


SCORE_SHAPES = {"score": (1000000, 10000000, "config[4] full-catalogue scoring + top-100: 1M users x 10M items"),
                "score_small": (138493, 26744, "full-catalogue scoring + top-100 at the config[2] shape")}


def run_score(args):
    """Full-catalogue scoring + top-100 (app.py:43-77 at scale).  A step scores `--batch` users (default 2048
    per GPU) against ALL items; metric = (user,item) pairs scored per second.  Users are sharded over the
    ranks, the folded item side is replicated."""
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
    users, items, desc = SCORE_SHAPES[args.workload]
    n = args.batch or 2048          # users per GPU per step: 16 tiles of 128 users for the tensor-core pre-filter
    k = 100
    users_local = (users + world - 1) // world
    model = build_model(users_local, items, dev, "fp32").eval()      # same seed on every rank: replicated item side
    scorer = ncf_b200.CatalogueScorer(model)
    g = torch.Generator().manual_seed(77 + rank)
    dev_b = [torch.randint(0, users_local, (n,), generator=g).to(dev) for _ in range(4)]
    host_b = [torch.randint(0, users_local, (n,), generator=g).pin_memory() for _ in range(4)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for s in range(args.warmup):
        scorer.topk(dev_b[s % 4], k)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = int(lib.ncf_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        scorer.topk(dev_b[s % 4], k)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.ncf_launch_count()) - l0
    t0 = time.perf_counter()
    for s in range(args.steps):
        idx, sc = scorer.topk(host_b[s % 4].to(dev, non_blocking=True), k)
        idx_h, sc_h = idx.cpu(), sc.cpu()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        pairs = world * n * items * args.steps
        pk = peaks()
        value = pairs / (ms_total / 1e3)
        tflops = value * 128 / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_score_rate(items, os.cpu_count() or 1)
        line = {"metric": "scoring_samples_per_s", "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: {desc}", "users_per_step_per_gpu": n, "items": items, "top_k": k,
                           "l2": f"folded item table {items * 260 / 1e6:.0f} MB streamed per user tile"},
                "users_per_s": world * n * args.steps / (ms_total / 1e3),
                "e2e": {"value": pairs / (e2e_ms / 1e3), "unit": "pairs/s", "h2d_bytes_per_step": n * 8,
                        "d2h_bytes_per_step": n * k * 12, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"kernel": "score_tc_kernel (tcgen05 bf16 upper-bound GEMM + exact fp32 re-scoring of survivors)"
                                       if scorer.img is not None and n >= scorer.TC_MIN_USERS else "score_topk_kernel (exact fp32)",
                             "bound": "tensor", "achieved": tflops, "peak": pk["bf16_tflops"],
                             "unit": "TFLOP/s", "frac": tflops / pk["bf16_tflops"], "traffic": None,
                             "peak_source": pk["source"] + " (bf16 burst); 128 flop per (user, item) pair; the kernel is bound by its "
                                            "per-tile epilogue (threshold scan of the TMEM accumulators), not by the tensor pipe"},
                "cpu_baseline": cpu}
        print(json.dumps(line))
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
    rate, ms = cpu_reference_rate(users, items, B_sample, args.steps, args.warmup, threads)
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": rate, "unit": "samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "interactions_per_step_per_gpu": B, "rows_per_interaction": S,
                       "table_update": "dense torch.optim.Adam on every row (the reference's own semantics)", "towers": "fp32",
                       "l2": "host DRAM; bounded sample of the step"},
            "cpu_baseline": {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{B_sample} of the {B} interactions of a step, {args.steps} steps"},
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

    if world > 1:
        # row-sharded tables: the model object only carries the replicated dense parameters
        from ncf_b200.sharding import ShardedNCFEngine
        model = build_model(1, 1, dev, precision)
        eng = ShardedNCFEngine(model, users, items, lr=1e-3, weight_decay=1e-5, table_mode=table_mode)
        # two sets of device staging buffers: while step s runs, the ids of step s+1 are already on the device, so
        # the engine can route them ahead (ShardedNCFEngine.train_step next_ids)
        dev_in = [[torch.empty(N, dtype=torch.long, device=dev), torch.empty(N, dtype=torch.long, device=dev),
                   torch.empty(N, dtype=torch.float32, device=dev)] for _ in range(2)]
        staged = {"slot": 0, "have": False}

        def stage(slot, batch):
            for d, h in zip(dev_in[slot], batch):
                d.copy_(h, non_blocking=True)

        def step_host(u, i, t, nxt=None):
            cur = staged["slot"]
            if not staged["have"]:
                stage(cur, (u, i, t))
            nids = None
            if nxt is not None:
                stage(cur ^ 1, nxt)
                nids = tuple(dev_in[cur ^ 1][:2])
            staged["slot"], staged["have"] = cur ^ 1, nxt is not None
            return float(eng.train_step(*dev_in[cur], next_ids=nids).item())
        eng.train_step_host = step_host
    else:
        if args.workload == "c3":
            raise SystemExit("workload c3 (169 GB of tables + Adam state) needs --gpus >= 2")
        model = build_model(users, items, dev, precision)
        eng = ncf_b200.NCFTrainEngine(model, lr=1e-3, weight_decay=1e-5, table_mode=table_mode, max_rows=N)
    nb = 4
    dev_batches = make_batches(users, items, B, nb, 1234 + rank, device=dev)
    host_batches = make_batches(users, items, B, nb, 4321 + rank, pin=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----
    for s in range(args.warmup):
        eng.train_step(*dev_batches[s % nb])
    barrier()
    if getattr(eng, "phase_ms", None):
        eng.phase_ms.clear()                 # NCF_SHARD_PROFILE: steady state only
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = int(lib.ncf_launch_count())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        if world > 1:      # sharded engine: the next batch's ids are known (input-pipeline look-ahead)
            eng.train_step(*dev_batches[s % nb], next_ids=dev_batches[(s + 1) % nb][:2] if s + 1 < args.steps else None)
        else:
            eng.train_step(*dev_batches[s % nb])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = int(lib.ncf_launch_count()) - launches0
    last_loss = float(eng.loss.item())

    # ---- end to end from pinned host buffers ("e2e") ----
    for s in range(3):
        eng.train_step_host(*host_batches[s % nb])
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        if world > 1:
            eng.train_step_host(*host_batches[s % nb], nxt=host_batches[(s + 1) % nb] if s + 1 < args.steps else None)
        else:
            eng.train_step_host(*host_batches[s % nb], next_batch=host_batches[(s + 1) % nb] if s + 1 < args.steps else None)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms_total, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms_total = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = world * N * args.steps / (ms_total / 1e3)
    e2e_value = world * N * args.steps / (e2e_ms_total / 1e3)

    if getattr(eng, "phase_ms", None):
        n = eng.phase_ms.pop("steps")
        sys.stderr.write(f"sharded phases, ms per step (rank {rank}): " + ", ".join(f"{k} {v / n:.3f}" for k, v in eng.phase_ms.items()) + "\n")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world > 1:
        line = {
            "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if precision == "bf16" else "fp32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "interactions_per_step_per_gpu": B,
                       "rows_per_interaction": S, "table_update": table_mode, "towers": precision,
                       "parallelism": f"tables row-sharded over {world} GPUs (all-to-all ids/rows/grads), towers "
                                      f"data-parallel (dense all-reduce)",
                       "l2": "per-step working set exceeds L2"},
            "interactions_per_s": value / S,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": N * (8 + 8 + 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms_total / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": None, "cpu_baseline": None, "last_loss": last_loss}
        # whole-step HBM view over all ranks (per-kernel pieces are measured by the N=1 run): algorithmic bytes
        # of the towers + embedding path per sample row, sparse table update
        pk = peaks()
        row_bytes = (BYTES_ATTN_FWD + BYTES_MLP_FWD + BYTES_MLP_BWD + BYTES_ATTN_BWD
                     + (BYTES_FWD_PER_INTERACTION + BYTES_BWD_PER_INTERACTION) / S)
        gbs = world * N * row_bytes / ms_step / 1e6
        line["roofline"] = {"kernel": "whole step, all ranks (NVLink exchange not counted)", "bound": "hbm", "achieved": gbs,
                            "peak": world * pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / (world * pk["hbm_gbs"]),
                            "traffic": None, "peak_source": pk["source"] + " (copy) x n_gpus",
                            "algorithmic_bytes_per_row": row_bytes}
        print(json.dumps(line))
        dist.destroy_process_group()
        return

    # ---- per-kernel roofline (rank 0, same inputs, CUDA events on the launch stream) ----
    # The step is timed piecewise through the C ABI: K1 (ncf_gather_ln_gmf_fwd), the whole forward
    # (ncf_forward), the tower backward alone (ncf_backward with emb_mode NONE), K6 (ncf_emb_bwd_adam_both)
    # and the dense-equivalent sweep; towers forward = forward - K1.
    pk = peaks()
    u, it, tg = dev_batches[0]
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
    ypm = torch.empty(N, 64, device=dev)
    yum = torch.empty(N, 64, device=dev)
    adam = _lib.AdamCfg()
    adam.lr, adam.beta1, adam.beta2, adam.eps, adam.weight_decay, adam.step = 1e-3, 0.9, 0.999, 1e-8, 1e-5, 7
    ews_bytes = int(lib.ncf_emb_bwd_workspace_bytes(N))
    ews = torch.empty(ews_bytes, dtype=torch.uint8, device=dev)
    dmf = torch.randn(N, device=dev) * 1e-6
    dx = torch.randn(N, 64, device=dev) * 1e-6

    k1_fn = lib.ncf_gather_ln_gmf_fwd_bf16 if precision == "bf16" else lib.ncf_gather_ln_gmf_fwd    # the variant the step runs

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
        _lib.check(lib.ncf_emb_bwd_adam_both(C.byref(adam), C.byref(tabs), _lib.ptr(flat), _lib.ptr(dgrad), _lib.ptr(u),
                                             _lib.ptr(it), N, _lib.ptr(dmf), _lib.ptr(dx), _lib.ptr(dx), _lib.ptr(ypm),
                                             _lib.ptr(yum), _lib.ptr(ews), ews_bytes, sptr))

    def sweep():
        adam.emb_mode = _lib.EMB_ADAM_DENSE_EQUIV
        _lib.check(lib.ncf_emb_adam_sweep(C.byref(adam), C.byref(tabs), sptr))

    # Stage calls run in place on the workspace a full forward left behind (include/ncf_b200.h); every stage moves
    # more bytes per call than the 126 MB L2 holds except K1/K6/sweep at this table size (noted below).
    fwd()
    k1_ms = time_kernel(k1, 20, sync)
    afwd_ms = time_kernel(attn_fwd, 10, sync)
    mfwd_ms = time_kernel(mlp_fwd, 10, sync)
    mbwd_ms = time_kernel(mlp_bwd, 10, sync)
    abwd_ms = time_kernel(attn_bwd, 10, sync)
    k6_ms = time_kernel(k6, 10, sync)
    sweep_ms = time_kernel(sweep, 10, sync) if table_mode == "fused_dense_equiv" else 0.0
    sus = pk["bf16_tflops_sustained"]
    hbm = pk["hbm_gbs"]

    def hbm_piece(ms, nbytes, **extra):
        d = {"ms": ms, "bound": "hbm", "achieved": nbytes / ms / 1e6, "unit": "GB/s", "peak": hbm,
             "frac": nbytes / ms / 1e6 / hbm, "algorithmic_bytes": nbytes}
        d.update(extra)
        return d

    def tensor_piece(ms, flop_per_row, bytes_per_row):
        # SURVEY 8d counts the towers against the tensor pipe; their arithmetic intensity (flop / byte of
        # activations a fused kernel has to move) is far below the bf16 ridge, so the HBM view is given as well
        tf = N * flop_per_row / ms / 1e9
        gb = N * bytes_per_row / ms / 1e6
        return {"ms": ms, "bound": "tensor", "achieved": tf, "unit": "TFLOP/s", "peak": sus, "frac": tf / sus,
                "algorithmic_flop": N * flop_per_row, "hbm_view": {"algorithmic_bytes": N * bytes_per_row, "achieved": gb,
                                                                   "unit": "GB/s", "peak": hbm, "frac": gb / hbm,
                                                                   "flop_per_byte": flop_per_row / bytes_per_row}}

    k1_bytes = B * BYTES_FWD_PER_INTERACTION
    k6_bytes = B * BYTES_BWD_PER_INTERACTION
    sweep_bytes = 2 * (users + items) * 1536
    tcp = precision == "bf16"
    kernels = {
        # SURVEY 8d "gather GB/s": every lookup counted (4 rows of 256 B per sample, no de-duplication), next to the
        # algorithmic figure (the S rows of an interaction share their user rows)
        "K1 gather_ln_gmf_fwd": hbm_piece(k1_ms, k1_bytes, gather_effective_gbs=4 * N * 256 / k1_ms / 1e6,
                                          gather_lookups=4 * N),
        "attention forward" + (" (attn_tc_fwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(afwd_ms, FLOP_ATTN_FWD, BYTES_ATTN_FWD),
        "MLP forward" + (" (mlp_tc_fwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(mfwd_ms, FLOP_MLP_FWD, BYTES_MLP_FWD),
        "MLP backward" + (" (head_bwd + mlp_tc_bwd + mlp_tc_wgrad kernels)" if tcp else " (fp32 kernels)"):
            tensor_piece(mbwd_ms, 2 * FLOP_MLP_FWD, BYTES_MLP_BWD),
        "attention backward" + (" (attn_tc_bwd_kernel)" if tcp else " (fp32 kernels)"): tensor_piece(abwd_ms, 3 * FLOP_ATTN_FWD, BYTES_ATTN_BWD),
        "K6 emb_bwd_adam_both (1 sort + segment-sum + apply, both sides)": hbm_piece(k6_ms, k6_bytes),
    }
    if sweep_ms:
        kernels["dense-equivalent Adam sweep"] = hbm_piece(sweep_ms, sweep_bytes,
                                                           note="tables that fit the 126 MB L2 read above the HBM copy peak")
    top = max(kernels, key=lambda k: kernels[k]["ms"])
    kt = kernels[top]
    traffic = None
    tpath = os.path.join(REPO, "profiles", "traffic.json")     # dram bytes per launch from the ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(args.workload, {}).get(top)
    emb_ms, emb_bytes = k1_ms + k6_ms, k1_bytes + k6_bytes
    pieces_ms = k1_ms + afwd_ms + mfwd_ms + mbwd_ms + abwd_ms + k6_ms + sweep_ms
    step_bytes = N * (BYTES_ATTN_FWD + BYTES_MLP_FWD + BYTES_MLP_BWD + BYTES_ATTN_BWD) + emb_bytes + (sweep_bytes if sweep_ms else 0)
    roofline = {"kernel": top, "bound": kt["bound"], "achieved": kt["achieved"], "peak": kt["peak"], "unit": kt["unit"],
                "frac": kt["frac"], "traffic": traffic,
                "peak_source": pk["source"] + (" (sustained bf16)" if kt["bound"] == "tensor" else " (copy)"),
                "embedding_path": {"ms": emb_ms, "achieved": emb_bytes / emb_ms / 1e6, "unit": "GB/s",
                                   "frac": emb_bytes / emb_ms / 1e6 / hbm},
                "whole_step_hbm_view": {"algorithmic_bytes": step_bytes, "ms": ms_step, "achieved": step_bytes / ms_step / 1e6,
                                        "unit": "GB/s", "frac": step_bytes / ms_step / 1e6 / hbm},
                "pieces_sum_ms": pieces_ms, "kernels": kernels}
    if "hbm_view" in kt:          # tower stages: SURVEY 8d counts them against the tensor pipe; their intensity says HBM
        roofline["hbm_view"] = kt["hbm_view"]

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        Bs = min(B, args.ref_batch)
        rate, ms = cpu_reference_rate(users, items, Bs, args.cpu_baseline_steps, 1, threads)
        cpu_baseline = {"value": rate, "unit": "samples/s", "cores": threads, "kind": "port", "ms_per_step": ms,
                        "sample": f"{Bs} of the {B} interactions of a step, {args.cpu_baseline_steps} steps"}

    line = {
        "metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "fp32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {desc}", "interactions_per_step_per_gpu": B, "rows_per_interaction": S,
                   "table_update": table_mode, "towers": precision,
                   "l2": f"inputs larger than L2: activations+ids of one step {N * 4316.8 / 1e6:.0f} MB algorithmic, "
                         f"{nb} rotating batches"},
        "interactions_per_s": value / S,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": N * (8 + 8 + 4), "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_ms_total / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
        "last_loss": last_loss,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
