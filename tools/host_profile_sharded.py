#!/usr/bin/env python3
"""Where the HOST spends its time in one sharded end-to-end step (torchrun, >= 2 GPUs): cProfile of rank 0 over
`ShardedNCFEngine.train_step_host`, printed by internal time.  usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/host_profile_sharded.py [steps]"""
import cProfile
import io
import os
import pstats
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from ncf_b200.sharding import ShardedNCFEngine
    users, items = bench.WORKLOADS["c2"][:2]
    B = 65536
    model = bench.build_model(1, 1, dev, "bf16")
    eng = ShardedNCFEngine(model, users, items, lr=1e-3, weight_decay=1e-5, table_mode="fused_dense_equiv")
    nb = 4
    host = bench.make_batches(users, items, B, nb, 4321 + rank, pin=True)
    for s in range(10):
        eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb])
    torch.cuda.synchronize()
    dist.barrier()
    # pass 0: wall time per step of three loops: the end-to-end step; the same without waiting for the loss; the
    # device-resident step with look-ahead (bench.py's `value` loop)
    devb = bench.make_batches(users, items, B, nb, 1234 + rank, device=dev)

    def loop(fn, n):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for s in range(n):
            fn(s)
        torch.cuda.synchronize()
        return 1e3 * (time.perf_counter() - t0) / n
    res = {}
    res["end to end"] = loop(lambda s: eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb]), steps)
    eng._debug_no_loss_wait = True
    res["end to end, loss not waited for"] = loop(lambda s: eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb]), steps)
    eng._debug_no_loss_wait = False
    res["device-resident + look-ahead"] = loop(lambda s: eng.train_step(*devb[s % nb], next_ids=devb[(s + 1) % nb][:2]), steps)
    def dev_early(s):
        eng._want_loss_event = True
        try:
            eng.train_step(*devb[s % nb], next_ids=devb[(s + 1) % nb][:2])
        finally:
            eng._want_loss_event = False
        eng.__dict__.pop("_loss_event", None)
        eng.__dict__.pop("_loss_from_counts", None)
    res["device-resident + early loss exchange"] = loop(dev_early, steps)
    side = torch.cuda.Stream(device=dev)
    scratch = [torch.empty_like(t) for t in devb[0]]

    def dev_h2d(s):
        with torch.cuda.stream(side):
            for d, h in zip(scratch, host[s % nb]):
                d.copy_(h, non_blocking=True)
        eng.train_step(*devb[s % nb], next_ids=devb[(s + 1) % nb][:2])
    res["device-resident + unrelated H2D copies"] = loop(dev_h2d, steps)
    res["end to end (again)"] = loop(lambda s: eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb]), steps)
    if rank == 0:
        for k, v in res.items():
            print(f"{k:36s} {v:.3f} ms per step (wall)")
    # pass 1: a timeline of the host inside one step (perf_counter stamps around the engine's phases, no profiler)
    stamps = {}
    t_step = [0.0]

    def wrap(obj, name, label=None):
        fn = getattr(obj, name)

        def timed(*a, **k):
            t_in = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                t_out = time.perf_counter()
                e = stamps.setdefault(label or name, [0.0, 0.0, 0])
                e[0] += t_in - t_step[0]
                e[1] += t_out - t_in
                e[2] += 1
        setattr(obj, name, timed)
    for name in ("_route", "_begin_count_gather", "_fill_plan", "phase_pull", "phase_forward_backward", "phase_owner_update",
                 "phase_dense_adam", "check_status", "_adopt"):
        wrap(eng, name)
    wrap(torch.cuda.Event, "synchronize", "Event.synchronize")
    # device-side latency of the early loss exchange: from "loss kernel done" (main stream) to "exchanged" (collective stream)
    pairs = []
    early = eng._early_loss

    def early_timed():
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(torch.cuda.current_stream(dev))
        early()
        if getattr(eng, "_coll", None) is not None:
            b.record(eng._coll)
            pairs.append((a, b))
    eng._early_loss = early_timed
    step_ev = []
    t0 = time.perf_counter()
    for s in range(steps):
        t_step[0] = time.perf_counter()
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(dev))
        step_ev.append(e)
        eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb])
    torch.cuda.synchronize()
    dt0 = time.perf_counter() - t0
    eng._early_loss = early
    if rank == 0 and pairs:
        lat = sorted(a.elapsed_time(b) for a, b in pairs)
        at = sorted(e.elapsed_time(a) for e, (a, _) in zip(step_ev, pairs))
        print(f"early loss exchange: {1e3 * lat[len(lat) // 2]:.0f} us median device latency behind the loss kernel "
              f"(max {1e3 * lat[-1]:.0f}); the loss kernel ends {1e3 * at[len(at) // 2]:.0f} us (median) after the step's first "
              f"enqueue reaches the device")
    if rank == 0:
        print(f"{1e3 * dt0 / steps:.3f} ms per step (wall, stamps only), {steps} steps, world {world}")
        print("phase: mean entry time after the step's start (us) / mean duration (us) / calls per step")
        for k, (a, d, n) in sorted(stamps.items(), key=lambda kv: kv[1][0] / kv[1][2]):
            print(f"  {k:28s} {1e6 * a / n:8.1f} {1e6 * d / n:8.1f} {n / steps:5.1f}")
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    if rank == 0:
        pr.enable()
    for s in range(steps):
        eng.train_step_host(*host[s % nb], next_batch=host[(s + 1) % nb])
    if rank == 0:
        pr.disable()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        out = io.StringIO()
        pstats.Stats(pr, stream=out).sort_stats("tottime").print_stats(40)
        print(f"{1e3 * dt / steps:.3f} ms per step (wall, with the profiler on rank 0), {steps} steps, world {world}")
        print(out.getvalue())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
