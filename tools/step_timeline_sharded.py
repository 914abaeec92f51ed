#!/usr/bin/env python3
"""Kernel timeline of one SHARDED training step on rank 0 (torch.profiler / CUPTI activity records; analysis only - numbers
taken under a profiler are never bench values): start, duration and stream of every kernel between two consecutive pull kernels.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/step_timeline_sharded.py [workload] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from torch.profiler import ProfilerActivity, profile
    from ncf_b200.sharding import ShardedNCFEngine
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    users, items, B = bench.WORKLOADS[wl][:3]
    model = bench.build_model(1, 1, dev, "bf16")
    mode = "fused_sparse" if wl in ("c3", "c3shard") else "fused_dense_equiv"
    eng = ShardedNCFEngine(model, users, items, lr=1e-3, weight_decay=1e-5, table_mode=mode)
    nb = 4
    batches = bench.make_batches(users, items, B, nb, 1234 + rank, device=dev)

    def step(s, last=False):
        eng.train_step(*batches[s % nb], next_ids=None if last else batches[(s + 1) % nb][:2])
    for s in range(8):
        step(s)
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for s in range(steps):
            step(s, last=s == steps - 1)
        torch.cuda.synchronize()
    if rank == 0:
        ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
        ev.sort(key=lambda e: e.time_range.start)
        pulls = [i for i, e in enumerate(ev) if "pull_rows" in e.name]
        # one step = from the first pull kernel of step k to the first pull kernel of step k + 1 (two pull kernels per step)
        a, b = pulls[-4], pulls[-2]
        t0 = ev[a].time_range.start
        print(f"{'start us':>9s} {'dur us':>8s}  kernel (rank 0, one step between two pulls, world {world}, {wl})")
        for e in ev[a:b]:
            s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
            print(f"{s:9.1f} {d:8.1f}  {e.name[:100]}")
        print(f"step: {ev[b].time_range.start - t0:.1f} us between the pulls of two consecutive steps")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
