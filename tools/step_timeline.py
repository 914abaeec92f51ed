#!/usr/bin/env python3
"""Kernel timeline of one single-GPU training step (torch.profiler / CUPTI activity records; analysis only - numbers taken
under a profiler are never bench values): start, duration, stream and the idle gap before every kernel of the main stream.
usage: python tools/step_timeline.py [workload] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    import ncf_b200
    wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    users, items, B = bench.WORKLOADS[wl][:3]
    dev = torch.device("cuda", 0)
    model = bench.build_model(users, items, dev, "bf16")
    eng = ncf_b200.NCFTrainEngine(model, lr=1e-3, weight_decay=1e-5, table_mode="fused_dense_equiv")
    batches = bench.make_batches(users, items, B, 4, 1234, device=dev)
    for s in range(8):
        eng.train_step(*batches[s % 4])
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for s in range(steps):
            eng.train_step(*batches[s % 4])
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ev.sort(key=lambda e: e.time_range.start)
    if not ev:
        print("no device activity records")
        return
    # the last step: from its first gather kernel on
    starts = [i for i, e in enumerate(ev) if "gather_ln_gmf_fwd" in e.name]
    first = starts[-1]
    # the auxiliary stream's kernels of that step may start before K1's record: include what overlaps
    t0 = ev[first].time_range.start
    step = [e for e in ev if e.time_range.end >= t0]
    main_stream = None
    for e in step:
        if "gather_ln_gmf_fwd" in e.name:
            main_stream = getattr(e, "stream", None)
    print(f"{'start us':>9s} {'dur us':>8s} {'gap us':>7s} stream  kernel")
    last_end = {}
    busy_main = 0.0
    for e in step:
        st = getattr(e, "stream", None)
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = s - last_end[st] if st in last_end else 0.0
        last_end[st] = s + d
        if st == main_stream:
            busy_main += d
        print(f"{s:9.1f} {d:8.1f} {gap:7.1f} {str(st):>6s}  {e.name[:90]}")
    end = max(e.time_range.end for e in step) - t0
    print(f"step: {end:.1f} us from K1's start to the last kernel's end; main stream busy {busy_main:.1f} us")


if __name__ == "__main__":
    main()
