#!/usr/bin/env python3
"""Run bench.verify_sharded alone under torchrun (the real NCCL / IPC path against the single-GPU engine), `repeats` times,
and print the detail.
usage: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/verify_sharded.py [repeats] [steps]"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    repeats = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    bad = 0
    for r in range(repeats):
        ok, detail = bench.verify_sharded(world, rank, dev, steps)
        if rank == 0:
            bad += not ok
            a, b = detail["adam_eps_1e-4"], detail["adam_eps_1e-8"]
            print(f"repeat {r}: parity {ok}  eps 1e-4: loss {a['max_abs_loss_diff']:.1e} MF {a['max_abs_diff_mf_tables']:.1e} "
                  f"MLP {a['max_abs_diff_mlp_tables']:.1e} MLP rows off {a['frac_mlp_rows_off_by_more_than_1e-7']:.1e} | eps 1e-8: loss {b['max_abs_loss_diff']:.1e} "
                  f"tables {b['max_abs_table_diff']:.1e} frac>8e-6 {b['frac_table_elements_off_by_more_than_8e-6']:.1e}", flush=True)
    if rank == 0:
        print(f"verify: {bad} of {repeats} repeats failed")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
