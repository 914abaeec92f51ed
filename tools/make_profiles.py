#!/usr/bin/env python3
"""Turn one `ncu --set full` capture of a training step (+ the launch list of the same command) into the
tracked summaries under profiles/:
  <tag>_launches.csv / .md   per-kernel launch list (copy + table)
  <tag>_ncu_full.json        per-kernel digest: time, DRAM bytes, pipes, occupancy, top stall reasons
  traffic.json               DRAM bytes per step of the stage bench.py reports as `roofline.kernel`
usage: python tools/make_profiles.py <tag> <workload> gpurun_out/launches.csv gpurun_out/step.ncu-rep "<command>" [full|metrics]

The .ncu-rep may also come from a capture with an explicit --metrics list (the METRICS below; ~1 minute of GPU time
instead of ~3.5 for --set full): the digest then has no stall reasons - say so in the note of the written JSON
(profiles/r01_final_ncu_metrics.json was made that way).
"""
import collections
import csv
import io
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tools"))

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
# which kernels make up the stages bench.py times through the C ABI
STAGES = {
    "K1 gather_ln_gmf_fwd": ["gather_ln_gmf_fwd_kernel", "gather_ln_gmf_fwd_grouped_kernel"],
    "attention forward (attn_tc_fwd_kernel)": ["attn_tc_fwd_kernel"],
    "MLP forward (mlp_tc_fwd_kernel)": ["mlp_tc_fwd_kernel", "mlp_tc_fwd2_kernel"],
    "MLP backward (head_bwd + mlp_tc_bwd + mlp_tc_wgrad kernels)": ["head_bwd", "mlp_tc_bwd_kernel", "mlp_tc_bwd2_kernel",
                                                                    "mlp_tc_wgrad_kernel", "mlp_wgrad_reduce_kernel"],
    "attention backward (attn_tc_bwd_kernel)": ["attn_tc_bwd_kernel", "attn_wgrad_reduce_kernel"],
    "K6 emb_bwd_adam_both (1 sort + segment-sum + apply, both sides)": ["emb_bwd_phase1", "emb_bwd_phase2_kernel", "DeviceRadixSort",
                                                                        "gather_sorted_kernel", "ids_to_keys2_kernel"],
    "dense-equivalent Adam sweep": ["emb_adam_sweep_kernel"],
}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def main(tag, workload, launches, rep, command, kind="full"):
    how = "ncu --set full" if kind == "full" else "ncu with an explicit metric list (time, DRAM bytes, pipes, occupancy; no stall reasons)"
    prof = os.path.join(REPO, "profiles")
    os.makedirs(prof, exist_ok=True)
    shutil.copy(launches, os.path.join(prof, f"{tag}_launches.csv"))
    table = subprocess.run([sys.executable, os.path.join(REPO, "tools", "summarize_launches.py"), launches],
                           capture_output=True, text=True).stdout
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    digest, per_kernel = [], collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for r in rows[2:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("ncf::", "")
        d = {"kernel": name}
        for m in METRICS:
            if m in hdr:
                v = num(r[hdr.index(m)])
                u = units[hdr.index(m)]
                if v is not None and m.startswith("dram__bytes"):
                    v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
                if v is not None and m == "gpu__time_duration.sum":
                    v *= {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
                d[m] = v
        st = sorted(((num(r[hdr.index(h)]) or 0.0, h.split("stalled_")[1].split("_per_")[0]) for h in stall), reverse=True)
        d["top_stalls_per_issue"] = {k: round(v, 2) for v, k in st[:5]}
        digest.append(d)
        a = per_kernel[name]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum") or 0
        a[2] += d.get("dram__bytes_read.sum") or 0
        a[3] += d.get("dram__bytes_write.sum") or 0
    with open(os.path.join(prof, f"{tag}_ncu_full.json"), "w") as f:
        json.dump({"command": command, "note": f"one training step captured with {how} --clock-control none; "
                   "cold-cache, serialised kernels: compare shares and bytes, not absolute times", "kernels": digest}, f, indent=1)
    stages = {}
    for label, pats in STAGES.items():
        b = sum(v[2] + v[3] for k, v in per_kernel.items() if any(p in k for p in pats))
        if b:
            stages[label] = b
    tpath = os.path.join(prof, "traffic.json")
    traffic = {}
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    traffic[workload] = stages
    traffic.setdefault("_source", {})[workload] = f"profiles/{tag}_ncu_full.json (dram__bytes_read.sum + dram__bytes_write.sum per step)"
    with open(tpath, "w") as f:
        json.dump(traffic, f, indent=1)
    with open(os.path.join(prof, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: launch list, workload {workload}\n\nCommand: `{command}`\n\n"
                "Cold-cache, serialised per-launch times (`--clock-control none`): read the shares, not the absolute times.\n\n")
        f.write(table)
        f.write(f"\n## DRAM traffic per step by stage ({how}, same command)\n\n| stage | kernels | time us | DRAM read MB | DRAM write MB |\n|---|---|---:|---:|---:|\n")
        for label, pats in STAGES.items():
            ks = {k: v for k, v in per_kernel.items() if any(p in k for p in pats)}
            if ks:
                f.write(f"| {label} | {', '.join(sorted(ks))[:80]} | {sum(v[1] for v in ks.values()):.1f} | "
                        f"{sum(v[2] for v in ks.values()) / 1e6:.1f} | {sum(v[3] for v in ks.values()) / 1e6:.1f} |\n")
    print(json.dumps(stages, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:7])
