#!/usr/bin/env python
"""Copy the ncu evidence of the last `scripts/gpu_profile_r02.sh` run from gpurun_out/ (scratch) into profiles/
(tracked): launch list, raw metric pages of the full captures, and profiles/ncu_traffic.json, which bench.py
reports as roofline.traffic.  Usage: python scripts/refresh_profiles.py [round-tag, default r02]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"


def page(rep, which, dst):
    with open(dst, "w") as f:
        subprocess.run(["ncu", "-i", rep, "--page", which] + (["--csv"] if which == "raw" else []), stdout=f,
                       stderr=subprocess.DEVNULL, check=True)


def first_row(path):
    rows = list(csv.reader(open(path)))
    return {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}


def to_bytes(unit, value):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[unit]
    return float(value) * mult


KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed_op_shared_ld.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]

with open(os.path.join(OUT, f"{tag}_launches.csv")) as f, open(os.path.join(PROF, f"{tag}_launches_bench.csv"), "w") as g:
    for line in f:
        if line.startswith('"'):
            g.write(line)
summary = {}
for rep, name in ((f"{tag}_prof_maxsim_dm", "maxsim_dm_full"), (f"{tag}_prof_maxsim_qm", "maxsim_qm_full"),
                  (f"{tag}_prof_c3_pair", "maxsim_tc_batched"), (f"{tag}_prof_topk_stream", "topk_stream")):
    src = os.path.join(OUT, rep + ".ncu-rep")
    if not os.path.exists(src):
        print("missing", src)
        continue
    raw = os.path.join(PROF, f"{tag}_{name}_raw.csv")
    page(src, "raw", raw)
    if name == "maxsim_dm_full":
        page(src, "details", os.path.join(PROF, f"{tag}_{name}_details.txt"))
    d = first_row(raw)
    summary[name] = {k: d[k] for k in KEEP if k in d}
json.dump(summary, open(os.path.join(PROF, f"{tag}_ncu_key_metrics.json"), "w"), indent=1)
if "maxsim_dm_full" in summary:
    d = first_row(os.path.join(PROF, f"{tag}_maxsim_dm_full_raw.csv"))
    rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
    json.dump({
        "kernel": d["Kernel Name"][1],
        "bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "algorithmic_bytes_per_launch": 32768000000,
        "source": f"profiles/{tag}_maxsim_dm_full_raw.csv (ncu --set full, bench.py --steps 2 --warmup 3 --no-secondary, 1M docs x 128 tokens)",
        "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"][1]),
    }, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    print(open(os.path.join(PROF, "ncu_traffic.json")).read())
print(json.dumps(summary, indent=1)[:3000])
