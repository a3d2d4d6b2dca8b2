#!/usr/bin/env python
"""Copy the ncu evidence of the last `scripts/gpu_profile.sh` run from gpurun_out/ (scratch) into profiles/
(tracked): launch list, raw metric pages of the two full captures, and profiles/ncu_traffic.json, which bench.py
reports as roofline.traffic.  Usage: python scripts/refresh_profiles.py [round-tag, default r01]"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


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


# launch list: keep the csv rows only (drop ncu's banner lines)
with open(os.path.join(OUT, "launches.csv")) as f, open(os.path.join(PROF, f"{tag}_launches_bench.csv"), "w") as g:
    for line in f:
        if line.startswith('"'):
            g.write(line)
page(os.path.join(OUT, "prof_maxsim_tc.ncu-rep"), "raw", os.path.join(PROF, f"{tag}_maxsim_tc_full_raw.csv"))
page(os.path.join(OUT, "prof_maxsim_tc.ncu-rep"), "details", os.path.join(PROF, f"{tag}_maxsim_tc_full_details.txt"))
page(os.path.join(OUT, "prof_c3.ncu-rep"), "raw", os.path.join(PROF, f"{tag}_maxsim_tc_batched_raw.csv"))
d = first_row(os.path.join(PROF, f"{tag}_maxsim_tc_full_raw.csv"))
rd, wr = to_bytes(*d["dram__bytes_read.sum"]), to_bytes(*d["dram__bytes_write.sum"])
json.dump({
    "kernel": d["Kernel Name"][1],
    "bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
    "algorithmic_bytes_per_launch": 32768000000,
    "source": f"profiles/{tag}_maxsim_tc_full_raw.csv (ncu --set full, bench.py --steps 2 --warmup 3, 1M docs x 128 tokens)",
    "gpu_time_ms_under_ncu": float(d["gpu__time_duration.sum"][1]),
}, open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(PROF, "ncu_traffic.json")).read())
