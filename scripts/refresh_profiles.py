#!/usr/bin/env python
"""Copies the evidence a `scripts/gpu_r2_step21.sh` run merged into gpurun_out/ to profiles/ and derives
profiles/r2_launches_summary.csv (per-kernel totals of the ncu launch list) and profiles/r2_ncu_traffic.json
(dram bytes per (row, frame) of the shipped 416-row recurrence launch, what bench.py's roofline.traffic is computed from)."""
import collections
import csv
import glob
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def main():
    for old in glob.glob(os.path.join(DST, "r2_ncu_*.txt")):
        os.remove(old)
    for f in glob.glob(os.path.join(SRC, "r2_ncu_*.txt")) + [os.path.join(SRC, n) for n in (
            "r2_launches.csv", "r2_frontend_microbench.txt", "r2_gemm_microbench.txt")]:
        if os.path.isfile(f):
            shutil.copy(f, DST)
    # launch list -> per-kernel totals
    rows = [r for r in csv.reader(open(os.path.join(SRC, "r2_launches.csv"))) if len(r) > 10]
    hdr = rows[0]
    name, val, unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        k = re.sub(r"^void |\(.*$", "", r[name]).replace("tssep::", "")
        v = float(r[val].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[unit], 1e-6)
        n, t = tot.get(k, (0, 0.0))
        tot[k] = (n + 1, t + v)
    total = sum(t for _, t in tot.values())
    with open(os.path.join(DST, "r2_launches_summary.csv"), "w") as f:
        f.write("kernel,launches,total_ms,share\n")
        for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            f.write(f'"{k}",{n},{t:.3f},{t / total:.4f}\n')
    # dram traffic of the shipped 416-row launch
    txt = open(os.path.join(SRC, "r2_ncu_rec_ts_416rows.txt")).read()

    def metric(key):
        m = re.search(rf"{re.escape(key)} = ([\d.]+) (\w+)", txt)
        return float(m.group(1)) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[m.group(2)]

    rd, wr = metric("dram__bytes_read.sum"), metric("dram__bytes_write.sum")
    rows_, frames, Up = 416, 2000, 304
    algo = rows_ * frames * (8 * Up * 2 + 2 * Up * 2)   # G read once (bf16), H written once (bf16)
    kernel = re.search(r"kernel: void (.*?)\(", txt).group(1)
    json.dump({
        "kernel": kernel,
        "capture": "profiles/r2_ncu_rec_ts_416rows.txt (ncu --set full --clock-control none, scripts/profile_rec.py "
                   "--rows 416 --clusters 32 --tiles 2 --subs 2 --frames 2000)",
        "rows": rows_, "frames": frames, "dram_bytes_read": rd, "dram_bytes_write": wr, "algorithmic_bytes": algo,
        "ratio": (rd + wr) / algo, "bytes_per_row_frame": (rd + wr) / (rows_ * frames),
        "note": "dram__bytes_read.sum + dram__bytes_write.sum of the shipped 32-rows-per-cluster launch (two sub-batches of "
                "16), scaled per (row, frame); bench.py multiplies by the rows and frames of its own launches",
    }, open(os.path.join(DST, "r2_ncu_traffic.json"), "w"), indent=1)
    print(open(os.path.join(DST, "r2_launches_summary.csv")).read())
    print(open(os.path.join(DST, "r2_ncu_traffic.json")).read())


if __name__ == "__main__":
    main()
