#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary kept under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__cluster_max_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct"]


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def main(rep, out):
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, vals = raw[0], raw[1], raw[2]
    lines = [f"# {rep}", f"kernel: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'}", ""]
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            lines.append(f"{k} = {vals[i]} {units[i]}")
    for h, u, v in zip(hdr, units, vals):
        if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.2:
            lines.append(f"{h} = {v}")
    src = list(csv.reader(io.StringIO(run([rep, "--page", "source", "--csv"]))))
    shdr = src[1]
    ix = {h: i for i, h in enumerate(shdr)}
    data = [r for r in src[2:] if len(r) == len(shdr)]
    tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
    stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    lines += ["", f"top instructions by warp-stall samples (total {tot}):"]
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:20]:
        s = int(r[ix["# Samples"]])
        dom = sorted(((int(r[ix[h]]), h) for h in stalls), reverse=True)[0]
        lines.append(f"{100 * s / tot:5.1f}%  {r[ix['Source']].strip()[:72]:72s} {dom[1]}")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:30]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
