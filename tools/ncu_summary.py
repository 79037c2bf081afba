#!/usr/bin/env python
"""Summarise an .ncu-rep of the scan kernel: key metrics, instruction/sample share per source region,
stall reasons.  usage: tools/ncu_summary.py gpurun_out/prof_scan.ncu-rep [bytes_per_launch]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
nbytes = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]
d = dict(zip(hdr, rows[2]))
print("kernel:", d.get("Kernel Name"))
for w in want:
    if w in d:
        print(f"  {w} = {d[w]} {units[hdr.index(w)]}")
inst_total = float(d["smsp__inst_executed.sum"].replace(",", ""))
if nbytes:
    print(f"  thread-instructions per byte = {inst_total * 32 / nbytes:.2f}  (per 16-byte chunk {inst_total * 32 / nbytes * 16:.0f})")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
cur, h = None, None
inst, smp, stall = collections.Counter(), collections.Counter(), collections.Counter()
by_line_i, by_line_s, text = collections.Counter(), collections.Counter(), {}
kernels = 0
seen_files = set()
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        if cur in seen_files:
            break               # second kernel instance: same files again
        seen_files.add(cur)
        continue
    if len(r) > 5 and r[0] == "Line No":
        h = r
        continue
    if h and len(r) == len(h) and r[0] != "":
        try:
            line = int(r[0])
        except ValueError:
            continue
        dd = dict(zip(h[4:], r[4:]))
        key = (cur, line)
        text[key] = r[1].strip()[:80]
        by_line_i[key] += int(dd["Instructions Executed"])
        by_line_s[key] += int(dd["# Samples"])
        for k, v in dd.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                stall[k] += int(v)
ti, ts = sum(by_line_i.values()), sum(by_line_s.values())
print("top source lines (instr share, sample share):")
for key, v in by_line_i.most_common(28):
    print(f"  {v / ti * 100:5.1f}% inst {by_line_s[key] / ts * 100:5.1f}% smp  {key[0]}:{key[1]}  {text[key]}")
print("top lines by samples:")
for key, v in by_line_s.most_common(8):
    print(f"  {v / ts * 100:5.1f}% smp {by_line_i[key] / ti * 100:5.1f}% inst  {key[0]}:{key[1]}  {text[key]}")
t = sum(stall.values())
print("stalls:", {k: round(v / t * 100, 1) for k, v in stall.most_common(9)})
