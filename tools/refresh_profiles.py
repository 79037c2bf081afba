#!/usr/bin/env python
"""Regenerate the tracked summaries under profiles/ from the scratch files a GPU run left in gpurun_out/:
launch list + shares (tools/gpu_round.sh, NCU=1), kernel table + scan-kernel summary + traffic.json
(tools/gpu_prof.sh), bench lines.  usage: tools/refresh_profiles.py [rNN]"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
G, P = "gpurun_out", "profiles"


def launches():
    src = os.path.join(G, "launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(P, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3, "s": v * 1e6}.get(r[ui].strip(), v)     # -> us
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    cmd = "python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
    log = os.path.join(G, "ncu_launches.log")
    whole = sum(tot.values())
    with open(os.path.join(P, f"{tag}_launch_shares.txt"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 120: {cmd} --no-file-leg (10 M reads, 3.46 GB per launch)\n")
        f.write(f"# {sum(cnt.values())} launches captured (includes workload generation, warm-up and the first steps); "
                "cold-cache, serialised: compare shares\n")
        for name, t in tot.most_common():
            f.write(f"{name:<62} n={cnt[name]:4d} total_us={t:10.1f} share={100 * t / whole:5.1f}% avg_us={t / cnt[name]:8.1f}\n")
        # share of the extraction kernels inside one step = the kernels a step launches
        step = [n for n in tot if not any(x in n for x in ("synth", "genome", "db_build"))]
        st = sum(tot[n] for n in step)
        ext = sum(tot[n] for n in step if any(x in n for x in ("kj_warp_filter", "kj_resolve", "kj_scan", "DeviceScan")))
        f.write(f"# extraction (kj_warp_filter + tile scan + kj_resolve; includes the small scans of the scoring path) share of the step kernels: {100 * ext / st:.1f}%\n")


def full():
    rep = os.path.join(G, "prof_scan.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]

    def col(r, name):
        return float(r[hdr.index(name)].replace(",", ""))
    units = dict(zip(hdr, rows[1]))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    tus = {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}[units["gpu__time_duration.sum"]]
    per = collections.defaultdict(list)
    lines = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")].strip()
        rd = col(r, "dram__bytes_read.sum") * scale[units["dram__bytes_read.sum"]]
        wr = col(r, "dram__bytes_write.sum") * scale[units["dram__bytes_write.sum"]]
        t = col(r, "gpu__time_duration.sum") * tus
        lines.append(f"{name} time_us {t:.1f} warp_inst {col(r, 'smsp__inst_executed.sum'):.0f} issue_active_pct "
                     f"{col(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} dram_read_MB {rd / 1e6:.1f} "
                     f"dram_write_MB {wr / 1e6:.1f} regs {col(r, 'launch__registers_per_thread'):.0f} grid "
                     f"{col(r, 'launch__grid_size'):.0f} block {col(r, 'launch__block_size'):.0f}")
        per[name.split("(")[0]].append(rd + wr)
    plain = None
    for cand in ("bench.json", "bench_quick.json", "plain_short.json"):
        q = os.path.join(G, cand)
        if os.path.exists(q) and os.path.getsize(q):
            plain = json.loads([l for l in open(q).read().strip().splitlines() if l.startswith("{")][-1])
            break
    nbytes = plain["fastq_bytes_per_gpu"]
    with open(os.path.join(P, f"{tag}_ncu_full_kernels.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:kj_warp_filter|kj_resolve, python bench.py --reads "
                f"{plain['config']['reads_per_gpu']} --steps 2 --warmup 3 ({nbytes / 1e6:.0f} MB FASTQ per launch); per launch\n")
        f.write("\n".join(lines) + "\n")
    traffic = sum(sum(v) / len(v) for v in per.values())
    alg = plain["roofline"]["algorithmic_bytes_per_launch"]
    json.dump({"source": f"profiles/{tag}_ncu_full_kernels.txt (ncu --set full at the bench's own size, {nbytes / 1e6:.0f} MB per launch: "
                         + "; ".join(f"{k.strip()} {sum(v) / len(v) / 1e6:.1f} MB" for k, v in per.items()) + ")",
               "dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": alg,
               "dram_bytes_per_input_byte": traffic / nbytes}, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    out = subprocess.run([sys.executable, "tools/ncu_summary.py", rep, str(nbytes)], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{tag}_scan_kernel_ncu_summary.txt"), "w").write(out)


def bench():
    for src, dst in (("bench.json", f"{tag}_bench_n1.json"), ("bench_ref.json", f"{tag}_bench_reference.json"),
                     ("bench_2gpu.json", f"{tag}_bench_n2.json")):
        p = os.path.join(G, src)
        if os.path.exists(p) and os.path.getsize(p):
            line = [l for l in open(p).read().strip().splitlines() if l.startswith("{")][-1]
            json.loads(line)
            open(os.path.join(P, dst), "w").write(line + "\n")


if __name__ == "__main__":
    launches()
    full()
    bench()
    print("profiles refreshed:", sorted(os.listdir(P)))
