#!/usr/bin/env python
"""DEVELOPER EXPERIMENT: GPU-side spans of the phases of one device-resident step (events recorded on the library's own
stream between the host calls) next to the host wall clock of the same phases and the kernel times the library accounts."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
from kmerjs_b200 import synth  # noqa: E402
from kmerjs_b200.context import Context  # noqa: E402
from kmerjs_b200.counts import Counts  # noqa: E402
from kmerjs_b200.matching import Match  # noqa: E402

args = argparse.Namespace(config="c3", reads=0, templates=-1, cpu_reads=0, score_mode="auto")
cfg = bench.resolve_config(args, 1)
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
ctx = Context(0, stream=stream.cuda_stream)
w = synth.Workload(n_reads=cfg["reads_per_gpu"], genome_len=cfg["genome_len"], seed=bench.SEED, first_read=0,
                   sub_rate=cfg["sub_rate"], ctx=ctx)
tdb = bench.build_db(cfg, w.genome_host())
tdb.device(ctx)
hint = 1 << 18
names = ["create+add_device", "finish", "first_match", "wta_rows", "free"]


def step(rec):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    t = [0.0] * 6
    ev[0].record(stream); t[0] = time.perf_counter()
    c = Counts(cfg["prefix"], cfg["k"], cfg["step"], capacity_hint=hint, ctx=ctx)
    c.add_device(w.fastq_ptr, w.n_bytes, final=True)
    ev[1].record(stream); t[1] = time.perf_counter()
    c.finish()
    ev[2].record(stream); t[2] = time.perf_counter()
    m = Match(c, tdb)
    ev[3].record(stream); t[3] = time.perf_counter()
    rows, _ = m.all_rows()
    ev[4].record(stream); t[4] = time.perf_counter()
    m.free(); c.free()
    ev[5].record(stream); t[5] = time.perf_counter()
    torch.cuda.synchronize(dev)
    if rec is not None:
        rec.append(([ev[i].elapsed_time(ev[i + 1]) for i in range(5)], [(t[i + 1] - t[i]) * 1e3 for i in range(5)]))


for _ in range(4):
    step(None)
ctx.enable_timers(True); ctx.reset_timers()
rec = []
for _ in range(10):
    step(rec)
scan_ms, scan_n, _ = ctx.scan_kernel_stats()
print(f"kernels accounted per step: extraction {scan_ms:.3f} ms (resolve part {ctx.verify_kernel_ms():.3f})")
for i, n in enumerate(names):
    g = sorted(r[0][i] for r in rec)[len(rec) // 2]
    h = sorted(r[1][i] for r in rec)[len(rec) // 2]
    print(f"{n:20s} GPU span between the events {g:7.3f} ms   host wall {h:7.3f} ms")
print("total GPU span", round(sum(sorted(sum(r[0]) for r in rec)[len(rec) // 2:len(rec) // 2 + 1]), 3), "ms; host", round(sorted(sum(r[1]) for r in rec)[len(rec) // 2], 3), "ms")
