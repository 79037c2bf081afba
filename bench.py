#!/usr/bin/env python
"""bench.py -- Gbases/s of the kmerjs hot path (FASTQ -> k-mer counts + template scores) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the whole path over one batch of synthetic reads: count (scan + extract +
hash count), first match against the template DB resident in HBM, winner-takes-all rows.
Workload (BASELINE.json configs[2]): 10 M x 150 bp Illumina-shaped reads per GPU, prefix ATGAC,
k = 16, step = 1, sampled from a 5 Mbp random genome; template DB = that genome, mutated relatives
and decoys.  With N > 1 every rank takes its own 10 M reads (weak scaling), k-mers travel to
their owner GPU in one NCCL all-to-all, per-template vectors are all-reduced.

`value`      device-resident throughput (inputs already in HBM when the timed region starts)
`e2e`        the same path through the C ABI with HOST buffers: pinned FASTQ bytes in, H2D staging,
             k-mer map arrays and rows back out, all inside the timed region
`roofline`   the scan kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
`cpu_baseline` / --impl reference: the CPU restatement of the reference algorithm (oracle/, one
             thread -- the reference is a single Node.js event loop) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gbases/s FASTQ->k-mer counts+template scores"
UNIT = "Gbases/s"
PREFIX, K, STEP = b"ATGAC", 16, 1
GENOME_LEN = 5_000_000
N_TEMPLATES = 32
SEED = 0x6B6D6572
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=10_000_000, help="reads per GPU (150 bp)")
    ap.add_argument("--cpu-reads", type=int, default=1_000_000, help="reads in the CPU-baseline sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--score-mode", default="auto", choices=["auto", "gather", "reduce"],
                    help="N>1: all-gather the matched set once (gather) or all-reduce the score vector every round (reduce)")
    ap.add_argument("--trace", action="store_true", help="print host-side phase timings of one device step to stderr")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"synthetic {args.reads // 1_000_000 if args.reads >= 1_000_000 else args.reads / 1e6:g}M x 150bp "
                        f"Illumina-shaped reads per GPU, prefix ATGAC, k=16, step=1 (BASELINE configs[2]); "
                        f"template DB of {N_TEMPLATES} synthetic templates over a {GENOME_LEN // 1_000_000} Mbp genome",
            "reads_per_gpu": args.reads, "read_len": 150, "prefix": "ATGAC", "k": K, "step": STEP,
            "templates": N_TEMPLATES, "l2": "inputs larger than L2 (3.46 GB FASTQ per GPU vs 126 MB)",
            "sharding": ("whole records per rank; owner all-to-all of counted k-mers; scoring: " +
                         ("per-round all-reduce of the template sums" if args.score_mode == "reduce" else
                          "one all-gather of the matched entries, winner-takes-all replicated")) if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------- CPU leg

def cpu_reference_sample(sample: bytes, tdb_lists, attrs, summary, n_reads: int):
    """Time the CPU restatement (oracle/) on `sample`: C count, Python scoring.  Returns
    (Gbases/s, seconds, n_unique, n_rows)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kmer_oracle as ko_py
    import ko as ko_c
    t0 = time.perf_counter()
    counts, _ = ko_c.count_fastq(sample, PREFIX, K, STEP)
    rows = 0
    try:
        db = ko_py.TemplateDB(tdb_lists, attrs, summary)
        templates, _ = ko_py.first_match(counts, db)
        for _ in ko_py.find_matches(templates, summary, counts, len(counts)):
            rows += 1
    except RuntimeError:
        pass
    dt = time.perf_counter() - t0
    return n_reads * 150 / dt / 1e9, dt, len(counts), rows


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (restated in oracle/; the reference is
    JavaScript and there is no Node.js here or on the GPU box) on the host cores, one thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from kmerjs_b200 import synth
    n_reads = args.cpu_reads
    # the same generator as the GPU arm; it needs the GPU only to produce the bytes
    w = synth.Workload(n_reads=n_reads, genome_len=GENOME_LEN, seed=SEED)
    sample = w.host_bytes()
    tdb = synth.template_db_from_genome(w.genome_host(), N_TEMPLATES, PREFIX, K)
    lists, attrs = tdb.to_lists()
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(sample[: len(sample) // 8], lists, attrs, tdb.summary, n_reads // 8)
    vals, secs = [], []
    for _ in range(args.steps):
        v, dt, nuniq, nrows = cpu_reference_sample(sample, lists, attrs, tdb.summary, n_reads)
        vals.append(v)
        secs.append(dt)
    total = sum(secs)
    value = args.steps * n_reads * 150 / total / 1e9
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{n_reads} reads of the same workload per step (C count + Python "
                                       f"scoring of oracle/; the reference is single-threaded Node.js)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count()}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- clocks

class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([x.strip() for x in ln.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- GPU arm

def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from kmerjs_b200 import _abi, synth
    from kmerjs_b200.context import Context
    from kmerjs_b200.counts import Counts
    from kmerjs_b200.matching import Match, NoHitsError
    from kmerjs_b200 import dist as kdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch N > 1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: kmerjs_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local_rank, stream=stream.cuda_stream)

    n_reads = args.reads
    w = synth.Workload(n_reads=n_reads, genome_len=GENOME_LEN, seed=SEED, first_read=rank * n_reads, ctx=ctx)
    tdb = synth.template_db_from_genome(w.genome_host(), N_TEMPLATES, PREFIX, K)
    dbh = tdb.device(ctx, rank, world)          # DB resident in HBM before the timed region (the reference's Redis is up)
    hint = 1 << 20
    state = {}

    trace = {}

    def tick(name, t0):
        trace[name] = trace.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    def step_device():
        if world == 1:
            t = time.perf_counter()
            c = Counts(PREFIX, K, STEP, capacity_hint=hint, flags=state.get("flags", 0), ctx=ctx)
            t = tick("counts_create", t)
            c.add_device(w.fastq_ptr, w.n_bytes, final=True)
            t = tick("add_device", t)
            c.finish()
            t = tick("finish", t)
            m = Match(c, tdb)
            t = tick("first_match", t)
            rows, _end = m.all_rows()          # findMatches, the whole generator (kj_wta_all)
            t = tick("wta_rows", t)
            state.update(occ=c.occurrences, uniq=c.size, rows=rows, lines=c.lines, bases=c.bases)
            m.free(); c.free()
            t = tick("free", t)
        else:
            t = time.perf_counter()
            owned = kdist.count_sharded(w.fastq_ptr, w.n_bytes, w.n_bytes, prefix=PREFIX, k=K, step=STEP, final=True,
                                        base_line=rank * n_reads * 4, capacity_hint=hint, flags=state.get("flags", 0),
                                        ctx=ctx, trace=state.get("fine_trace"))
            t = tick("count+exchange", t)
            dm = kdist.DistMatch(owned, tdb, torch_stream=stream, mode=args.score_mode)
            t = tick("first_match+reduce", t)
            rows = []
            try:
                for r in dm.rows():      # the generator may end by throwing (query exhausted): keep what it yielded
                    rows.append(r)
            except NoHitsError:
                pass
            t = tick("wta_rows", t)
            state.update(occ=owned.occurrences, uniq=getattr(owned, "global_size", owned.size), rows=rows,
                         lines=owned.lines, bases=owned.bases)
            dm.free(); owned.free()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, n):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()        # from the warm-up to the end of the end-to-end timing: both timed regions are under it (20 ms period)
    # the first warm-up step also sums the sequence-line lengths on the device (KJ_F_COUNT_BASES) so that
    # the bases the throughput is quoted on are checked against what the kernel saw
    state["flags"] = _abi.KJ_F_COUNT_BASES
    step_device()
    assert state["bases"] == n_reads * 150 * world and state["lines"] == 4 * n_reads * world, state
    state["flags"] = 0
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    if args.trace:            # every rank runs the step (it contains collectives); rank 0 prints
        trace.clear()
        step_device()
        coarse = dict(trace)
        trace.clear()
        state["fine_trace"] = trace         # synchronising marks inside count_sharded: a second, slower step
        step_device()
        state["fine_trace"] = None
        fine = {k: round(v, 3) for k, v in trace.items() if k[:2] in ("c.", "x.")}
        trace.clear(); trace.update(coarse)
        if rank == 0 and fine:
            print("trace (ms, count+exchange phases, synchronised):", json.dumps(fine), file=sys.stderr)
        if rank == 0:
            print("trace (ms, one device-resident step):", json.dumps({k: round(v, 3) for k, v in trace.items()}), file=sys.stderr)
    ctx.enable_timers(True)
    ctx.reset_timers()
    l0 = ctx.launches
    ms_total = timed(step_device, args.steps)
    launches = ctx.launches - l0
    scan_ms, scan_n, scan_bytes = ctx.scan_kernel_stats()
    verify_ms = ctx.verify_kernel_ms()
    ctx.enable_timers(False)
    bases_per_step = n_reads * 150 * world
    assert state["lines"] == 4 * n_reads * world, state
    value = bases_per_step * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers --------------------------------------
    pinned = torch.empty(w.n_bytes, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(w.fastq._t[: w.n_bytes])
    torch.cuda.synchronize(dev)
    d2h = {"bytes": 0}
    dev_in = torch.empty(w.n_bytes + 64, dtype=torch.uint8, device=dev) if world > 1 else None

    def step_e2e():
        if world == 1:
            t = time.perf_counter()
            c = Counts(PREFIX, K, STEP, ctx=ctx)
            c.add_host(pinned, final=True)
            t = tick("e2e.add_host", t)
            c.finish()
            t = tick("e2e.finish", t)
            keys, lens, cnts = c.export_arrays()                     # the k-mer map, back on the host
            t = tick("e2e.export", t)
            m = Match(c, tdb)
            t = tick("e2e.first_match", t)
            rows, _end = m.all_rows()
            t = tick("e2e.wta_rows", t)
            d2h["bytes"] = keys.nbytes + lens.nbytes + cnts.nbytes + len(rows) * 136
            m.free(); c.free()
            tick("e2e.free", t)
        else:
            with torch.cuda.stream(stream):
                dev_in[: w.n_bytes].copy_(pinned, non_blocking=True)
            stream.synchronize()
            owned = kdist.count_sharded(dev_in.data_ptr(), w.n_bytes, w.n_bytes, prefix=PREFIX, k=K, step=STEP,
                                        final=True, base_line=rank * n_reads * 4, ctx=ctx)
            dm = kdist.DistMatch(owned, tdb, torch_stream=stream, mode=args.score_mode)
            rows = []
            try:
                for r in dm.rows():
                    rows.append(r)
            except NoHitsError:
                pass
            keys, lens, cnts = owned.export_arrays()
            d2h["bytes"] = keys.nbytes + lens.nbytes + cnts.nbytes + len(rows) * 136
            dm.free(); owned.free()

    step_e2e()
    if args.trace and rank == 0 and world == 1:
        trace.clear()
        step_e2e()
        print("trace (ms, one end-to-end step):", json.dumps({k: round(v, 3) for k, v in trace.items() if k.startswith("e2e.")}),
              file=sys.stderr)
    e2e_steps = max(1, args.e2e_steps)
    ms_e2e = timed(step_e2e, e2e_steps)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = bases_per_step * e2e_steps / (ms_e2e * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (scan: newline phase + extract + count) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    per_rank_occ = state["occ"] / world
    per_rank_uniq = state["uniq"] / world
    alg_bytes = scan_bytes / max(scan_n, 1) + 32.0 * per_rank_occ          # F + 32 * N_occ per launch
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    # DRAM bytes per launch from the committed `ncu --set full` capture of this workload (profiles/), scaled to
    # the bytes of this launch when the capture was taken at another size
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_input_byte"] * (scan_bytes / max(scan_n, 1))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": "kj_scan_filter_kernel<5,0> + kj_verify_kernel (extraction + count)",
                "kernel_ms": scan_ms, "scan_kernel_ms": scan_ms - verify_ms, "verify_kernel_ms": verify_ms,
                "launches_averaged": scan_n, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": scan_ms * scan_n / max(ms_total, 1e-9)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": w.n_bytes * world,
                    "d2h_bytes_per_step": d2h["bytes"], "steps": e2e_steps, "ms_per_step": ms_e2e / e2e_steps,
                    "api": "kj_counts_add_buffer(KJ_MEM_HOST, pinned) -> kj_counts_finish -> kj_counts_export -> "
                           "kj_first_match -> kj_wta_all" if world == 1 else
                           "pinned H2D -> dist.count_sharded -> DistMatch.rows -> export"},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "result": {"unique_kmers": int(state["uniq"]), "occurrences": int(state["occ"]),
                       "rows": len(state["rows"]), "winner": state["rows"][0]["template"] if state["rows"] else None},
            "fastq_bytes_per_gpu": w.n_bytes}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = min(args.cpu_reads, n_reads)
        sample = w.host_bytes(n_cpu)
        lists, attrs = tdb.to_lists()
        v, dt, nuniq, nrows = cpu_reference_sample(sample, lists, attrs, tdb.summary, n_cpu)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                                "host_cores": os.cpu_count(),
                                "sample": f"first {n_cpu} reads of the same workload (C count + Python scoring of "
                                          f"oracle/, one thread: the reference is a single Node.js event loop)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
