#!/usr/bin/env python
"""bench.py -- Gbases/s of the kmerjs hot path (FASTQ -> k-mer counts + template scores) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config c3|c4|c5]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the whole path over one batch of synthetic reads: count (scan + extract + hash
count), first match against the template DB resident in HBM, winner-takes-all rows.

--config (BASELINE.json `configs`):
  c3 (default; the configuration the metric is quoted on at N = 1)  10 M x 150 bp Illumina-shaped reads per
      GPU, prefix ATGAC, k = 16, step = 1, sampled from a 5 Mbp genome; DB of 10 000 templates in genera
      that share k-mers (template 0 = the genome).  N > 1: every rank brings its own 10 M reads (weak).
  c4  100 M reads in total, split over the ranks (strong), against 10 000 templates / ~1e8 (k-mer, template)
      pairs, DB sharded by k-mer owner.
  c5  empty prefix, k = 31: every window is an emission (hash-table bound); 1/10 of BASELINE's size: 100 M reads
      in total over the ranks, count + owner exchange, no DB.

`value`      device-resident throughput (inputs already in HBM when the timed region starts)
`e2e`        the same path through the C ABI with HOST buffers: pinned FASTQ bytes in, H2D staging, k-mer map
             arrays and rows back out, all inside the timed region; `e2e.file` is the entry point a user calls,
             kmerjs(path, ...) on a tmpfs file (kj_counts_add_file: reader threads, two pinned buffers)
`roofline`   extraction + count kernels against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
`parity_checked`  outside the timed region: the GPU path on the first reads of the workload against the CPU
             oracle (map keys, counts, Map order, rows)
`cpu_baseline` / --impl reference: the CPU restatement of the reference algorithm (oracle/, one thread -- the
             reference is a single Node.js event loop) on a bounded sample.
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
SEED = 0x6B6D6572
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md, used only without MEASURED_PEAKS.json

CONFIGS = {
    "c3": dict(name="BASELINE configs[2]", reads_per_gpu=10_000_000, genome_len=5_000_000, prefix=b"ATGAC", k=16,
               step=1, templates=10_000, per_template=0, scaling="weak", sub_rate=0.005, cpu_reads=1_000_000),
    "c4": dict(name="BASELINE configs[3]", total_reads=100_000_000, genome_len=5_000_000, prefix=b"ATGAC", k=16,
               step=1, templates=10_000, per_template=10_000, scaling="strong", sub_rate=0.005, cpu_reads=1_000_000),
    "c5": dict(name="BASELINE configs[4] at 1/10 scale", total_reads=100_000_000, genome_len=50_000_000, prefix=b"",
               k=31, step=1, templates=0, per_template=0, scaling="strong", sub_rate=0.001, cpu_reads=50_000),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override: reads per GPU (c3) / in total (c4, c5)")
    ap.add_argument("--templates", type=int, default=-1, help="override the number of DB templates")
    ap.add_argument("--cpu-reads", type=int, default=0, help="reads in the CPU-baseline / parity sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file-leg", action="store_true")
    ap.add_argument("--score-mode", default="auto", choices=["auto", "gather", "reduce"],
                    help="N>1: all-gather the matched set once (gather) or all-reduce the score vector every round (reduce)")
    ap.add_argument("--trace", action="store_true", help="print host-side phase timings of one device step to stderr")
    return ap.parse_args()


def resolve_config(args, world):
    cfg = dict(CONFIGS[args.config])
    if cfg["scaling"] == "weak":
        cfg["reads_per_gpu"] = args.reads or cfg["reads_per_gpu"]
        cfg["total_reads"] = cfg["reads_per_gpu"] * world
    else:
        cfg["total_reads"] = args.reads or cfg["total_reads"]
        cfg["reads_per_gpu"] = (cfg["total_reads"] + world - 1) // world
    if args.templates >= 0:
        cfg["templates"] = args.templates
    cfg["cpu_reads"] = min(args.cpu_reads or cfg["cpu_reads"], cfg["reads_per_gpu"])
    return cfg


def workload_config(args, cfg, world):
    reads = cfg["reads_per_gpu"]
    db = (f"template DB of {cfg['templates']} synthetic templates in genera of 20 sharing k-mers (template 0 = the "
          f"{cfg['genome_len'] // 1_000_000} Mbp sample genome)" if cfg["templates"] else "no template DB (count + exchange only)")
    return {"workload": f"synthetic {reads / 1e6:g}M x 150bp Illumina-shaped reads per GPU ({cfg['total_reads'] / 1e6:g}M in total), "
                        f"prefix {cfg['prefix'].decode() or '(empty)'}, k={cfg['k']}, step={cfg['step']} ({cfg['name']}); {db}",
            "config": args.config, "reads_per_gpu": reads, "total_reads": cfg["total_reads"], "read_len": 150,
            "prefix": cfg["prefix"].decode(), "k": cfg["k"], "step": cfg["step"], "templates": cfg["templates"],
            "l2": f"inputs larger than L2 ({reads * 346 / 1e9:.2f} GB FASTQ per GPU vs 126 MB)",
            "sharding": ("whole records per rank; owner all-to-all of counted k-mers" + ("; scoring: " +
                         ("per-round all-reduce of the template sums" if args.score_mode == "reduce" else
                          "one all-gather of the matched entries, winner-takes-all replicated") if cfg["templates"] else ""))
            if world > 1 else "single GPU"}


# ---------------------------------------------------------------------------------------------- workload pieces

def sample_keys_of(genome, prefix: bytes, k: int):
    """2-bit keys of the prefix-filtered k-mers of the genome's forward strand (a template's k-mer set)."""
    import numpy as np
    g = np.frombuffer(genome, dtype=np.uint8) if not hasattr(genome, "dtype") else genome
    code = np.zeros(256, dtype=np.uint64)
    for ch, v in ((65, 0), (67, 1), (84, 2), (71, 3)):
        code[ch] = v
    c = code[g]
    n = c.size - k + 1
    key = np.zeros(n, dtype=np.uint64)
    for i in range(k):
        key = (key << np.uint64(2)) | c[i:i + n]
    pk = 0
    for b in prefix:
        pk = (pk << 2) | int(code[b])
    sel = (key >> np.uint64(2 * (k - len(prefix)))) == np.uint64(pk)
    return np.unique(key[sel])


def build_db(cfg, genome):
    from kmerjs_b200 import synth
    if not cfg["templates"]:
        return None
    keys = sample_keys_of(genome, cfg["prefix"], cfg["k"])
    return synth.genus_template_db(keys, cfg["templates"], cfg["per_template"] or int(keys.size), prefix=cfg["prefix"],
                                   k=cfg["k"], seed=11)


# ---------------------------------------------------------------------------------------------- CPU leg

def oracle_query_arrays(counts, tdb):
    """The oracle's k-mer map against the DB as arrays: (qcount, qoff, qt) for oracle/ko.wta_arrays."""
    import numpy as np
    k = int(tdb.kmer_len[0])
    keys = list(counts.keys())
    qcount = np.fromiter(counts.values(), dtype=np.uint64, count=len(keys))
    code = np.full(256, 255, dtype=np.uint8)
    for ch, v in ((65, 0), (67, 1), (84, 2), (71, 3)):
        code[ch] = v
    regular = np.array([len(x) == k for x in keys], dtype=bool)
    raw = np.frombuffer(b"".join(x if len(x) == k else b"N" * k for x in keys), dtype=np.uint8).reshape(-1, k)
    cs = code[raw]
    regular &= (cs != 255).all(axis=1)
    key = np.zeros(len(keys), dtype=np.uint64)
    for i in range(k):
        key = (key << np.uint64(2)) | (cs[:, i] & 3).astype(np.uint64)
    order = np.argsort(tdb.keys_u64)
    sk = tdb.keys_u64[order]
    at = np.searchsorted(sk, key)
    at = np.minimum(at, sk.size - 1)
    hit = regular & (sk[at] == key)
    pos = np.where(hit, order[at], 0)
    off = tdb.list_off.astype(np.int64)
    lens = np.where(hit, off[pos + 1] - off[pos], 0)
    qoff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    starts = np.repeat(off[pos], lens)
    within = np.arange(int(lens.sum())) - np.repeat(qoff[:-1].astype(np.int64), lens)
    qt = tdb.tmpl_ids[starts + within] if lens.sum() else np.zeros(1, dtype=np.uint32)
    return qcount, qoff, qt


def cpu_reference_sample(sample, cfg, tdb, n_reads: int):
    """The CPU restatement (oracle/) on `sample`: C count; scoring = the reference's full recount per round in C
    with the exact-decimal gate and rows in Python.  Returns (Gbases/s, seconds, counts, rows, error text)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ko as ko_c
    t0 = time.perf_counter()
    counts, lines = ko_c.count_fastq(sample, cfg["prefix"], cfg["k"], cfg["step"])
    rows, err = [], None
    if tdb is not None:
        qcount, qoff, qt = oracle_query_arrays(counts, tdb)
        attrs = cpu_reference_sample.attrs.get(id(tdb))
        if attrs is None:
            attrs = {n: {"lengths": int(tdb.lengths[i]), "ulength": int(tdb.ulengths[i]), "species": tdb.species[i]}
                     for i, n in enumerate(tdb.names)}
            cpu_reference_sample.attrs[id(tdb)] = attrs
        _first, _hits, rows, err = ko_c.wta_arrays(qcount, qoff, qt, tdb.names, attrs, tdb.summary, len(counts))
    dt = time.perf_counter() - t0
    return n_reads * 150 / dt / 1e9, dt, counts, rows, err, lines


cpu_reference_sample.attrs = {}


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm (restated in oracle/; the reference is JavaScript and
    there is no Node.js here or on the GPU box) on the host cores, one thread.  Nothing of the GPU library is
    loaded: the sample comes from the numpy restatement of the generator (oracle/synth_ref.py)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import synth_ref
    cfg = resolve_config(args, 1)
    n_reads = cfg["cpu_reads"]
    genome = synth_ref.genome(SEED, cfg["genome_len"])
    sample = synth_ref.fastq(SEED, n_reads, genome, sub_rate=cfg["sub_rate"])
    tdb = build_db(cfg, genome)
    for _ in range(min(args.warmup, 1)):
        cpu_reference_sample(sample[: (n_reads // 8) * 346], cfg, tdb, n_reads // 8)
    secs = []
    for _ in range(args.steps):
        _v, dt, counts, rows, _err, _lines = cpu_reference_sample(sample, cfg, tdb, n_reads)
        secs.append(dt)
    total = sum(secs)
    value = args.steps * n_reads * 150 / total / 1e9
    conf = workload_config(args, cfg, 1)
    conf["sample"] = f"every step counts and scores the first {n_reads} reads of the workload (a rate: comparable with the GPU arm's)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic", "impl": "reference",
            "config": conf,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{n_reads} reads of the same workload per step (C count + C recount loop + Python "
                                       f"exact-decimal rows of oracle/; the reference is single-threaded Node.js)"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "host_cores": os.cpu_count(),
            "result": {"unique_kmers": len(counts), "rows": len(rows)}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- NUMA

def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank on the cores of the NUMA node its GPU hangs off, so that the pinned staging buffers it allocates next
    (first touch) are local to the GPU's PCIe root: with eight ranks copying at once a remote buffer halves the H2D rate.
    Returns the node, or None when the platform does not say (virtualised PCI topology, one node)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:                                   # noqa: BLE001  (best effort: the bench runs without it)
        return None


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
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    stdout_fd = None
    if world > 1:
        # NCCL prints its version banner on the process's stdout when the first communicator comes up: the one JSON line is
        # the only thing this program may write there, so file descriptor 1 points at stderr until the line is printed
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = Context(local_rank, stream=stream.cuda_stream)

    cfg = resolve_config(args, world)
    PREFIX, K, STEP = cfg["prefix"], cfg["k"], cfg["step"]
    n_reads = cfg["reads_per_gpu"]
    first_read = rank * n_reads
    if cfg["scaling"] == "strong":
        n_reads = max(0, min(n_reads, cfg["total_reads"] - first_read))
    w = synth.Workload(n_reads=n_reads, genome_len=cfg["genome_len"], seed=SEED, first_read=first_read,
                       sub_rate=cfg["sub_rate"], ctx=ctx)
    genome = w.genome_host()
    tdb = build_db(cfg, genome)
    if tdb is not None:
        tdb.device(ctx, rank, world)            # DB resident in HBM before the timed region (the reference's Redis is up)
    scoring = tdb is not None
    # distinct k-mers one GPU meets: error variants dominate (c3: 175 k from 2.6 M occurrences)
    hint = (1 << 18) * max(1, n_reads // 10_000_000) if PREFIX else 0
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
            rows = []
            if scoring:
                m = Match(c, tdb)
                t = tick("first_match", t)
                rows, _end = m.all_rows()          # findMatches, the whole generator (kj_wta_all)
                t = tick("wta_rows", t)
                m.free()
            state.update(occ=c.occurrences, uniq=c.size, rows=rows, lines=c.lines, bases=c.bases)
            c.free()
            t = tick("free", t)
        else:
            t = time.perf_counter()
            kw = dict(prefix=PREFIX, k=K, step=STEP, final=True, base_line=first_read * 4, capacity_hint=hint,
                      flags=state.get("flags", 0), ctx=ctx, trace=state.get("fine_trace"))
            rows = []
            kdist._lean_mark(None, dev)
            if scoring:
                # count, owner exchange, matched gather; after the first job of a kind the exchanges have fixed capacities
                owned, dm = kdist.count_and_match(w.fastq_ptr, w.n_bytes, w.n_bytes, tdb, torch_stream=stream,
                                                  mode=args.score_mode, **kw)
                t = tick("count+exchange+first_match", t)
                try:
                    for r in dm.rows():      # the generator may end by throwing (query exhausted): keep what it yielded
                        rows.append(r)
                except NoHitsError:
                    pass
                t = tick("wta_rows", t)
                kdist._lean_mark("l.rows", dev)
                dm.free()
            else:
                # count + owner exchange; after the first job of a kind the exchange has fixed capacities
                owned = kdist.count_only(w.fastq_ptr, w.n_bytes, w.n_bytes, torch_stream=stream, **kw)
                t = tick("count+exchange", t)
            uniq = getattr(owned, "global_size", None)
            if uniq is None:
                uniq = kdist.global_size(owned)
            state.update(occ=owned.occurrences, uniq=uniq, rows=rows, lines=owned.lines, bases=owned.bases)
            if getattr(owned, "_local", None) is not None:
                owned._local.free()
            owned.free()

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
    total_reads = cfg["total_reads"]
    assert state["bases"] == total_reads * 150 and state["lines"] == 4 * total_reads, state
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
        if rank == 0 and kdist._LEAN["sink"]:
            print("trace (ms, fixed-capacity path, every phase followed by a device wait):", json.dumps(kdist._LEAN["sink"]), file=sys.stderr)
    ctx.enable_timers(True)
    ctx.reset_timers()
    l0 = ctx.launches
    ms_total = timed(step_device, args.steps)
    launches = ctx.launches - l0
    scan_ms, scan_n, scan_bytes = ctx.scan_kernel_stats()
    verify_ms = ctx.verify_kernel_ms()
    ctx.enable_timers(False)
    bases_per_step = total_reads * 150
    assert state["lines"] == 4 * total_reads, state
    value = bases_per_step * args.steps / (ms_total * 1e-3) / 1e9

    # ---- end to end through the C ABI with host buffers --------------------------------------
    pinned = torch.empty(max(w.n_bytes, 16), dtype=torch.uint8, pin_memory=True)
    pinned[: w.n_bytes].copy_(w.fastq._t[: w.n_bytes])
    torch.cuda.synchronize(dev)
    d2h = {"bytes": 0}
    dev_in = torch.empty(w.n_bytes + 64, dtype=torch.uint8, device=dev) if world > 1 else None

    def step_e2e():
        if world == 1:
            t = time.perf_counter()
            c = Counts(PREFIX, K, STEP, ctx=ctx)
            c.add_host(pinned[: w.n_bytes], final=True)
            t = tick("e2e.add_host", t)
            c.finish()
            t = tick("e2e.finish", t)
            keys, lens, cnts = c.export_arrays()                     # the k-mer map, back on the host
            t = tick("e2e.export", t)
            rows = []
            if scoring:
                m = Match(c, tdb)
                t = tick("e2e.first_match", t)
                rows, _end = m.all_rows()
                t = tick("e2e.wta_rows", t)
                m.free()
            d2h["bytes"] = keys.nbytes + lens.nbytes + cnts.nbytes + len(rows) * 136
            c.free()
            tick("e2e.free", t)
        else:
            with torch.cuda.stream(stream):
                dev_in[: w.n_bytes].copy_(pinned[: w.n_bytes], non_blocking=True)
            stream.synchronize()
            kw = dict(prefix=PREFIX, k=K, step=STEP, final=True, base_line=first_read * 4, capacity_hint=hint, ctx=ctx)
            rows = []
            if scoring:
                owned, dm = kdist.count_and_match(dev_in.data_ptr(), w.n_bytes, w.n_bytes, tdb, torch_stream=stream,
                                                  mode=args.score_mode, **kw)
                try:
                    for r in dm.rows():
                        rows.append(r)
                except NoHitsError:
                    pass
                dm.free()
            else:
                owned = kdist.count_sharded(dev_in.data_ptr(), w.n_bytes, w.n_bytes, fixed=False, **kw)
            keys, lens, cnts = owned.export_arrays()
            d2h["bytes"] = keys.nbytes + lens.nbytes + cnts.nbytes + len(rows) * 136
            if getattr(owned, "_local", None) is not None:
                owned._local.free()
            owned.free()

    e2e = None
    if args.config != "c5":         # c5's map (10^8 .. 10^9 keys) is not exported key by key
        step_e2e()
        if args.trace and rank == 0 and world == 1:
            trace.clear()
            step_e2e()
            print("trace (ms, one end-to-end step):", json.dumps({k: round(v, 3) for k, v in trace.items() if k.startswith("e2e.")}),
                  file=sys.stderr)
        e2e_steps = max(1, args.e2e_steps)
        ms_e2e = timed(step_e2e, e2e_steps)
        e2e = {"value": bases_per_step * e2e_steps / (ms_e2e * 1e-3) / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": w.n_bytes * world, "d2h_bytes_per_step": d2h["bytes"], "steps": e2e_steps,
               "ms_per_step": ms_e2e / e2e_steps,
               "api": "kj_counts_add_buffer(KJ_MEM_HOST, pinned) -> kj_counts_finish -> kj_counts_export -> "
                      "kj_first_match -> kj_wta_all" if world == 1 else
                      "pinned H2D -> dist.count_sharded -> DistMatch.rows -> export"}
        # the entry point a user calls: kmerjs(path, prefix, k, step) on a file (tmpfs, so the disk is not what is timed)
        if world == 1 and not args.no_file_leg and os.path.isdir("/dev/shm"):
            import kmerjs_b200
            path = f"/dev/shm/kmerjs_b200_bench_{os.getpid()}.fastq"
            try:
                free = os.statvfs("/dev/shm")
                if free.f_bavail * free.f_frsize > w.n_bytes + (1 << 28):
                    with open(path, "wb") as f:
                        f.write(memoryview(pinned[: w.n_bytes].numpy()))
                    kmerjs_b200.kmerjs(path, PREFIX.decode(), K, STEP).result(timeout=600)      # warm-up (page cache, pinned buffers)
                    torch.cuda.synchronize(dev)
                    t0 = time.perf_counter()
                    m = kmerjs_b200.kmerjs(path, PREFIX.decode(), K, STEP).result(timeout=600)
                    dt = time.perf_counter() - t0
                    e2e["file"] = {"value": bases_per_step / dt / 1e9, "unit": UNIT, "ms_per_step": dt * 1e3,
                                   "keys": len(m), "api": "kmerjs(path, prefix, k, step) -> kj_counts_add_file (two pinned "
                                                          "buffers, reader threads) -> finish -> export -> Map", "timer": "host wall clock"}
            finally:
                if os.path.exists(path):
                    os.unlink(path)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernels (scan + resolve: extraction + count) ----------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    else:
        peak, peak_src = FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"
    per_rank_occ = state["occ"] / world
    # F + 32 * N_occ per launch: a step may take several launches (config 5 counts in 8 MiB pieces), the
    # occurrences of a step are spread over them
    launches_per_step = max(scan_n, 1) / max(args.steps, 1)
    alg_bytes = scan_bytes / max(scan_n, 1) + 32.0 * per_rank_occ / max(launches_per_step, 1.0)
    achieved = alg_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms > 0 else 0.0
    # DRAM bytes per launch from the committed `ncu --set full` capture of this workload (profiles/), scaled to
    # the bytes of this launch when the capture was taken at another size
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.config != "c5":
        tj = json.load(open(tpath))
        traffic = tj["dram_bytes_per_input_byte"] * (scan_bytes / max(scan_n, 1))
    kname = ("kj_scan_dense_kernel (extraction + count)" if not PREFIX else
             "kj_warp_filter_kernel<5> + cub tile scan + kj_resolve_kernel<4> (extraction + count)")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": kname,
                "kernel_ms": scan_ms, "scan_kernel_ms": scan_ms - verify_ms, "resolve_kernel_ms": verify_ms,
                "launches_averaged": scan_n, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "kernel_share_of_step": scan_ms * scan_n / max(ms_total, 1e-9)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": dict(workload_config(args, cfg, world), numa_node_of_rank0=numa_node),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "result": {"unique_kmers": int(state["uniq"]), "occurrences": int(state["occ"]),
                       "rows": len(state["rows"]), "winner": state["rows"][0]["template"] if state["rows"] else None},
            "fastq_bytes_per_gpu": w.n_bytes}

    # ---- parity at benchmark inputs + CPU baseline beside it (rank 0, N = 1 only) ---------------
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = cfg["cpu_reads"]
        sample = w.host_bytes(n_cpu)
        v, dt, o_counts, o_rows, o_err, o_lines = cpu_reference_sample(sample, cfg, tdb, n_cpu)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "seconds": dt,
                                "host_cores": os.cpu_count(),
                                "sample": f"first {n_cpu} reads of the same workload (C count + C recount loop + Python "
                                          f"exact-decimal rows of oracle/, one thread: the reference is a single Node.js event loop)"}
        # the same reads through the GPU path, outside every timed region: map (keys, counts, Map order), lines, rows
        c = Counts(PREFIX, K, STEP, capacity_hint=hint, ctx=ctx)
        c.add_device(w.fastq_ptr, n_cpu * w.record_bytes, final=True).finish()
        keys, lens, cnts = c.export_arrays()
        raw = keys.tobytes()
        g_items = [(raw[32 * i:32 * i + int(lens[i])], int(cnts[i])) for i in range(len(lens))]
        assert c.lines == o_lines, (c.lines, o_lines)
        assert g_items == list(o_counts.items()), "k-mer map differs from the oracle's (keys, counts or Map order)"
        g_rows, g_err = [], None
        if scoring:
            m = Match(c, tdb)
            g_rows, g_err = m.all_rows()
            m.free()
            assert [r["template"] for r in g_rows] == [r["template"] for r in o_rows], "winner order differs from the oracle's"
            for g, e in zip(g_rows, o_rows):
                for f in e:
                    if f == "probability":
                        assert abs(g[f] - e[f]) <= 1e-9 * abs(e[f]), (g["template"], f)
                    else:
                        assert g[f] == e[f], (g["template"], f, g[f], e[f])
            assert (str(g_err) if g_err else None) == o_err
        c.free()
        line["parity_checked"] = {"reads": n_cpu, "keys": len(g_items), "occurrences": int(sum(v for _, v in g_items)),
                                  "rows": len(g_rows), "against": "oracle/ (bit-exact map and Map order, rows exact, probability 1e-9)"}
    if stdout_fd is not None:
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        os.close(stdout_fd)
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
