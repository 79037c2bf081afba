import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kmerjs_b200 import _abi, synth
from kmerjs_b200.counts import Counts
from kmerjs_b200.context import default_context
ctx = default_context()
n_reads = 1000000
for label, kw in (("with N", {}), ("no N", dict(n_rate=0.0, lead_n_rate=0.0))):
    w = synth.Workload(n_reads=n_reads, genome_len=5_000_000, seed=11, sub_rate=0.001, **kw)
    for hint in (n_reads * 240, 40_000_000):
        for flags, name in ((0, "ordered"), (_abi.KJ_F_NO_ORDER, "no-order")):
            ctx.enable_timers(True); ctx.reset_timers()
            torch.cuda.synchronize(); t = time.perf_counter()
            c = Counts(b"", 31, 1, flags=flags, capacity_hint=hint)
            c.add_device(w.fastq_ptr, w.n_bytes, final=True)
            torch.cuda.synchronize(); t1 = time.perf_counter()
            c.finish()
            torch.cuda.synchronize(); dt = time.perf_counter() - t
            ms, n, b = ctx.scan_kernel_stats()
            print(f"{label} hint={hint} {name}: total {dt*1e3:.1f} ms (add {1e3*(t1-t):.1f}), kernels {ms*n:.1f} ms in {n} launches -> "
                  f"{c.occurrences/ (ms*n*1e-3)/1e9:.2f} G emissions/s in-kernel; unique {c.size}", flush=True)
            c.free()
