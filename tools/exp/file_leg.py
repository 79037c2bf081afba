#!/usr/bin/env python
"""DEVELOPER EXPERIMENT: the file leg (kmerjs(path, ...) on a tmpfs file) against the number of reader threads
(environment KJ_READ_THREADS, read by kj_counts_add_file at every call) and the staging chunk size."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

import kmerjs_b200  # noqa: E402
from kmerjs_b200 import synth  # noqa: E402
from kmerjs_b200.context import default_context  # noqa: E402

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ctx = default_context()
w = synth.Workload(n_reads=n_reads, genome_len=5_000_000, seed=1, first_read=0, ctx=ctx)
host = w.host_bytes()
path = f"/dev/shm/kmerjs_b200_fileleg_{os.getpid()}.fastq"
with open(path, "wb") as f:
    f.write(host)
try:
    kmerjs_b200.kmerjs(path, "ATGAC", 16, 1).result(timeout=600)
    for chunk_mb in (64, 256):
        ctx.set_stage_chunk(chunk_mb << 20)
        kmerjs_b200.kmerjs(path, "ATGAC", 16, 1).result(timeout=600)
        for nt in (4, 8, 12, 16, 24, 32):
            os.environ["KJ_READ_THREADS"] = str(nt)
            best = 1e9
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                m = kmerjs_b200.kmerjs(path, "ATGAC", 16, 1).result(timeout=600)
                best = min(best, time.perf_counter() - t0)
            print(f"chunk {chunk_mb} MiB, {nt:2d} reader threads: {best * 1e3:7.1f} ms  {len(host) / best / 1e9:6.2f} GB/s  "
                  f"{n_reads * 150 / best / 1e9:5.2f} Gbases/s  ({len(m)} keys, {os.cpu_count()} host cores)", flush=True)
finally:
    os.unlink(path)
