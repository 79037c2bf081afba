"""H2D staging chunk sweep for the end-to-end path (host pinned FASTQ -> counts)."""
import os, sys, time, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = '''
import sys, time, torch
sys.path.insert(0, %r)
from kmerjs_b200 import synth
from kmerjs_b200.counts import Counts
w = synth.Workload(n_reads=10_000_000, genome_len=5_000_000, seed=0x6B6D6572)
pinned = torch.empty(w.n_bytes, dtype=torch.uint8, pin_memory=True); pinned.copy_(w.fastq._t[:w.n_bytes]); torch.cuda.synchronize()
dev = torch.empty(w.n_bytes, dtype=torch.uint8, device="cuda")
for _ in range(2):
    t = time.perf_counter(); dev.copy_(pinned, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("plain cudaMemcpy H2D GB/s", w.n_bytes / dt / 1e9)
for _ in range(3):
    t = time.perf_counter(); c = Counts(b"ATGAC", 16, 1); c.add_host(pinned, final=True).finish(); dt = time.perf_counter() - t; n = c.size; c.free()
print("chunk", %r, "MB: count from pinned host", dt * 1e3, "ms", w.n_bytes / dt / 1e9, "GB/s", n)
'''
for mb in (16, 64, 256, 1024):
    env = dict(os.environ, KJ_STAGE_CHUNK_MB=str(mb))
    print(subprocess.run([sys.executable, "-c", code % (root, mb)], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
