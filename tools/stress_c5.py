"""BASELINE config 5 at reduced scale on one GPU: empty prefix, k = 31 (every window is an emission; the
hash table / atomics are the bound).  Prints throughput and checks totals + a sampled oracle comparison."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
from kmerjs_b200 import _abi, synth
from kmerjs_b200.counts import Counts

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
w = synth.Workload(n_reads=n_reads, genome_len=5_000_000, seed=11, sub_rate=0.001)
for flags, name in ((0, "ordered"), (_abi.KJ_F_NO_ORDER, "no-order")):
    torch.cuda.synchronize(); t = time.perf_counter()
    c = Counts(b"", 31, 1, flags=flags, capacity_hint=n_reads * 240)
    c.add_device(w.fastq_ptr, w.n_bytes, final=True).finish()
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"{name}: {n_reads} reads, {w.n_bytes/1e6:.0f} MB, {dt*1e3:.1f} ms -> {n_reads*150/dt/1e9:.3f} Gbases/s, "
          f"{c.occurrences/dt/1e9:.3f} G emissions/s; unique {c.size}, occurrences {c.occurrences} "
          f"(expected {n_reads*2*120})", flush=True)
    assert c.occurrences == n_reads * 2 * 120 and c.lines == 4 * n_reads
    if name == "ordered":
        import ko
        sample = w.host_bytes(2000)
        exp, lines = ko.count_fastq(sample, b"", 31, 1)
        cs = Counts(b"", 31, 1); cs.add_host(sample).finish()
        got = cs.to_dict()
        assert [(k.encode(), v) for k, v in got.items()] == list(exp.items()), "parity on the 2000-read sample"
        print("sample parity ok:", len(got), "k-mers")
        cs.free()
    c.free()
