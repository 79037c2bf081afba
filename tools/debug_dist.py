"""2-rank debug of the distributed scoring path (run under torchrun on a 2-GPU box)."""
import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from kmerjs_b200 import _abi, synth, dist as kdist
from kmerjs_b200.context import Context
from kmerjs_b200.counts import Counts
from kmerjs_b200.matching import Match, NoHitsError

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device(f"cuda:{lr}")
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev)
ctx = Context(lr, stream=stream.cuda_stream)
n_reads = 200000
w = synth.Workload(n_reads=n_reads, genome_len=1_000_000, seed=5, first_read=rank * n_reads, ctx=ctx)
tdb = synth.template_db_from_genome(w.genome_host(), 8, b"ATGAC", 16)
owned = kdist.count_sharded(w.fastq_ptr, w.n_bytes, w.n_bytes, prefix=b"ATGAC", k=16, step=1, final=True,
                            base_line=rank * n_reads * 4, ctx=ctx)
print(rank, "owned", owned.size, "global", owned.global_size, "lines", owned.lines, flush=True)
for ts in (None, stream):
    try:
        dm = kdist.DistMatch(owned, tdb, torch_stream=ts)
        print(rank, "hits", dm.hits, "templates", list(dm.templates().items())[:3], flush=True)
        rows = list(dm.rows())
        print(rank, "rows", len(rows), rows[:1], flush=True)
        dm.free()
    except Exception as e:
        print(rank, "EXC", repr(e), flush=True); traceback.print_exc()
    # a second scoring needs the alive mask back: recount
    owned.free()
    owned = kdist.count_sharded(w.fastq_ptr, w.n_bytes, w.n_bytes, prefix=b"ATGAC", k=16, step=1, final=True,
                                base_line=rank * n_reads * 4, ctx=ctx)
# single-rank truth on rank 0: all reads of both ranks
if rank == 0:
    w2 = synth.Workload(n_reads=2 * n_reads, genome_len=1_000_000, seed=5, first_read=0, ctx=ctx)
    c = Counts(b"ATGAC", 16, 1, ctx=ctx); c.add_device(w2.fastq_ptr, w2.n_bytes, final=True).finish()
    m = Match(c, tdb); rows = []
    try:
        while True:
            r = m.next_row()
            if r is None: break
            rows.append(r)
    except NoHitsError as e:
        print("single: ", e)
    print("single-rank size", c.size, "hits", m.hits, "rows", len(rows), rows[:1], flush=True)
dist.barrier(); dist.destroy_process_group()
