"""The multi-GPU path over real NCCL (needs >= 2 GPUs; skipped otherwise): reads sharded over ranks,
owner all-to-all, sharded DB, all-reduced score vectors or the all-gathered matched set, replicated WTA -- every rank must produce the
rows of the single-GPU run over all the reads, in both scoring modes ("reduce", "gather"), each synchronous
and stream-ordered."""
import os
import socket

import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from kmerjs_b200 import dist as kdist, synth
    from kmerjs_b200.context import Context
    from kmerjs_b200.counts import Counts
    from kmerjs_b200.matching import Match, NoHitsError
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device(f"cuda:{rank}")
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        stream = torch.cuda.Stream(device=dev)
        ctx = Context(rank, stream=stream.cuda_stream)
        n_reads = 100000
        w = synth.Workload(n_reads=n_reads, genome_len=1_000_000, seed=5, first_read=rank * n_reads, ctx=ctx)
        tdb = synth.template_db_from_genome(w.genome_host(), 8, b"ATGAC", 16)
        out = {}
        for mode, ts, how in (("sync", None, "reduce"), ("stream", stream, "reduce"),
                              ("gather-sync", None, "gather"), ("gather-stream", stream, "auto")):
            owned = kdist.count_sharded(w.fastq_ptr, w.n_bytes, w.n_bytes, prefix=b"ATGAC", k=16, step=1, final=True,
                                        base_line=rank * n_reads * 4, ctx=ctx)
            dm = kdist.DistMatch(owned, tdb, torch_stream=ts, mode=how)
            assert dm.mode == ("gather" if how == "auto" else how)
            rows, err = [], None
            try:
                for r in dm.rows():
                    rows.append(r)
            except NoHitsError as exc:
                err = str(exc)
            out[mode] = dict(size=owned.global_size, lines=owned.lines, hits=dm.hits, order=list(dm.templates()),
                             rows=rows, err=err, counts=owned.to_dict())
            dm.free(); owned.free()
        # the lean path: after the two-phase exchange above has been seen, the same kind of job goes through the
        # fixed-capacity exchange (one all-to-all, one all-gather, no size round trips); then with capacities that are
        # too small: every rank must fall back together and still get the same answer
        for name, force in (("fixed", None), ("fixed-overflow", (256, 16, 256, 256))):
            key = kdist._caps_key(ctx, world, b"ATGAC", 16, 1)
            assert key in kdist._CAPS
            if force:
                kdist._CAPS[key] = force
            owned, dm = kdist.count_and_match(w.fastq_ptr, w.n_bytes, w.n_bytes, tdb, torch_stream=stream, mode="auto",
                                              prefix=b"ATGAC", k=16, step=1, final=True, base_line=rank * n_reads * 4,
                                              capacity_hint=1 << 18, ctx=ctx)
            took_fixed = getattr(owned, "_fixed_caps", None) is not None
            assert took_fixed == (force is None), (name, took_fixed)
            rows, err = [], None
            try:
                for r in dm.rows():
                    rows.append(r)
            except NoHitsError as exc:
                err = str(exc)
            out[name] = dict(size=owned.global_size, lines=owned.lines, hits=dm.hits, order=list(dm.templates()),
                             rows=rows, err=err, counts=owned.to_dict())
            dm.free()
            if getattr(owned, "_local", None) is not None:
                owned._local.free()
            owned.free()
        # count-only jobs (BASELINE config 5 shape: empty prefix, k = 31, reads with N -> byte-string k-mers travel too):
        # two-phase exchange first, then the fixed-capacity one, then one that is forced to overflow and must fall back
        n_dense = 4000
        wd = synth.Workload(n_reads=n_dense, genome_len=200_000, seed=9, first_read=rank * n_dense, ctx=ctx)
        dkey = kdist._caps_key(ctx, world, b"", 31, 1)
        for name in ("dense-two-phase", "dense-fixed", "dense-overflow"):
            if name == "dense-overflow":
                kdist._CAPS[dkey] = (1024, 16, 4096, 4096)
            owned = kdist.count_only(wd.fastq_ptr, wd.n_bytes, wd.n_bytes, torch_stream=stream, prefix=b"", k=31, step=1,
                                     final=True, base_line=rank * n_dense * 4, ctx=ctx)
            took_fixed = getattr(owned, "_fixed_caps", None) is not None
            assert took_fixed == (name == "dense-fixed"), (name, took_fixed)
            assert dkey in kdist._CAPS
            out[name] = dict(size=kdist.global_size(owned), lines=owned.lines, occ=owned.occurrences, counts=owned.to_dict())
            if getattr(owned, "_local", None) is not None:
                owned._local.free()
            owned.free()
        if rank == 0:
            wd2 = synth.Workload(n_reads=world * n_dense, genome_len=200_000, seed=9, first_read=0, ctx=ctx)
            cd = Counts(b"", 31, 1, ctx=ctx)
            cd.add_device(wd2.fastq_ptr, wd2.n_bytes, final=True).finish()
            out["dense-single"] = dict(size=cd.size, lines=cd.lines, occ=cd.occurrences, counts=cd.to_dict())
            cd.free()
        if rank == 0:      # single-GPU truth over the reads of all ranks
            w2 = synth.Workload(n_reads=world * n_reads, genome_len=1_000_000, seed=5, first_read=0, ctx=ctx)
            c = Counts(b"ATGAC", 16, 1, ctx=ctx)
            c.add_device(w2.fastq_ptr, w2.n_bytes, final=True).finish()
            counts = c.to_dict()
            m = Match(c, tdb)
            rows, err = [], None
            hits, order = m.hits, list(m.templates())
            try:
                while True:
                    r = m.next_row()
                    if r is None:
                        break
                    rows.append(r)
            except NoHitsError as exc:
                err = str(exc)
            out["single"] = dict(size=c.size, lines=c.lines, hits=hits, order=order, rows=rows, err=err, counts=counts)
            # and the CPU oracle over the same reads: map (keys, counts, Map order), first match, rows
            import sys
            sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
            import ko as ko_c
            import kmer_oracle as ko_py
            from collections import OrderedDict
            o_counts, o_lines = ko_c.count_fastq(w2.host_bytes(), b"ATGAC", 16, 1)
            lists, attrs = tdb.to_lists()
            o_first, o_hits, o_rows, o_err = ko_c.find_matches_fast(OrderedDict(o_counts),
                                                                     ko_py.TemplateDB(lists, attrs, tdb.summary))
            out["oracle"] = dict(counts=[(k.decode("latin-1"), v) for k, v in o_counts.items()], lines=o_lines,
                                 hits=o_hits, order=list(o_first), rows=[dict(r) for r in o_rows], err=o_err)
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_equal_single_gpu():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {}
    import queue as _queue
    import time as _time
    deadline = _time.monotonic() + 300
    try:
        while len(res) < world:                 # a rank that died must not cost the whole time limit
            try:
                k, v = q.get(timeout=2)
                res[k] = v
            except _queue.Empty:
                dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
                assert not dead, f"a rank exited with {dead}"
                assert _time.monotonic() < deadline, "ranks did not answer in 300 s"
        for p in procs:
            p.join(timeout=120)
            assert p.exitcode == 0
    finally:
        for p in procs:                         # a rank left waiting in a collective for a peer that is gone
            if p.is_alive():
                p.terminate()
    dsingle = res[0]["dense-single"]
    for name in ("dense-two-phase", "dense-fixed", "dense-overflow"):
        merged_d = {}
        for rank in range(world):
            o = res[rank][name]
            assert (o["size"], o["lines"], o["occ"]) == (dsingle["size"], dsingle["lines"], dsingle["occ"]), (name, rank)
            assert not (set(o["counts"]) & set(merged_d))                  # every k-mer has one owner
            merged_d.update(o["counts"])
        assert merged_d == dsingle["counts"], name
    single = res[0]["single"]
    assert len(single["rows"]) >= 1
    merged = {}
    for rank in range(world):
        for mode in ("sync", "stream", "gather-sync", "gather-stream", "fixed", "fixed-overflow"):
            o = res[rank][mode]
            for f in ("size", "lines", "hits", "order", "rows", "err"):
                assert o[f] == single[f], (rank, mode, f)
        assert not (set(res[rank]["sync"]["counts"]) & set(merged))       # every k-mer has one owner
        merged.update(res[rank]["sync"]["counts"])
        assert res[rank]["fixed"]["counts"] == res[rank]["sync"]["counts"]
        assert list(res[rank]["fixed"]["counts"]) == list(res[rank]["sync"]["counts"])      # Map order too
    assert merged == single["counts"]
    # the oracle: the merged map of the two ranks key for key, the single-GPU map in Map order, and the rows
    oracle = res[0]["oracle"]
    assert merged == dict(oracle["counts"])
    assert list(single["counts"].items()) == oracle["counts"] and single["lines"] == oracle["lines"]
    assert (single["hits"], single["order"], single["err"]) == (oracle["hits"], oracle["order"], oracle["err"])
    assert [r["template"] for r in single["rows"]] == [r["template"] for r in oracle["rows"]]
    for g, e in zip(single["rows"], oracle["rows"]):
        for f in e:
            if f == "probability":
                assert g[f] == pytest.approx(e[f], rel=1e-9)
            else:
                assert g[f] == e[f], (g["template"], f)
