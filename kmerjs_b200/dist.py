"""Multi-GPU partitioning of the path: one process per GPU, torch.distributed for the
plumbing (NCCL over NVLink on the GPU box; gloo in the CPU tests of this host logic).

  reads      contiguous byte ranges of the FASTQ per rank (+ halo); the record phase of a range is
             fixed by the number of '\\n' before it (allgather of one integer per rank)
  k-mers     owner(key) = hash(key) mod world: locally counted (key, count, first-seen ordinal)
             records are exchanged with ONE all-to-all; owners merge with add / min
  templates  the DB is sharded by the same owner function; per-template partial vectors are
             all-reduced (sum for uScore/tScore/hits, min for the first-encounter keys)
  WTA        every rank holds the global score vector, takes the same argmax and gate, removes the
             winner's k-mers from its own shard, and the partial sums are all-reduced again.
"""
from __future__ import annotations

import numpy as np

from . import _abi
from .counts import Counts, count_newlines_device
from .matching import Match, NoHitsError

_SIGN = -(2 ** 63)


def plan_ranges(n_bytes: int, world: int, halo: int = 64, align: int = 16):
    """[(lo, own_n, read_n)] per rank: rank r owns window starts in [lo, lo+own_n) and may read
    read_n >= own_n bytes (halo for windows / lines that cross the cut).  Cuts are 16-byte aligned
    (device buffers are read with 16-byte vector loads)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    cuts = [min(n_bytes, (n_bytes * r // world) // align * align) for r in range(world)] + [n_bytes]
    out = []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        out.append((lo, hi - lo, min(n_bytes, hi + halo) - lo if hi < n_bytes else hi - lo))
    return out


def phase_of_ranges(newline_counts, last_newline_plus1, range_lens):
    """Given per-rank (number of '\\n', offset+1 of the last '\\n' or 0) of the OWNED ranges, return
    per-rank (base_line, base_col): lines before the range, bytes of the current line before it."""
    base_line, base_col = [], []
    lines, col = 0, 0
    for cnt, last, n in zip(newline_counts, last_newline_plus1, range_lens):
        base_line.append(lines)
        base_col.append(col)
        lines += cnt
        col = (n - last) if cnt else col + n
    return base_line, base_col


def _to_signed_order(t):
    """u64 bit patterns held in int64 -> int64 values with the same ORDER (for MIN reductions)."""
    return t ^ _SIGN


def allreduce_u64(t, op: str, group=None):
    """In-place all-reduce of a tensor of u64 bit patterns stored as int64.  'sum' wraps like u64;
    'min' compares as unsigned."""
    import torch.distributed as dist
    if op == "sum":
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    elif op == "min":
        t.bitwise_xor_(_SIGN)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        t.bitwise_xor_(_SIGN)
    else:
        raise ValueError(op)
    return t


def exchange_records(send, send_sizes, group=None):
    """All-to-all of owner-grouped records.  send: int64 tensor [n, 3] grouped by destination rank,
    send_sizes[r] = records for rank r.  Returns the int64 tensor [m, 3] this rank received."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = torch.tensor(send_sizes, dtype=torch.int64, device=send.device)
    recv_sizes = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv_sizes, sizes, group=group)
    rs = [int(x) for x in recv_sizes.tolist()]
    recv = torch.empty((sum(rs), 3), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send.reshape(-1, 3), output_split_sizes=rs,
                           input_split_sizes=[int(x) for x in send_sizes], group=group)
    return recv


class _CudaView:
    """Zero-copy torch view of device memory owned by libkmerjs_b200 (__cuda_array_interface__)."""

    def __init__(self, ptr: int, n_i64: int):
        self.__cuda_array_interface__ = {"shape": (n_i64,), "typestr": "<i8", "data": (ptr, False),
                                         "version": 2, "strides": None}


_IRR_INLINE = 1024 * 56        # bytes of irregular records that ride in the totals all-gather
_STAGE = {}


def _pinned_stage(device: int):
    import torch
    t = _STAGE.get(device)
    if t is None:
        t = torch.zeros(40 + _IRR_INLINE + 24, dtype=torch.uint8).pin_memory()
        _STAGE[device] = t
    return t


class _Trace:
    """Optional phase timing (a dict the caller passes): synchronises at every mark, so only for diagnosis."""

    def __init__(self, sink, dev):
        import time
        self.sink, self.dev, self.t = sink, dev, time.perf_counter()

    def mark(self, name):
        if self.sink is None:
            return
        import time
        import torch
        torch.cuda.synchronize(self.dev)
        now = time.perf_counter()
        self.sink[name] = self.sink.get(name, 0.0) + (now - self.t) * 1e3
        self.t = now


def exchange_counts(local: Counts, group=None, trace=None) -> Counts:
    """Local table -> the table of the k-mers this rank owns (collective)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device(f"cuda:{local.ctx.device}")
    tr = _Trace(trace, dev)
    ptr, sizes = local.partition(world)
    tr.mark("x.partition")
    n = sum(sizes)
    if n:
        send = torch.as_tensor(_CudaView(ptr, 3 * n), device=dev).reshape(-1, 3)
    else:
        send = torch.empty((0, 3), dtype=torch.int64, device=dev)
    recv = exchange_records(send, sizes, group)
    torch.cuda.synchronize(dev)
    tr.mark("x.all_to_all")
    owned = Counts(local.prefix, local.k, local.step, flags=local.flags & ~(_abi.KJ_F_FORWARD_ONLY),
                   capacity_hint=max(int(recv.shape[0]), 1024), ctx=local.ctx)
    if recv.shape[0]:
        owned.merge_records(recv.data_ptr(), int(recv.shape[0]))
    tr.mark("x.merge")
    # one small all-gather carries everything the ranks need from each other besides the records: line /
    # base / occurrence / byte totals and -- while they fit _IRR_INLINE bytes, the usual case: k-mers with
    # N or other non-ACGT bytes are rare -- the irregular records themselves (owned by hash like the rest)
    irr = local.irregular_records()
    head = np.array([irr.size, local.lines, local.bases, local.occurrences, local.bytes_read], dtype=np.int64)
    stage = _pinned_stage(local.ctx.device)
    stage[:40] = torch.from_numpy(head.view(np.uint8))
    n_inline = min(int(irr.size), _IRR_INLINE)
    if n_inline:
        stage[40:40 + n_inline] = torch.from_numpy(irr[:n_inline])
    mine_v = stage.to(dev, non_blocking=True)
    all_v = torch.empty((world, stage.numel()), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_v, mine_v, group=group)
    host_v = all_v.cpu().numpy()
    rows = [host_v[r, :40].view(np.int64).tolist() for r in range(world)]
    sizes_irr = [int(r[0]) for r in rows]
    tr.mark("x.totals_allgather")
    if max(sizes_irr) > _IRR_INLINE:
        pad = max(sizes_irr)
        mine = torch.zeros(pad, dtype=torch.uint8, device=dev)
        if irr.size:
            mine[:irr.size] = torch.from_numpy(irr.copy()).to(dev)
        parts = torch.empty((world, pad), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(parts, mine, group=group)
        host = parts.cpu().numpy()
        owned.merge_irregular(np.concatenate([host[rr, :sz] for rr, sz in enumerate(sizes_irr)]), rank, world)
    elif max(sizes_irr) > 0:
        # every rank holds all records after the all-gather and keeps the ones it owns
        owned.merge_irregular(np.concatenate([host_v[rr, 40:40 + sz] for rr, sz in enumerate(sizes_irr)]), rank, world)
    tr.mark("x.irregular")
    owned.finish()
    tr.mark("x.finish_owned")
    # a rank's line count already includes the lines before its range (base_line): the last rank's is the file's
    lines = max(int(r[1]) for r in rows)
    bases, occ, nbytes = (sum(int(r[i]) for r in rows) for i in (2, 3, 4))
    owned.set_totals(lines, bases, occ, nbytes)
    owned.global_size = None        # kmerMap.size over all ranks: global_size() / DistMatch fill it in
    tr.mark("x.set_totals")
    return owned


def global_size(owned: Counts, group=None) -> int:
    """kmerMap.size of the whole query = sum of the owners' table sizes (collective; cached on the handle).
    DistMatch folds this sum into its own first all-gather, so the scoring path never calls it."""
    import torch
    import torch.distributed as dist
    if getattr(owned, "global_size", None) is None:
        qs = torch.tensor([owned.size], dtype=torch.int64, device=torch.device(f"cuda:{owned.ctx.device}"))
        dist.all_reduce(qs, group=group)
        owned.global_size = int(qs.item())
    return owned.global_size


def count_sharded(dev_ptr: int, n_own: int, n_read: int, *, prefix=b"ATGAC", k=16, step=1, final: bool,
                  base_line: int | None = None, base_col: int = 0, capacity_hint: int = 0, flags: int = 0,
                  group=None, ctx=None, trace=None) -> Counts:
    """Count this rank's byte range and exchange.  With base_line None the ranks agree on the record
    phase first (newline counts of the owned ranges, allgathered)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if base_line is None:
        cnt, last = count_newlines_device(dev_ptr, n_own, ctx)
        dev = torch.device(f"cuda:{(ctx.device if ctx else 0)}")
        mine = torch.tensor([cnt, last, n_own], dtype=torch.int64, device=dev)
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine, group=group)
        rows = [[int(x) for x in v.tolist()] for v in allv]
        bl, bc = phase_of_ranges([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows])
        base_line, base_col = bl[rank], bc[rank]
    tr = _Trace(trace, torch.device(f"cuda:{(ctx.device if ctx else 0)}"))
    local = Counts(prefix, k, step, flags=flags, base_line=base_line, base_col=base_col,
                   capacity_hint=capacity_hint, ctx=ctx)
    tr.mark("c.create")
    local.add_device(dev_ptr, n_read, own_n=n_own, final=final)
    tr.mark("c.add_device")
    local.finish()
    tr.mark("c.finish_local")
    owned = exchange_counts(local, group, trace)
    tr.t = __import__("time").perf_counter()
    local.free()
    tr.mark("c.free_local")
    return owned


class DistMatch:
    """findFirstMatch + findMatches over ranks: every rank returns the same rows.

    mode "gather" (the default while the matched set is small enough): the ranks all-gather the query
    entries that hit their DB shard together with those entries' template lists -- one bandwidth-bound
    exchange over NVLink -- and each runs the winner-takes-all loop on the whole matched set, so no round
    of the loop waits on a collective.  mode "reduce": the matched set stays sharded and every round
    all-reduces the per-template sums ({u, tau, H}); memory per rank stays 1/N, each round pays a
    collective's latency.  "auto" picks "gather" up to gather_limit_pairs template-list entries."""

    def __init__(self, owned: Counts, db, group=None, torch_stream=None, mode: str = "auto",
                 gather_limit_pairs: int = 1 << 26):
        """torch_stream: the torch.cuda.Stream whose handle the Context was created on.  With it the
        collectives are stream-ordered (no host synchronisation between copy, collective and copy back)."""
        import torch
        import torch.distributed as dist
        if mode not in ("auto", "gather", "reduce"):
            raise ValueError(mode)
        self.group = group
        self.torch_stream = torch_stream
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device(f"cuda:{owned.ctx.device}")
        self.local = Match(owned, db, local_only=True, part=rank, n_parts=world)
        self.m = self.local
        self._buf = {}
        ne, npairs = self.local.matched_size()
        sz = torch.tensor([ne, npairs, owned.size], dtype=torch.int64, device=self.dev)
        all_sz = torch.empty((world, 3), dtype=torch.int64, device=self.dev)
        dist.all_gather_into_tensor(all_sz, sz, group=group)
        gathered = all_sz.tolist()
        sizes = [(int(a), int(b)) for a, b, _ in gathered]
        qsize = sum(int(c) for _, _, c in gathered)           # kmerMap.size (lib/kmerFinderClient.js:242)
        if getattr(owned, "global_size", None) is None:
            owned.global_size = qsize
        total_pairs = sum(b for _, b in sizes)
        self.mode = mode if mode != "auto" else ("gather" if total_pairs <= gather_limit_pairs else "reduce")
        if self.mode == "gather":
            self._gather(db, sizes, qsize, world, rank)
        else:
            self._reduce(_abi.KJ_VEC_SCORES, "sum")
            self._reduce(_abi.KJ_VEC_FIRST_ORD, "min")
            self._reduce(_abi.KJ_VEC_FIRST_IDX, "min")
        self.m.set_query_size(qsize)
        self.m.commit()
        self._gathered = None
        if self.m.hits == 0:
            raise NoHitsError("No hits were found!")

    def _gather(self, db, sizes, qsize, world, rank):
        import contextlib
        import torch
        import torch.distributed as dist
        max_e = max(max(a for a, _ in sizes), 1)
        max_p = max(max(b for _, b in sizes), 1)
        seg_bytes = max_e * 32 + ((max_p * 4 + 15) // 16) * 16          # {entries | template ids}, one payload
        mine = torch.empty(seg_bytes, dtype=torch.uint8, device=self.dev)
        allb = torch.empty((world, seg_bytes), dtype=torch.uint8, device=self.dev)
        ordered = torch.cuda.stream(self.torch_stream) if self.torch_stream is not None else contextlib.nullcontext()
        with ordered:
            self.local.export_matched(mine.data_ptr(), max_e, mine.data_ptr() + max_e * 32, max_p)
            if self.torch_stream is None:
                torch.cuda.synchronize(self.dev)
            dist.all_gather_into_tensor(allb, mine, group=self.group)
            if self.torch_stream is None:
                torch.cuda.synchronize(self.dev)
            segs = [(allb[r].data_ptr(), sizes[r][0], allb[r].data_ptr() + max_e * 32, sizes[r][1])
                    for r in range(world)]
            self.m = Match.from_matched(self.local.ctx, db, segs, qsize, part=rank, n_parts=world)
        self._gathered = (mine, allb)       # alive until commit() has synchronised

    def _reduce(self, which: int, op: str):
        import torch
        n = self.m.vec_len(which)
        t = self._buf.get(which)
        if t is None:
            t = torch.empty(max(n, 1), dtype=torch.int64, device=self.dev)
            self._buf[which] = t
        if self.torch_stream is not None:
            with torch.cuda.stream(self.torch_stream):      # everything below is ordered on the context's stream
                self.m.get(which, t.data_ptr(), sync=False)
                allreduce_u64(t[:n], op, self.group)
                self.m.set(which, t.data_ptr(), sync=False)
            return
        self.m.get(which, t.data_ptr())
        allreduce_u64(t[:n], op, self.group)
        torch.cuda.synchronize(self.dev)
        self.m.set(which, t.data_ptr())

    @property
    def hits(self):
        return self.m.hits

    def templates(self):
        return self.m.templates()

    def rows(self, max_hits: int = 100):
        self.m.set_max_hits(max_hits)
        if self.mode == "gather":
            yield from self.m.rows(max_hits)       # the whole loop in one call
            return
        self.m.defer_rows(True)
        while True:
            state, row = self.m.next_row_begin()
            if state == 0:
                return
            # the partial sums changed (the winner's k-mers are being removed): reduce them for the next
            # round; with a pending row the exact-decimal arithmetic runs on the host while the GPUs reduce
            self._reduce(_abi.KJ_VEC_SCORES, "sum")
            if state == 2:
                row = self.m.finish_row()
            yield row

    def free(self):
        if self.m is not self.local:
            self.m.free()
        self.local.free()
