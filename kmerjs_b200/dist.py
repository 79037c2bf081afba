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
    x_seen = [max(sizes) if sizes else 0, 0]        # what the fixed-capacity exchange would have needed per peer
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
    # learn the capacities of the fixed-capacity exchange from what all ranks saw (identical on every rank)
    # (irregular records are owned by hash like the rest: a peer gets about 1/world of a rank's; generous slack)
    x_seen[1] = (max(sizes_irr) // 56 + world - 1) // world * 2 + 2048
    mx = torch.tensor(x_seen, dtype=torch.int64, device=dev)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    owned._seen = [int(v) for v in mx.tolist()]
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


# Capacities of the fixed-capacity exchange, learned from the sizes the two-phase exchange saw for the same kind of
# job: {(device, world, prefix, k, step): (records per peer, irregular records per peer, matched entries, matched pairs)}.
# Every rank derives them from the same collective results, so all ranks take the same path.
_CAPS = {}
_XBUF = {}


def _caps_key(ctx, world, prefix, k, step):
    return (ctx.device if ctx else 0, world, bytes(prefix), int(k), int(step))


def _round_cap(n: int, floor: int) -> int:
    """1.5 x what was seen, rounded up to a multiple of 1024."""
    return max(floor, (int(n * 1.5) + 1023) // 1024 * 1024)


def _xbuf(dev, name: str, nbytes: int):
    """Persistent device buffers of the exchange (allocated once per size)."""
    import torch
    key = (dev.index, name)
    t = _XBUF.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _XBUF[key] = t
    return t[:nbytes]


_LEAN = {"on": bool(__import__("os").environ.get("KJ_TRACE_LEAN")), "t": None, "sink": {}}


def _lean_mark(name: str, dev=None):
    """Developer trace of the fixed-capacity path (environment KJ_TRACE_LEAN=1): every mark waits for the device, so the
    numbers are phase costs, not what the overlapped step pays.  name None: start."""
    if not _LEAN["on"]:
        return
    import time
    import torch
    torch.cuda.synchronize(dev)
    now = time.perf_counter()
    if name is not None and _LEAN["t"] is not None:
        _LEAN["sink"][name] = round((now - _LEAN["t"]) * 1e3, 3)
    _LEAN["t"] = now


def exchange_counts_fixed(local: Counts, caps, group=None, torch_stream=None) -> Counts:
    """Local table -> the table of the k-mers this rank owns, with ONE equal-split all-to-all and no host wait before
    it: the sender scatters its table into per-owner segments of fixed capacity (record counts and its totals in the
    segment headers), the owner merges what it receives, counts read on the device.  Overflow surfaces as KjError in
    the returned handle's finish(), which the caller has not called yet."""
    import contextlib
    import torch
    import torch.distributed as dist
    from .counts import segment_bytes
    world = dist.get_world_size(group)
    dev = torch.device(f"cuda:{local.ctx.device}")
    cap_reg, cap_irr = caps[0], caps[1]
    seg = segment_bytes(cap_reg, cap_irr)
    send = _xbuf(dev, "send", world * seg)
    recv = _xbuf(dev, "recv", world * seg)
    ordered = torch.cuda.stream(torch_stream) if torch_stream is not None else contextlib.nullcontext()
    with ordered:
        _lean_mark("l.count", dev)
        local.partition_segments(world, send.data_ptr(), cap_reg, cap_irr)
        if torch_stream is None:
            torch.cuda.synchronize(dev)
        _lean_mark("l.partition", dev)
        dist.all_to_all_single(recv, send, group=group)
        if torch_stream is None:
            torch.cuda.synchronize(dev)
        _lean_mark("l.all_to_all", dev)
        owned = Counts(local.prefix, local.k, local.step, flags=local.flags & ~(_abi.KJ_F_FORWARD_ONLY),
                       capacity_hint=world * cap_reg, ctx=local.ctx)
        owned.merge_segments(recv.data_ptr(), world, cap_reg, cap_irr)
        _lean_mark("l.merge", dev)
    owned.global_size = None
    return owned


def count_sharded(dev_ptr: int, n_own: int, n_read: int, *, prefix=b"ATGAC", k=16, step=1, final: bool,
                  base_line: int | None = None, base_col: int = 0, capacity_hint: int = 0, flags: int = 0,
                  group=None, ctx=None, trace=None, torch_stream=None, fixed: bool = True) -> Counts:
    """Count this rank's byte range and exchange.  With base_line None the ranks agree on the record
    phase first (newline counts of the owned ranges, allgathered).

    With a capacity hint, once a job of the same kind has gone through the two-phase exchange on these ranks, the
    fixed-capacity exchange is used (`fixed`): nothing waits on the host between counting and the owner's finish().
    If it overflows, finish() raises KjError(KJ_E_RANGE) on every affected rank and DistMatch / the caller falls back
    (redo with fixed=False)."""
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
    caps = _CAPS.get(_caps_key(local.ctx, world, prefix, k, step))
    # (fixed == "always": count_only, which agrees on the outcome itself, also takes it without a capacity hint)
    if fixed and (capacity_hint or fixed == "always") and caps is not None and trace is None:
        owned = exchange_counts_fixed(local, caps, group, torch_stream)
        owned._local = local            # its table backs nothing any more, but its buffers must outlive the queued kernels
        owned._fixed_caps = caps
        return owned
    local.finish()
    tr.mark("c.finish_local")
    owned = exchange_counts(local, group, trace)
    tr.t = __import__("time").perf_counter()
    local.free()
    tr.mark("c.free_local")
    return owned


def count_only(dev_ptr: int, n_own: int, n_read: int, *, group=None, torch_stream=None, **kw) -> Counts:
    """count_sharded for jobs without scoring (collective), finished: the first job of a kind goes through the two-phase
    exchange and leaves the capacities for the fixed-capacity exchange of the next ones; whether that one fitted is agreed by
    all ranks (one small all-reduce), and if it did not the job is redone the robust way."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if kw.get("fixed", True):
        kw["fixed"] = "always"
    owned = count_sharded(dev_ptr, n_own, n_read, group=group, torch_stream=torch_stream, **kw)
    key = _caps_key(owned.ctx, world, owned.prefix, owned.k, owned.step)
    if getattr(owned, "_fixed_caps", None) is None:
        seen = getattr(owned, "_seen", None)
        if seen is not None:
            _CAPS[key] = (_round_cap(seen[0], 4096), _round_cap(seen[1], 1024), 4096, 4096)
        return owned
    bad = 0
    try:
        owned.finish()
    except _abi.KjError as exc:
        if exc.code != _abi.KJ_E_RANGE:
            raise
        bad = 1
    flag = torch.tensor([bad], dtype=torch.int64, device=torch.device(f"cuda:{owned.ctx.device}"))
    dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
    if int(flag.item()) == 0:
        return owned
    _CAPS.pop(key, None)
    getattr(owned, "_local", owned).free()
    owned.free()
    kw["fixed"] = False
    return count_only(dev_ptr, n_own, n_read, group=group, torch_stream=torch_stream, **kw)


class ExchangeRetry(RuntimeError):
    """The fixed-capacity exchange did not fit (on some rank): every rank raises this at the same point; redo the job
    with count_sharded(..., fixed=False).  count_and_match() does that by itself."""


def count_and_match(dev_ptr: int, n_own: int, n_read: int, db, *, torch_stream=None, mode: str = "auto", group=None, **kw):
    """count_sharded + DistMatch with the fall-back from the fixed-capacity exchange handled: (owned, dist_match)."""
    if mode == "reduce":
        kw["fixed"] = False                     # the per-round all-reduce keeps the matched set sharded: two-phase exchange
    owned = count_sharded(dev_ptr, n_own, n_read, torch_stream=torch_stream, group=group, **kw)
    try:
        return owned, DistMatch(owned, db, group=group, torch_stream=torch_stream, mode=mode)
    except ExchangeRetry:
        _CAPS.pop(_caps_key(owned.ctx, __import__("torch").distributed.get_world_size(group), owned.prefix, owned.k, owned.step), None)
        getattr(owned, "_local", owned).free()
        owned.free()
        kw["fixed"] = False
        owned = count_sharded(dev_ptr, n_own, n_read, torch_stream=torch_stream, group=group, **kw)
        return owned, DistMatch(owned, db, group=group, torch_stream=torch_stream, mode=mode)


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
        self.m = self.local = None
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        self.dev = torch.device(f"cuda:{owned.ctx.device}")
        self._buf = {}
        self._gathered = None
        caps = getattr(owned, "_fixed_caps", None)
        if caps is not None:
            if mode == "reduce":
                raise ValueError("a handle from the fixed-capacity exchange goes with the gathered match (mode 'auto' or 'gather')")
            self._init_fixed(owned, db, caps, world, rank)
            return
        self.local = Match(owned, db, local_only=True, part=rank, n_parts=world)
        self.m = self.local
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
        seen = getattr(owned, "_seen", None)
        if seen is not None and self.mode == "gather":
            # what the fixed-capacity exchange of the next job of this kind will be sized for
            _CAPS[_caps_key(owned.ctx, world, owned.prefix, owned.k, owned.step)] = (
                _round_cap(seen[0], 4096), _round_cap(seen[1], 1024),
                _round_cap(max(a for a, _ in sizes), 4096), _round_cap(max(b for _, b in sizes), 4096))
        if self.m.hits == 0:
            raise NoHitsError("No hits were found!")

    def _init_fixed(self, owned, db, caps, world, rank):
        """The lean path: owner tables merged from fixed-capacity segments (not finished yet), matched entries
        gathered in fixed-capacity segments.  Three host waits in all: the owner's finish(), the commit, the rows."""
        import contextlib
        import torch
        import torch.distributed as dist
        from .matching import matched_segment_bytes
        cap_e, cap_p = caps[2], caps[3]
        seg = matched_segment_bytes(cap_e, cap_p)
        mine = _xbuf(self.dev, "mseg", seg)
        allb = _xbuf(self.dev, "mall", world * seg)
        ordered = torch.cuda.stream(self.torch_stream) if self.torch_stream is not None else contextlib.nullcontext()
        self.local = None
        # the short way: this rank's matched entries straight from the owner's hash table -- its finish() (compaction,
        # counters back, a host wait) is queued behind the gather and no longer sits between the two collectives
        with ordered:
            fast = owned.export_matched_segment(db.device(owned.ctx, rank, world), mine.data_ptr(), cap_e, cap_p)
        ok = True
        if not fast:
            try:
                owned.finish()
            except _abi.KjError as exc:
                if exc.code != _abi.KJ_E_RANGE:
                    raise
                ok = False                                  # this rank's exchange overflowed: tell everybody
            _lean_mark("l.finish_owned", self.dev)
        with ordered:
            if fast:
                _lean_mark("l.export_from_table", self.dev)
            elif ok:
                self.local = Match(owned, db, local_only=True, part=rank, n_parts=world)
                _lean_mark("l.first_match_local", self.dev)
                self.local.export_segment(mine.data_ptr(), cap_e, cap_p, owned.size, 0)
                _lean_mark("l.export_segment", self.dev)
            else:
                mine[:32].copy_(torch.tensor([0, 0, 0, 1], dtype=torch.int64).view(torch.uint8), non_blocking=False)
            if self.torch_stream is None:
                torch.cuda.synchronize(self.dev)
            dist.all_gather_into_tensor(allb.view(world, seg), mine, group=self.group)
            if self.torch_stream is None:
                torch.cuda.synchronize(self.dev)
            _lean_mark("l.all_gather", self.dev)
            self.m = Match.from_segments(owned.ctx, db, world, allb.data_ptr(), cap_e, cap_p, part=rank, n_parts=world)
            _lean_mark("l.from_segments", self.dev)
        finish_failed = False
        if fast:
            try:
                owned.finish()                          # its wait also covers the gather and the import
            except _abi.KjError as exc:
                if exc.code != _abi.KJ_E_RANGE:
                    raise
                finish_failed = True                    # the header of this rank's segment says so too: commit fails everywhere
            _lean_mark("l.finish_owned", self.dev)
        self._gathered = (mine, allb)                   # the template lists are used in place
        self.mode = "gather"
        try:
            self.m.commit()
        except _abi.KjError as exc:
            if exc.code != _abi.KJ_E_RANGE:
                raise
            self.free()
            raise ExchangeRetry(str(exc)) from exc
        _lean_mark("l.commit", self.dev)
        if finish_failed:
            self.free()
            raise ExchangeRetry("the owner's exchange overflowed")
        owned.global_size = self.m.query_size
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
        if getattr(self, "m", None) is not None and self.m is not self.local:
            self.m.free()
        if getattr(self, "local", None) is not None:
            self.local.free()
        self.m = self.local = None
