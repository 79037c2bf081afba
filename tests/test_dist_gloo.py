"""Host-side multi-rank logic over gloo (world_size 2, CPU): the owner exchange and the unsigned
reductions that the GPU ranks run over NCCL.  The records are produced by the oracle, the owner
function is the library's (kj_owner, host-side)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ko as ko_c
from conftest import read_golden

from kmerjs_b200 import _abi
from kmerjs_b200 import dist as kdist

CODE = {65: 0, 67: 1, 84: 2, 71: 3}


def pack(key: bytes) -> int:
    v = 0
    for b in key:
        v = (v << 2) | CODE[b]
    return v


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        data = read_golden("test_long.kmer.fastq")
        ranges = kdist.plan_ranges(len(data), world, halo=64)
        lo, own, rd = ranges[rank]
        # record phase by allgather (what count_sharded does with the device newline kernel)
        mine = torch.tensor([data[lo:lo + own].count(b"\n"), data[lo:lo + own].rfind(b"\n") + 1, own])
        allv = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        rows = [v.tolist() for v in allv]
        bl, bc = kdist.phase_of_ranges([r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows])
        assert bl[rank] == data[:lo].count(b"\n")
        # local count of this rank's lines (oracle stands in for the GPU kernel in this CPU test):
        # whole lines whose start lies in the owned range
        lines = data.split(b"\n")
        pos, local = 0, {}
        for i, ln in enumerate(lines[:-1]):
            if lo <= pos < lo + own and i % 4 == 1:
                c, _ = ko_c.count_fastq(b"@\n" + ln + b"\n")
                for kk, v in c.items():
                    local[kk] = local.get(kk, 0) + v
            pos += len(ln) + 1
        L = _abi.lib()
        recs = [[] for _ in range(world)]
        for kk, v in local.items():
            if all(b in CODE for b in kk):
                recs[L.kj_owner(kk, len(kk), world)].append((pack(kk), v, rank * 10**6 + len(recs[0])))
        send = torch.tensor([r for part in recs for r in part] or np.zeros((0, 3)), dtype=torch.int64).reshape(-1, 3)
        recv = kdist.exchange_records(send, [len(p) for p in recs])
        owned = {}
        for key, cnt, _ in recv.tolist():
            assert L.kj_owner(_unpack(key), 16, world) == rank
            owned[key] = owned.get(key, 0) + cnt
        # unsigned reductions: min must treat 0xFFFF... (empty) as the largest value
        t = torch.tensor([-1, 5 + rank, -(2 ** 63) + rank], dtype=torch.int64)
        kdist.allreduce_u64(t, "min")
        assert t.tolist() == [-1, 5, -(2 ** 63)] or t.tolist() == [-1, 5, -(2 ** 63) + 0]
        t2 = torch.tensor([-1 if rank == 0 else 7, 3], dtype=torch.int64)
        kdist.allreduce_u64(t2, "min")
        assert t2.tolist() == [7, 3]
        s = torch.tensor([2 ** 62, rank + 1], dtype=torch.int64)
        kdist.allreduce_u64(s, "sum")
        assert s.tolist()[1] == world * (world + 1) // 2
        gathered = [None] * world
        dist.all_gather_object(gathered, owned)
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


def _unpack(key: int) -> bytes:
    letters = b"ACTG"
    return bytes(letters[(key >> (2 * (15 - i))) & 3] for i in range(16))


def test_owner_exchange_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # union of the owned shards == the whole-file count (regular keys), no key on two ranks
    whole, _ = ko_c.count_fastq(read_golden("test_long.kmer.fastq"))
    exp = {pack(k): v for k, v in whole.items() if all(b in CODE for b in k)}
    merged = {}
    for shard in gathered:
        assert not (set(shard) & set(merged))
        merged.update(shard)
    assert merged == exp
