"""The multi-GPU protocol of the C ABI, with the ranks emulated one after the other on one device
(no collectives here: the reductions NCCL would do are done with numpy): byte-range sharding with
newline phase, owner partition + merge, sharded DB, reduced score vectors, replicated WTA.
The result must equal the single-device run and the oracle."""
import ctypes as C
import json
import random
from collections import OrderedDict

import numpy as np
import pytest

import kmer_oracle as ko
import ko as ko_c
from conftest import read_golden
from util import DevBuf, dev_u64, emulated, random_fastq, synthetic_db

from kmerjs_b200 import _abi
from kmerjs_b200 import dist as kdist
from kmerjs_b200.counts import Counts, count_newlines_device
from kmerjs_b200.db import TemplateDB
from kmerjs_b200.matching import Match, NoHitsError

pytestmark = pytest.mark.gpu


def read_device(ptr, n_bytes):
    if emulated():
        return np.frombuffer(C.string_at(ptr, n_bytes), dtype=np.uint8).copy()
    import torch
    out = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    C.cdll.LoadLibrary  # noqa: B018
    # copy through torch: build a view with the CUDA array interface
    view = torch.as_tensor(kdist._CudaView(ptr, n_bytes // 8), device="cuda")
    return view.cpu().numpy().view(np.uint8).copy()


def sharded_counts(data, world, prefix=b"ATGAC", k=16, step=1, halo=64):
    """count every rank's range, partition by owner, merge per owner -> list of owned Counts."""
    ranges = kdist.plan_ranges(len(data), world, halo=halo)
    whole = DevBuf(data)
    stats = [count_newlines_device(whole.ptr + lo, own) for lo, own, _ in ranges]
    for (lo, own, _), (cnt, last) in zip(ranges, stats):
        seg = data[lo:lo + own]
        assert cnt == seg.count(b"\n") and last == seg.rfind(b"\n") + 1
    bl, bc = kdist.phase_of_ranges([s[0] for s in stats], [s[1] for s in stats], [r[1] for r in ranges])
    locals_ = []
    for r, (lo, own, rd) in enumerate(ranges):
        c = Counts(prefix, k, step, flags=_abi.KJ_F_COUNT_BASES, base_line=bl[r], base_col=bc[r])
        c.add_device(whole.ptr + lo, rd, own_n=own, final=(r == world - 1))
        c.finish()
        locals_.append(c)
    parts = [c.partition(world) for c in locals_]
    owned = []
    for o in range(world):
        oc = Counts(prefix, k, step)
        for r, (ptr, sizes) in enumerate(parts):
            off = sum(sizes[:o])
            if sizes[o]:
                oc.merge_records(ptr + 24 * off, sizes[o])
        for c in locals_:                     # byte-string k-mers: every owner is shown all, keeps its own
            oc.merge_irregular(c.irregular_records(), o, world)
        oc.finish()
        owned.append(oc)
    totals = (max(c.lines for c in locals_), sum(c.bases for c in locals_), sum(c.occurrences for c in locals_))
    for c in locals_:
        c.free()
    return owned, totals, whole


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_count_equals_whole(world):
    rng = random.Random(50 + world)
    data = random_fastq(rng, 500, p_n=0.02, blank_lines=0.03)
    exp, lines = ko_c.count_fastq(data)
    owned, totals, _keep = sharded_counts(data, world)
    merged = {}
    for o, oc in enumerate(owned):
        d = oc.to_dict()
        assert not (set(d) & set(merged))                         # every key has one owner
        for kk in d:
            assert _abi.lib().kj_owner(kk.encode("latin-1"), len(kk), world) == o
        merged.update(d)
    assert {k.encode("latin-1"): v for k, v in merged.items()} == dict(exp)
    assert totals[0] == lines                        # the last rank's count includes the lines before it
    assert totals[1] == sum(len(l) for i, l in enumerate(ko.split_lines(data)) if i % 4 == 1)
    assert totals[2] == sum(exp.values())
    for oc in owned:
        oc.free()


def test_sharded_scoring_equals_single_device():
    golden_reads = read_golden("test_long.kmer.fastq")
    world = 2
    exp_counts, _ = ko_c.count_fastq(golden_reads)
    rng = random.Random(314)
    lists, attrs, summary = synthetic_db(list(exp_counts.keys()), rng, n_templates=20, decoys=80, share=0.7)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    # oracle
    db = ko.TemplateDB(lists, attrs, summary)
    q = OrderedDict(exp_counts)
    templates, hits = ko.first_match(q, db)
    e_order = list(templates.keys())
    e_rows, e_err = [], None
    try:
        for r in ko.find_matches(templates, summary, q, len(exp_counts)):
            e_rows.append(r)
    except RuntimeError as exc:
        e_err = str(exc)
    # sharded run
    owned, totals, _keep = sharded_counts(golden_reads, world)
    qsize = sum(oc.size for oc in owned)
    assert qsize == len(exp_counts)
    ms = [Match(oc, tdb, local_only=True, part=r, n_parts=world) for r, oc in enumerate(owned)]

    def reduce(which, op):
        n = ms[0].vec_len(which)
        vals = []
        for m in ms:
            ptr, keep, back = dev_u64(np.zeros(max(n, 1), dtype=np.uint64))
            m.get(which, ptr)
            vals.append(back()[:n])
        red = np.sum(vals, axis=0, dtype=np.uint64) if op == "sum" else np.min(vals, axis=0)
        for m in ms:
            ptr, keep, _ = dev_u64(red if n else np.zeros(1, dtype=np.uint64))
            m.set(which, ptr)

    reduce(_abi.KJ_VEC_SCORES, "sum")
    reduce(_abi.KJ_VEC_FIRST_ORD, "min")
    reduce(_abi.KJ_VEC_FIRST_IDX, "min")
    for m in ms:
        m.set_query_size(qsize)
        m.commit()
        assert m.hits == hits and list(m.templates().keys()) == e_order
        assert {n: (t["uScore"], t["tScore"]) for n, t in m.templates().items()} == \
            {n: (t["uScore"], t["tScore"]) for n, t in templates.items()}
    g_rows, g_err = [], None
    try:
        while True:
            rows = [m.next_row() for m in ms]                   # every rank takes the same decision
            assert all(r == rows[0] for r in rows)
            if rows[0] is None:
                break
            g_rows.append(rows[0])
            reduce(_abi.KJ_VEC_SCORES, "sum")
    except NoHitsError as exc:
        g_err = str(exc)
    assert [r["template"] for r in g_rows] == [r["template"] for r in e_rows]
    for g, e in zip(g_rows, e_rows):
        for f in ko.ROW_KEYS:
            if f == "probability":
                assert g[f] == pytest.approx(e[f], rel=1e-9)
            else:
                assert g[f] == e[f], f
    assert g_err == e_err
    for m in ms:
        m.free()
    for oc in owned:
        oc.free()


@pytest.mark.parametrize("world", [1, 3])
def test_gathered_match_equals_single_device(world):
    """kj_match_matched_size / export_matched / from_matched: the matched entries of every (emulated) rank,
    gathered, give the rows, hits and first-encounter order of the single-device run -- and of the oracle."""
    golden_reads = read_golden("test_long.kmer.fastq")
    exp_counts, _ = ko_c.count_fastq(golden_reads)
    rng = random.Random(2718)
    lists, attrs, summary = synthetic_db(list(exp_counts.keys()), rng, n_templates=20, decoys=80, share=0.7)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    db = ko.TemplateDB(lists, attrs, summary)
    q = OrderedDict(exp_counts)
    templates, hits = ko.first_match(q, db)
    e_rows, e_err = [], None
    try:
        for r in ko.find_matches(templates, summary, q, len(exp_counts)):
            e_rows.append(r)
    except RuntimeError as exc:
        e_err = str(exc)
    owned, totals, _keep = sharded_counts(golden_reads, world)
    qsize = sum(oc.size for oc in owned)
    ms = [Match(oc, tdb, local_only=True, part=r, n_parts=world) for r, oc in enumerate(owned)]
    sizes = [m.matched_size() for m in ms]
    assert sum(p for _, p in sizes) == hits
    segs, keep = [], []
    for m, (ne, npairs) in zip(ms, sizes):
        pe, ke, back_e = dev_u64(np.zeros(max(4 * ne, 4), dtype=np.uint64))
        pt, kt, back_t = dev_u64(np.zeros(max((npairs + 1) // 2, 1), dtype=np.uint64))
        m.export_matched(pe, ne, pt, npairs)
        assert m.matched_size() == (ne, npairs)                 # sizing is repeatable (and synchronises)
        ent = back_e()[:4 * ne].reshape(-1, 4)
        # every entry's list lies inside the rank's template array and the lists tile it exactly
        spans = sorted((int(o), int(o + l)) for _, _, o, l in ent)
        pos = 0
        for lo, hi in spans:
            assert lo == pos
            pos = hi
        assert pos == npairs
        tm = back_t().view(np.uint32)[:npairs]
        assert tm.size == 0 or int(tm.max()) < tdb.n_templates
        segs.append((pe, ne, pt, npairs))
        keep += [ke, kt]
    g = Match.from_matched(ms[0].ctx, tdb, segs, qsize, part=0, n_parts=world)
    g.commit()
    assert g.hits == hits and list(g.templates().keys()) == list(templates.keys())
    assert {n: (t["uScore"], t["tScore"]) for n, t in g.templates().items()} == \
        {n: (t["uScore"], t["tScore"]) for n, t in templates.items()}
    with pytest.raises(_abi.KjError):
        g.template_kmers(0)
    g_rows, g_err = [], None
    try:
        while True:
            r = g.next_row()
            if r is None:
                break
            g_rows.append(r)
    except NoHitsError as exc:
        g_err = str(exc)
    assert [r["template"] for r in g_rows] == [r["template"] for r in e_rows]
    for a, e in zip(g_rows, e_rows):
        for f in ko.ROW_KEYS:
            if f == "probability":
                assert a[f] == pytest.approx(e[f], rel=1e-9)
            else:
                assert a[f] == e[f], f
    assert g_err == e_err
    # a record pointing outside its segment is refused
    if sizes[0][0]:
        bad = np.zeros(4, dtype=np.uint64)
        bad[2], bad[3] = 0, sizes[0][1] + 1
        pb, kb, _ = dev_u64(bad)
        with pytest.raises(_abi.KjError):
            Match.from_matched(ms[0].ctx, tdb, [(pb, 1, segs[0][2], sizes[0][1])], qsize)
    g.free()
    for m in ms:
        m.free()
    for oc in owned:
        oc.free()


def test_matched_segment_straight_from_the_table():
    """kj_counts_export_matched_segment: the matched entries of a handle that has NOT been finished, read from its hash
    table, give the same gathered match (hits, first-encounter order, scores, rows) as the finished handle and the
    oracle; the handle can still be finished afterwards.  A DB with byte-string k-mers does not take the short cut."""
    from kmerjs_b200.context import default_context
    from kmerjs_b200.matching import matched_segment_bytes
    golden_reads = read_golden("test_long.kmer.fastq")
    exp_counts, _ = ko_c.count_fastq(golden_reads)
    rng = random.Random(1618)
    regular = [k for k in exp_counts.keys() if set(k) <= set(b"ACGT") and len(k) == 16]
    lists, attrs, summary = synthetic_db(regular, rng, n_templates=20, decoys=80, share=0.7)
    assert all(set(k) <= set(b"ACGT") and len(k) == 16 for k in lists)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    db = ko.TemplateDB(lists, attrs, summary)
    q = OrderedDict(exp_counts)
    templates, hits = ko.first_match(q, db)
    e_rows, e_err = [], None
    try:
        for r in ko.find_matches(templates, summary, q, len(exp_counts)):
            e_rows.append(r)
    except RuntimeError as exc:
        e_err = str(exc)
    ctx = default_context()
    c = Counts(b"ATGAC", 16, 1)
    c.add_host(golden_reads, final=True)                       # counted, not finished
    cap_e, cap_p = 1024, 16384
    p, keep, _back = dev_u64(np.zeros(matched_segment_bytes(cap_e, cap_p) // 8 + 1, dtype=np.uint64))
    assert c.export_matched_segment(tdb.device(ctx, 0, 1), p, cap_e, cap_p)
    g = Match.from_segments(ctx, tdb, 1, p, cap_e, cap_p)
    g.commit()
    assert g.query_size == len(exp_counts)
    assert g.hits == hits and list(g.templates().keys()) == list(templates.keys())
    assert {n: (t["uScore"], t["tScore"]) for n, t in g.templates().items()} == \
        {n: (t["uScore"], t["tScore"]) for n, t in templates.items()}
    g_rows, g_err = [], None
    try:
        while True:
            r = g.next_row()
            if r is None:
                break
            g_rows.append(r)
    except NoHitsError as exc:
        g_err = str(exc)
    assert [r["template"] for r in g_rows] == [r["template"] for r in e_rows]
    for a, e in zip(g_rows, e_rows):
        for f in ko.ROW_KEYS:
            if f == "probability":
                assert a[f] == pytest.approx(e[f], rel=1e-9)
            else:
                assert a[f] == e[f], f
    assert g_err == e_err
    g.free()
    c.finish()
    assert c.size == len(exp_counts)
    # a segment that is too small: nothing may follow a list outside it (the walk is queued before the sizes are known to
    # the host), and the commit refuses the match
    c3 = Counts(b"ATGAC", 16, 1)
    c3.add_host(golden_reads, final=True)
    small_e, small_p = 64, 8
    p2, keep2, _b2 = dev_u64(np.zeros(matched_segment_bytes(small_e, small_p) // 8 + 1, dtype=np.uint64))
    assert c3.export_matched_segment(tdb.device(ctx, 0, 1), p2, small_e, small_p)
    g2 = Match.from_segments(ctx, tdb, 1, p2, small_e, small_p)
    with pytest.raises(_abi.KjError) as ei:
        g2.commit()
    assert ei.value.code == _abi.KJ_E_RANGE
    g2.free()
    c3.free()
    del keep2
    # a DB that holds a byte-string k-mer (here: one with N) must be matched the long way
    lists_n = OrderedDict(lists)
    lists_n[b"ATGACNNNNNNNNNNN"] = [next(iter(attrs))] if isinstance(attrs, dict) else ["T0000"]
    tdb_n = TemplateDB.from_lists(lists_n, attrs, summary)
    c2 = Counts(b"ATGAC", 16, 1)
    c2.add_host(golden_reads, final=True)
    assert not c2.export_matched_segment(tdb_n.device(ctx, 0, 1), p, cap_e, cap_p)
    c2.free()
    c.free()
    del keep


@pytest.mark.parametrize("prefix,k", [(b"ATGAC", 16), (b"", 31)])
def test_fixed_capacity_exchange_emulated(prefix, k):
    """kj_counts_partition_segments / kj_counts_merge_segments with the ranks emulated on one device: every rank scatters
    its (unfinished) table into one fixed-capacity segment per owner, the owners merge what they are sent and finish; the
    union of the owners' maps is the single-device map, totals included.  Reads with N: byte-string k-mers travel too.
    Then with segments that are too small: an owner that was sent an overflowed segment must refuse in finish (KJ_E_RANGE),
    never return a partial map."""
    from kmerjs_b200.counts import segment_bytes
    rng = random.Random(99)
    n_reads = 60 if emulated() else 2000
    data = random_fastq(rng, n_reads, min_len=60, max_len=150, p_n=0.02)
    world = 3
    single = Counts(prefix, k, 1)
    single.add_host(data, final=True)
    single.finish()
    exp = single.to_dict()
    for cap_reg, cap_irr, fits in ((1 << 17, 1 << 15, True), (8, 2, False)):     # 2000 dense reads: ~45 k records per segment
        ranges = kdist.plan_ranges(len(data), world, halo=64)
        whole = DevBuf(data)
        stats = [count_newlines_device(whole.ptr + lo, own) for lo, own, _ in ranges]
        bl, bc = kdist.phase_of_ranges([s_[0] for s_ in stats], [s_[1] for s_ in stats], [r[1] for r in ranges])
        seg = segment_bytes(cap_reg, cap_irr)
        sent, locals_, keep = [], [], []
        for r, (lo, own, rd) in enumerate(ranges):
            c = Counts(prefix, k, 1, base_line=bl[r], base_col=bc[r])
            c.add_device(whole.ptr + lo, rd, own_n=own, final=(r == world - 1))      # not finished
            p, kp, back = dev_u64(np.zeros(world * seg // 8, dtype=np.uint64))
            c.partition_segments(world, p, cap_reg, cap_irr)
            c.finish()                                                               # (waits for the scatter)
            sent.append(back().view(np.uint8).reshape(world, seg))
            locals_.append(c)
            keep.append(kp)
        merged, failed = {}, 0
        lines = occ = 0
        for o in range(world):
            recv = np.concatenate([sent[r][o] for r in range(world)])
            p, kp, _ = dev_u64(recv.view(np.uint64))
            oc = Counts(prefix, k, 1, capacity_hint=world * cap_reg)
            oc.merge_segments(p, world, cap_reg, cap_irr)
            try:
                oc.finish()
            except _abi.KjError as exc:
                assert exc.code == _abi.KJ_E_RANGE
                failed += 1
                oc.free()
                continue
            d = oc.to_dict()
            assert not (set(d) & set(merged))
            merged.update(d)
            lines, occ = oc.lines, oc.occurrences
            oc.free()
            del kp
        if fits:
            assert failed == 0 and merged == exp
            assert (lines, occ) == (single.lines, single.occurrences)
        else:
            # an owner whose segments all fitted is complete for the k-mers it owns; the others refused
            assert failed >= 1 and all(exp.get(key) == v for key, v in merged.items())
        for c in locals_:
            c.free()
    single.free()
