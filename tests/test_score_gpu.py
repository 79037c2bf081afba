"""Parity of template scoring on the GPU (kj_first_match / kj_wta_next / kj_standard_scoring)
with the CPU oracle: per-template uScore/tScore, hits, first-encounter order and winner order are
bit-exact; rounded row fields are identical; 'probability' within 1e-9 relative."""
import copy
import json
import random
from collections import OrderedDict

import numpy as np
import pytest

import kmer_oracle as ko
from conftest import read_golden
from util import random_fastq, synthetic_db, emulated as util_emulated

import kmerjs_b200
from kmerjs_b200 import _abi
from kmerjs_b200.counts import Counts
from kmerjs_b200.db import TemplateDB
from kmerjs_b200.kmer_finder_client import counts_from_map
from kmerjs_b200.matching import Match, NoHitsError

pytestmark = pytest.mark.gpu
REL = 1e-9      # tolerance of the one unrounded float field (north star: 1e-9 relative)


def oracle_rows(qmap, kmer_lists, attrs, summary, max_hits=100):
    """(first-match templates, hits, rows, error text or None) from the CPU oracle."""
    db = ko.TemplateDB(kmer_lists, attrs, summary)
    q = OrderedDict(qmap)
    templates, hits = ko.first_match(q, db)
    first = OrderedDict((n, dict(uScore=t["uScore"], tScore=t["tScore"])) for n, t in templates.items())
    rows, err = [], None
    try:
        for r in ko.find_matches(templates, summary, q, len(qmap), max_hits):
            rows.append(r)
    except RuntimeError as exc:
        err = str(exc)
    return first, hits, rows, err, q


def gpu_rows(counts, tdb, max_hits=100):
    m = Match(counts, tdb)
    first = m.templates()
    hits = m.hits
    m.set_max_hits(max_hits)
    rows, err = [], None
    try:
        while True:
            r = m.next_row()
            if r is None:
                break
            rows.append(r)
            assert m.last_row.z_device == pytest.approx(float(ko.z_score(
                r["score"], r["kmers-template"], int(m.last_row.hits), tdb.summary["uniqueLens"]).toNumber()), rel=REL)
    except NoHitsError as exc:
        err = str(exc)
    return first, hits, rows, err, m


def check_rows(got, exp):
    assert [r["template"] for r in got] == [r["template"] for r in exp]        # winner order
    for g, e in zip(got, exp):
        assert list(g.keys()) == ko.ROW_KEYS                                     # JSON key order
        for f in ko.ROW_KEYS:
            if f == "probability":
                assert g[f] == pytest.approx(e[f], rel=REL)
            else:
                assert g[f] == e[f], (g["template"], f)


def test_c2_golden_query_vs_seeded_db():
    """BASELINE config 2 as re-stated in SURVEY.md 8c: query = kmers_long.json (the golden
    findKmers output of test_long.fastq), DB = seeded synthetic templates over those keys."""
    golden = json.loads(read_golden("kmers_long.json"))
    qmap = OrderedDict((k.encode("latin-1"), v) for k, v in golden.items())
    rng = random.Random(2026)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=60, decoys=500)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    counts = counts_from_map(golden, "ATGAC", 16, 1)
    assert counts.size == 6191 and counts.to_dict() == golden and list(counts.to_dict()) == list(golden)
    # the `kmers` Set of every template (lib/kmerFinderServer.js:190-199), in insertion order
    o_templates, _ = ko.first_match(OrderedDict(qmap), ko.TemplateDB(lists, attrs, summary))
    m0 = Match(counts, tdb)
    with_sets = m0.templates(with_kmers=True, keys=list(golden.keys()))
    for name, t in o_templates.items():
        assert [k.encode("latin-1") for k in with_sets[name]["kmers"]] == list(t["kmers"].keys()), name
    m0.free()
    e_first, e_hits, e_rows, e_err, e_q = oracle_rows(qmap, lists, attrs, summary)
    g_first, g_hits, g_rows, g_err, m = gpu_rows(counts, tdb)
    assert g_hits == e_hits == sum(t["uScore"] for t in e_first.values())      # KA8 invariant
    assert list(g_first.keys()) == list(e_first.keys())                          # first-encounter order
    for n in e_first:
        assert (g_first[n]["uScore"], g_first[n]["tScore"]) == (e_first[n]["uScore"], e_first[n]["tScore"]), n
    assert len(e_rows) >= 3
    check_rows(g_rows, e_rows)
    assert g_err == e_err
    # the query after the loop: exactly the winner k-mers are gone (kmerMap.delete)
    alive = counts.alive()
    keys = list(golden.keys())
    assert [k.encode("latin-1") for k, a in zip(keys, alive) if a] == list(e_q.keys())
    # standard scoring (lib/kmerFinderServer.js:857-874) on a fresh match
    m.free(); counts.free()


def test_ka7_reference_row(known):
    """The only stats known-answer in the reference (test/kmerFinderServer.js:70-82), through the
    exact-decimal routine kj_wta_next uses; both bignumber rounding modes."""
    import ctypes as C
    ka = known["KA7_best_match"]
    summary = json.loads(read_golden("summary.json"))
    for mode in (4, 2):
        r, ok = _abi.kj_row(), C.c_int()
        _abi.check(_abi.lib().kj_stats_row(mode, ka["score"], ka["tScore"], ka["score"], ka["tScore"], 10000,
                                           ka["kmers-template"], ka["hits"], ka["kmerMapSize"],
                                           summary["templates"], summary["uniqueLens"], C.byref(r), C.byref(ok)))
        assert ok.value == 1
        assert (r.score, r.expected, r.z, r.frac_q, r.frac_d, r.depth, r.kmers_template) == \
            (ka["score"], ka["expected"], ka["z"], ka["frac-q"], ka["frac-d"], ka["depth"], ka["kmers-template"])
        assert (r.total_frac_q, r.total_frac_d, r.total_temp_cover) == (ka["total-frac-q"], ka["total-frac-d"], 0.36)
        assert r.probability == pytest.approx(ka["probability"], rel=1e-12)


@pytest.mark.parametrize("seed", range(5))
def test_fastq_to_rows_end_to_end(seed, tmp_path):
    """FASTQ -> counts on the GPU -> scoring on the GPU, against oracle count -> oracle scoring;
    irregular (N) k-mers take part in the DB."""
    rng = random.Random(400 + seed)
    data = random_fastq(rng, 250, p_n=0.03 if seed % 2 else 0.0, plant=(b"ATGAC", 0.8), min_len=40)
    qmap, _ = ko.count_fastq(data)
    assert len(qmap) > 50
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=12 + 7 * seed, decoys=50,
                                         share=0.6)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    path = tmp_path / "reads.fastq"
    path.write_bytes(data)
    client = kmerjs_b200.KmerFinderClient(str(path), "node", "ATGAC", 16, 1, 1, False, tdb)
    kmap = client.findKmers().promise.result(timeout=300)
    assert {k.encode("latin-1"): v for k, v in kmap.items()} == dict(qmap) and \
        [k.encode("latin-1") for k in kmap] == list(qmap)
    e_first, e_hits, e_rows, e_err, e_q = oracle_rows(qmap, lists, attrs, summary)
    winner = client.findFirstMatch(kmap).result(timeout=300)
    assert winner["hits"] == e_hits and list(winner["templates"]) == list(e_first)
    assert winner["summary"] == summary
    got, err = [], None
    try:
        for row in client.findMatches(winner, kmap):
            got.append(row)
    except NoHitsError as exc:
        err = str(exc)
    check_rows(got, e_rows)
    assert err == e_err
    # the caller's map lost the winner k-mers, like the reference's Map
    assert [k.encode("latin-1") for k in kmap if k not in ("db", "collection")] == list(e_q.keys())
    client.close()


def test_no_hits_errors():
    qmap = {"ATGACAAAAAAAAAAA": 3, "ATGACCCCCCCCCCCC": 1}
    lists = {b"ATGACGGGGGGGGGGG": ["T1"]}
    attrs = {"T1": {"lengths": 1000, "ulength": 10, "species": "s"}}
    summary = {"templates": 1, "uniqueLens": 10, "totalLen": 1000}
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    c = counts_from_map(qmap, "ATGAC", 16, 1)
    with pytest.raises(NoHitsError, match=r"^No hits were found!$"):           # lib/kmerFinderClient.js:161
        Match(c, tdb)
    # a match whose best template fails the evalue gate: zero rows -> the second error text
    lists = {b"ATGACAAAAAAAAAAA": ["T1"]}
    attrs = {"T1": {"lengths": 1000, "ulength": 100000, "species": "s"}}
    summary = {"templates": 5000, "uniqueLens": 100000, "totalLen": 1000}
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    e = oracle_rows(OrderedDict((k.encode(), v) for k, v in qmap.items()), lists, attrs, summary)
    assert e[2] == [] and "kmerResults.length === 0" in e[3]
    m = Match(c, tdb)
    with pytest.raises(NoHitsError, match=r"kmerResults\.length === 0"):       # lib/kmerFinderClient.js:284
        m.next_row()
    m.free(); c.free()


def test_max_hits_and_ties():
    rng = random.Random(77)
    keys = [bytes(b"ATGAC") + bytes(rng.choice(b"ACGT") for _ in range(11)) for _ in range(400)]
    qmap = OrderedDict((k, rng.randint(1, 9)) for k in dict.fromkeys(keys))
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=30, decoys=10, share=0.9)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    for max_hits in (1, 3, 100):
        c = counts_from_map({k.decode(): v for k, v in qmap.items()}, "ATGAC", 16, 1)
        e_first, e_hits, e_rows, e_err, _ = oracle_rows(qmap, lists, attrs, summary, max_hits)
        g_first, g_hits, g_rows, g_err, m = gpu_rows(c, tdb, max_hits)
        assert list(g_first) == list(e_first) and g_hits == e_hits
        check_rows(g_rows, e_rows)
        assert g_err == e_err and len(g_rows) <= max_hits
        m.free(); c.free()


def test_deferred_rows_equal_immediate_rows():
    """kj_match_defer_rows: kj_wta_next hands back the integers, kj_wta_row finishes the row later (the
    multi-GPU layer puts its all-reduce in between); the rows must be the same."""
    rng = random.Random(123)
    keys = [bytes(b"ATGAC") + bytes(rng.choice(b"ACGT") for _ in range(11)) for _ in range(600)]
    qmap = OrderedDict((k, rng.randint(1, 5)) for k in dict.fromkeys(keys))
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=16, decoys=20, share=0.8)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    _, _, e_rows, e_err, _ = oracle_rows(qmap, lists, attrs, summary)
    c = counts_from_map({k.decode(): v for k, v in qmap.items()}, "ATGAC", 16, 1)
    m = Match(c, tdb)
    m.defer_rows(True)
    got, err, pending = [], None, 0
    try:
        while True:
            state, row = m.next_row_begin()
            if state == 0:
                break
            if state == 2:
                pending += 1
                row = m.finish_row()
            got.append(row)
    except NoHitsError as exc:
        err = str(exc)
    check_rows(got, e_rows)
    assert err == e_err and pending >= 1
    m.free(); c.free()


@pytest.mark.parametrize("seed,max_hits", [(123, 100), (7, 3), (31, 1)])
def test_all_rows_equal_the_generator(seed, max_hits):
    """kj_wta_all (the whole loop in one call) against the stepwise generator and the oracle:
    same rows, same terminal error, also when maxHits cuts the loop short."""
    rng = random.Random(seed)
    keys = [bytes(b"ATGAC") + bytes(rng.choice(b"ACGT") for _ in range(11)) for _ in range(600)]
    qmap = OrderedDict((k, rng.randint(1, 5)) for k in dict.fromkeys(keys))
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=16, decoys=20, share=0.8)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    c = counts_from_map({k.decode(): v for k, v in qmap.items()}, "ATGAC", 16, 1)
    # the stepwise path on one handle, the one-call path on another
    m1 = Match(c, tdb)
    m1.set_max_hits(max_hits)
    step_rows, step_err = [], None
    try:
        while True:
            r = m1.next_row()
            if r is None:
                break
            step_rows.append(r)
    except NoHitsError as exc:
        step_err = str(exc)
    m1.free(); c.free()
    c = counts_from_map({k.decode(): v for k, v in qmap.items()}, "ATGAC", 16, 1)
    m2 = Match(c, tdb)
    rows, err = m2.all_rows(max_hits)
    assert rows == step_rows and (str(err) if err else None) == step_err
    assert len(rows) <= max_hits
    if max_hits == 100:
        _, _, e_rows, e_err, _ = oracle_rows(qmap, lists, attrs, summary)
        check_rows(rows, e_rows)
        assert (str(err) if err else None) == e_err
    with pytest.raises(_abi.KjError):       # a row buffer smaller than maxHits is refused
        m2.set_max_hits(100)
        buf = (_abi.kj_row * 2)()
        import ctypes as C
        n, end = C.c_uint32(), C.c_int()
        _abi.check(_abi.lib().kj_wta_all(m2.handle, buf, 2, C.byref(n), C.byref(end)), m2.ctx.handle)
    m2.free(); c.free()


def test_standard_scoring():
    golden = json.loads(read_golden("test_long.json"))          # 6045-key sorted subset of the golden map
    qmap = OrderedDict((k.encode("latin-1"), v) for k, v in golden.items())
    rng = random.Random(9)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=25, decoys=100)
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    c = counts_from_map(golden, "ATGAC", 16, 1)
    m = Match(c, tdb)
    got = m.standard_scoring()
    # oracle: matchSummary of every first-match template, stable sort by score desc
    # (lib/kmerFinderServer.js:857-874,684-693)
    db = ko.TemplateDB(lists, attrs, summary)
    templates, hits = ko.first_match(OrderedDict(qmap), db)
    exp = []
    for name, t in templates.items():
        r = ko.match_summary(len(qmap), templates, name, t, hits, summary)
        if r is not None:
            exp.append(r)
    exp.sort(key=lambda r: -r["score"])
    check_rows(got, exp)
    m.free(); c.free()


def test_device_stats_match_exact_decimal(ctx):
    import ctypes as C
    rng = np.random.default_rng(5)
    n = 512
    n1 = rng.integers(1, 2_000_000, n).astype(np.uint64)
    r1 = (rng.random(n) * n1).astype(np.uint64)
    n2 = rng.integers(1_000, 50_000_000, n).astype(np.uint64)
    r2 = (rng.random(n) * np.minimum(n2, 500_000)).astype(np.uint64)
    z = np.zeros(n); p = np.zeros(n)
    _abi.check(_abi.lib().kj_stats_zscore_device(ctx.handle, n, r1.ctypes.data, n1.ctypes.data, r2.ctypes.data,
                                                 n2.ctypes.data, z.ctypes.data, p.ctypes.data), ctx.handle)
    for i in range(n):
        ez = ko.z_score(int(r1[i]), int(n1[i]), int(r2[i]), int(n2[i]))
        assert z[i] == pytest.approx(ez.toNumber(), rel=REL, abs=1e-12)
        # the step function may only differ when z sits within the tolerance of a threshold
        if all(abs(z[i] - thr) > 1e-6 for thr, _ in ko._FASTP_TABLE):
            assert p[i] == ko.fastp(ez).toNumber()


def test_exact_decimal_host_functions():
    """kj_stats_zscore / kj_stats_fastp_text (host-only entry points) against the oracle's BN."""
    rng = random.Random(1)
    for mode in (4, 2, 6):
        ko.BNConfig.rounding_mode = mode
        kmerjs_b200.stats.set_rounding_mode(mode)
        try:
            for _ in range(60):
                n1 = rng.randint(1, 10**6); r1 = rng.randint(0, n1)
                n2 = rng.randint(1, 10**8); r2 = rng.randint(0, min(n2, 10**6))
                ez = ko.z_score(r1, n1, r2, n2)
                gz = kmerjs_b200.zScore(r1, n1, r2, n2)
                assert gz == __import__("decimal").Decimal(ez.n).scaleb(-ez.e)
                assert float(kmerjs_b200.fastp(gz)) == ko.fastp(ez).toNumber()
        finally:
            ko.BNConfig.rounding_mode = 4
            kmerjs_b200.stats.set_rounding_mode(4)


@pytest.mark.parametrize("max_hits", [100, 7])
def test_ten_thousand_templates_vs_oracle(max_hits):
    """BASELINE config 4's shape at a size the oracle follows in seconds: 10 000 templates in genera that
    share k-mers, > 1e6 (k-mer, template) pairs, a query that hits 3 000 of them.  T > 8192 takes the
    global-atomic walk (kj_walk_kernel<ACCUM, false>); the loop is cut by maxHits; uScore ties are decided by
    first-encounter order.  Oracle: the reference's full recount per round (oracle/kmer_oracle.c ko_wta,
    pinned to the Python restatement in test_oracle_golden.py) with the exact-decimal gate and rows."""
    import ko as ko_c
    from kmerjs_b200 import synth
    T = 10_000 if not util_emulated() else 600
    rng = np.random.default_rng(4242)
    pk = 0b0010110001                                                   # ATGAC
    sample = (np.uint64(pk) << np.uint64(22)) | rng.integers(0, 1 << 22, 150, dtype=np.uint64)
    tdb = synth.genus_template_db(sample, T, 135, seed=3)
    assert T < 10_000 or (tdb.tmpl_ids.size > 1_000_000 and tdb.n_templates > 8192)
    # the query: 70 % of the k-mers of the first 30 % of the genera (enriched in them, like a sample is in its
    # relatives) plus k-mers the DB does not hold, in random (Map) order
    n_db = tdb.n_kmers
    pick = np.nonzero(rng.random(int(0.3 * n_db)) < 0.7)[0]
    miss = (np.uint64(pk) << np.uint64(22)) | rng.integers(0, 1 << 22, 5000, dtype=np.uint64)
    miss = miss[~np.isin(miss, tdb.keys_u64)]
    qkeys = np.concatenate([tdb.keys_u64[pick], np.unique(miss)])
    qidx = np.concatenate([pick, np.full(qkeys.size - pick.size, -1)])
    order = rng.permutation(qkeys.size)
    qkeys, qidx = qkeys[order], qidx[order]
    qcount = rng.integers(1, 40, qkeys.size).astype(np.uint64)
    rec = np.stack([qkeys, qcount, np.arange(qkeys.size, dtype=np.uint64)], axis=1)
    c = Counts(b"ATGAC", 16, 1)
    c.merge_host_records(rec)
    c.finish()
    assert c.size == qkeys.size
    # oracle arrays: every query entry with its DB list (DB order)
    off = tdb.list_off.astype(np.int64)
    lens = np.where(qidx >= 0, off[np.maximum(qidx, 0) + 1] - off[np.maximum(qidx, 0)], 0)
    qoff = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    starts = np.repeat(off[np.maximum(qidx, 0)], lens)
    within = np.arange(int(lens.sum())) - np.repeat(qoff[:-1].astype(np.int64), lens)
    qt = tdb.tmpl_ids[starts + within]
    attrs = {n: {"lengths": int(tdb.lengths[i]), "ulength": int(tdb.ulengths[i]), "species": tdb.species[i]}
             for i, n in enumerate(tdb.names)}
    e_first, e_hits, e_rows, e_err = ko_c.wta_arrays(qcount, qoff, qt, tdb.names, attrs, tdb.summary,
                                                     int(qkeys.size), max_hits)
    m = Match(c, tdb)
    g_first = m.templates()
    assert m.hits == e_hits == int(lens.sum())
    assert list(g_first.keys()) == list(e_first.keys())                       # first-encounter order, 10 k templates
    assert all((g_first[n]["uScore"], g_first[n]["tScore"]) == (e_first[n]["uScore"], e_first[n]["tScore"])
               for n in e_first)
    rows, err = m.all_rows(max_hits)
    assert T < 10_000 or len(e_rows) == max_hits                              # the loop is cut by maxHits
    check_rows(rows, e_rows)
    assert (str(err) if err else None) == e_err
    us = [r["score"] for r in e_rows]
    assert T < 10_000 or max_hits < 100 or len(set(us)) < len(us)             # ties took part
    m.free(); c.free()


@pytest.mark.parametrize("fmt", ["kmer_docs", "redis_strings", "template_docs", "kmerfinder_map", "packed"])
def test_native_db_load_scores_like_the_oracle(fmt, tmp_path):
    """kj_db_load (C ABI): each on-disk layout of the reference (lib/kmerFinderServer.js:68-92,184-199,
    src/kmerPyToMongo.py:35-42, lib/index.js:184-192) and the packed binary, file -> GPU, scored against the oracle
    fed with the same database."""
    from kmerjs_b200.db import load as load_db, load_native
    golden = json.loads(read_golden("test_long.json"))
    qmap = OrderedDict((k.encode("latin-1"), v) for k, v in golden.items())
    rng = random.Random(31)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=20, decoys=60)
    names = list(attrs)
    spath = tmp_path / "summary.json"
    spath.write_text(json.dumps(summary))
    path = tmp_path / ("db." + ("kjdb" if fmt == "packed" else "json"))
    if fmt in ("kmer_docs", "redis_strings"):
        rec = lambda t: {"sequence": t, "lengths": attrs[t]["lengths"], "ulengths": attrs[t]["ulength"], "species": attrs[t]["species"]}
        docs = [{"kmer": k.decode("latin-1"), "templates": [json.dumps(rec(t)) if fmt == "redis_strings" else rec(t) for t in lst]}
                for k, lst in lists.items()]
        path.write_text(json.dumps(docs))
        host = load_db(str(path), str(spath))
    elif fmt == "template_docs":
        docs = [{"sequence": t, "lengths": attrs[t]["lengths"], "ulenght": attrs[t]["ulength"], "species": attrs[t]["species"],
                 "reads": [k.decode("latin-1") for k, lst in lists.items() if t in lst]} for t in names]
        path.write_text(json.dumps(docs))
        host = load_db(str(path), str(spath))
    elif fmt == "kmerfinder_map":
        path.write_text(json.dumps({k.decode("latin-1"): ",".join(lst) for k, lst in lists.items()}))
        side = {"lengths": {t: attrs[t]["lengths"] for t in names}, "ulengths": {t: attrs[t]["ulength"] for t in names},
                "descriptions": {t: attrs[t]["species"] for t in names}}
        for suffix, d in side.items():
            (tmp_path / f"db.json.{suffix}.json").write_text(json.dumps(d))
        host = load_db(str(path), str(spath), **side)
    else:
        TemplateDB.from_lists(lists, attrs, summary).save(str(path))
        host = load_db(str(path))
    native = load_native(str(path), None if fmt == "packed" else str(spath))
    assert native.summary == summary and native.n_templates == host.n_templates
    assert sorted(native.names) == sorted(host.names)
    o_lists, o_attrs = host.to_lists()                       # the oracle sees what the host loader read from the same file
    e_first, e_hits, e_rows, e_err, _ = oracle_rows(qmap, o_lists, o_attrs, summary)
    counts = counts_from_map(golden, "ATGAC", 16, 1)
    g_first, g_hits, g_rows, g_err, m = gpu_rows(counts, native)
    assert g_hits == e_hits and list(g_first.keys()) == list(e_first.keys())
    assert all((g_first[n]["uScore"], g_first[n]["tScore"]) == (e_first[n]["uScore"], e_first[n]["tScore"]) for n in e_first)
    assert len(e_rows) >= 2
    check_rows(g_rows, e_rows)
    assert g_err == e_err
    m.free(); counts.free()
