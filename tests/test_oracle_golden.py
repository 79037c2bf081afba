"""The CPU oracle against every known answer the reference's tests and fixtures hold for this
path, and the C restatement against the Python one."""
import json
import os
import random

import pytest

import kmer_oracle as ko_py
import ko as ko_c

from conftest import read_golden
from util import random_fastq


def as_bytes_map(d):
    return {k.encode("latin-1"): v for k, v in d.items()}


def test_ka1_complement(known):
    ka = known["KA1_complement"]
    assert ko_py.complement(ka["in"].encode()) == ka["out"].encode()
    assert ko_c.complement(ka["in"].encode()) == ka["out"].encode()
    assert ko_py.complement(b"ANxaT\r") == b"\rAaxNT"          # only upper-case ACGT map (lib/kmers.js:32)


def test_ka2_first_key(known):
    ka = known["KA2_first_key"]
    counts = {}
    ko_py.kmers_in_line(ka["line"].encode(), counts)
    assert next(iter(counts)) == ka["first"].encode()


def test_ka3_test_short(known):
    data = read_golden("test_short.fastq")
    exp = [(k.encode(), v) for k, v in known["KA3_test_short"]["map"]]
    for impl in (ko_py, ko_c):
        counts, lines = impl.count_fastq(data)
        assert list(counts.items()) == exp                       # content AND insertion order
        assert lines == 40


def test_ka4_ka5_test_long_kmer(known):
    data = read_golden("test_long.kmer.fastq")
    golden = as_bytes_map(json.loads(read_golden("kmers_long.json")))
    for impl in (ko_py, ko_c):
        counts, lines = impl.count_fastq(data)
        assert len(counts) == known["KA4_test_long_kmer_size"]["size"]
        assert lines == 9000
        # the file is a subset of the reads of test_long.fastq: its map is dominated by the golden map
        assert all(k in golden and v <= golden[k] for k, v in counts.items())


def test_ka6_golden_map_shape(known):
    golden = json.loads(read_golden("kmers_long.json"))
    ka = known["KA6_kmers_long"]
    assert len(golden) == ka["size"] and sum(golden.values()) == ka["sum"]
    assert sum("N" in k for k in golden) == ka["n_keys_with_N"]
    assert all(k.startswith("ATGAC") and len(k) == 16 for k in golden)


def test_ka7_stats_closed_form(known):
    """test/kmerFinderServer.js:70-82 with the integers of db_long_results.json / summary.json."""
    ka = known["KA7_best_match"]
    summary = json.loads(read_golden("summary.json"))
    res = json.loads(read_golden("db_long_results.json"))
    assert res["templateentries"][ka["template"]] == ka["score"]
    assert res["templateentriestot"][ka["template"]] == ka["tScore"]
    for mode in (ko_py.ROUND_HALF_UP, ko_py.ROUND_CEIL):           # client default / after kmerFinderServer.js:7
        ko_py.BNConfig.rounding_mode = mode
        try:
            match = {"uScore": ka["score"], "tScore": ka["tScore"], "ulength": ka["kmers-template"],
                     "lengths": 10000, "species": ka["species"]}
            first = {ka["template"]: dict(match)}
            row = ko_py.match_summary(ka["kmerMapSize"], first, ka["template"], match, ka["hits"], summary)
            assert row is not None
            for f in ("score", "expected", "z", "frac-q", "frac-d", "depth", "kmers-template",
                      "total-frac-q", "total-frac-d", "total-temp-cover"):
                assert row[f] == ka[f], f
            assert row["probability"] == pytest.approx(ka["probability"], rel=1e-12)
        finally:
            ko_py.BNConfig.rounding_mode = ko_py.ROUND_HALF_UP


def test_ka8_hits_invariant(known):
    for name, key in (("db_long_results.json", "db_long"), ("db_short_results.json", "db_short")):
        res = json.loads(read_golden(name))
        assert sum(res["templateentries"].values()) == res["hits"] == known["KA8_hits"][key]


def test_ka9_out_format(known):
    head = known["KA9_out_head"]["head"]
    pairs = [p.split(": ") for p in head[2:].split(",")[:3]]
    text = ko_py.output_file_text({k.encode(): int(v) for k, v in pairs})
    assert head.startswith(text[:-3])                                  # '{\nK: V,K: V,' prefix


def test_digests_are_stable():
    dig = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")))
    import hashlib
    for key, exp in dig.items():
        f, prefix, k, step = key.split("|")
        counts, lines = ko_c.count_fastq(read_golden(f), prefix.encode(), int(k), int(step))
        h = hashlib.sha256()
        for kk in sorted(counts):
            h.update(kk + b"\t" + str(counts[kk]).encode() + b"\n")
        assert (len(counts), sum(counts.values()), lines, h.hexdigest()[:16]) == \
            (exp["unique"], exp["total"], exp["lines"], exp["sha"]), key


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_equals_python_oracle(seed):
    rng = random.Random(seed)
    data = random_fastq(rng, 60, p_n=0.02, p_lower=0.01, crlf=(seed == 1), blank_lines=0.1 if seed == 2 else 0.0,
                        trailing_newline=(seed != 3))
    for prefix, k, step in [(b"ATGAC", 16, 1), (b"", 7, 1), (b"AC", 9, 4), (b"ATGAC", 4, 1)]:
        a, la = ko_py.count_fastq(data, prefix, k, step)
        b, lb = ko_c.count_fastq(data, prefix, k, step)
        assert list(a.items()) == list(b.items()) and la == lb


def test_step_quirk_short_windows():
    """lib/kmers.js:89-99: with step > 1 the loop still runs L-k+1 times; tail windows are clipped."""
    counts = {}
    ko_py.kmers_in_line(b"ACGTACGTAC", counts, k=4, step=3, prefix=b"")
    assert list(counts.items()) == [(b"ACGT", 1), (b"TACG", 1), (b"GTAC", 1), (b"C", 1), (b"", 3)]


def test_wta_oracle_small():
    """Hand-checkable winner-takes-all: T1 wins, its k-mers leave the query, T2 follows."""
    q = {b"AAAA": 2, b"CCCC": 1, b"GGGG": 5, b"TTTT": 1}
    db = ko_py.TemplateDB({b"AAAA": ["T1", "T2"], b"CCCC": ["T1"], b"GGGG": ["T1"], b"TTTT": ["T2"]},
                          {"T1": {"lengths": 100, "ulength": 3, "species": "one"},
                           "T2": {"lengths": 100, "ulength": 2, "species": "two"}},
                          {"templates": 2, "uniqueLens": 1000000, "totalLen": 200})
    templates, hits = ko_py.first_match(q, db)
    assert hits == 5 and list(templates) == ["T1", "T2"]
    assert (templates["T1"]["uScore"], templates["T1"]["tScore"]) == (3, 8)
    rows = []
    # both winners are yielded; the query is then empty and the NEXT getMatches throws
    # (lib/kmerFinderClient.js:264-266 runs before the maxHits/notFound test can end the loop)
    with pytest.raises(RuntimeError, match="nHits === 0"):
        for r in ko_py.find_matches(templates, db.summary, dict(q), len(q)):
            rows.append(r)
    assert [r["template"] for r in rows] == ["T1", "T2"]
    assert rows[1]["score"] == 1 and rows[1]["total-frac-d"] == 100.0 and rows[1]["frac-d"] == 50.0


@pytest.mark.parametrize("seed", range(8))
def test_c_wta_equals_python_oracle(seed):
    """oracle/kmer_oracle.c ko_wta (integer loop in C, exact-decimal gate + rows in Python) against the
    readable restatement kmer_oracle.first_match + find_matches: first-encounter order, first scores, hits,
    every row, the terminating error; seeded DBs with ties, dead templates, maxHits cuts."""
    from collections import OrderedDict
    from util import synthetic_db
    rng = random.Random(400 + seed)
    n_q = rng.choice([30, 200, 1200])
    qmap = OrderedDict()
    while len(qmap) < n_q:
        qmap[bytes(rng.choice(b"ACGT") for _ in range(16))] = rng.randint(1, 9)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=rng.choice([3, 17, 90]), decoys=50,
                                         share=rng.choice([0.1, 0.35, 0.9]))
    if seed == 5:
        # the query is no richer in any template than the DB as a whole: the gate rejects the first winner
        summary = dict(summary, uniqueLens=sum(len(lists.get(q, ())) for q in qmap) + 1)
    for max_hits in (100, 2):
        db = ko_py.TemplateDB(lists, attrs, summary)
        exp_rows, exp_err = [], None
        try:
            templates, hits = ko_py.first_match(OrderedDict(qmap), db)
            first = OrderedDict((n, {"uScore": t["uScore"], "tScore": t["tScore"]}) for n, t in templates.items())
            try:
                for r in ko_py.find_matches(templates, summary, OrderedDict(qmap), len(qmap), max_hits):
                    exp_rows.append(r)
            except RuntimeError as exc:
                exp_err = str(exc)
        except RuntimeError as exc:
            first, hits, exp_err = OrderedDict(), 0, str(exc)
        g_first, g_hits, g_rows, g_err = ko_c.find_matches_fast(OrderedDict(qmap), db, max_hits)
        assert g_hits == hits and list(g_first.items()) == list(first.items())
        assert g_rows == exp_rows and g_err == exp_err, (seed, max_hits)


def test_c_wta_no_hits():
    from collections import OrderedDict
    db = ko_py.TemplateDB({b"AAAA": ["T1"]}, {"T1": {"lengths": 10, "ulength": 1, "species": "s"}},
                          {"templates": 1, "uniqueLens": 10, "totalLen": 10})
    first, hits, rows, err = ko_c.find_matches_fast(OrderedDict([(b"CCCC", 1)]), db)
    assert (hits, rows, err) == (0, [], "No hits were found!")
