"""Wire formats and command line around the path against the oracle: the findFirstMatch exchange
(lib/kmerFinderClient.js:132-161), POST /kmers (server/app.js:22-54), the CLI flags (lib/cli.js:9-20)."""
import json
import random
from collections import OrderedDict

import pytest

import kmer_oracle as ko
from conftest import read_golden
from util import random_fastq, synthetic_db

from kmerjs_b200 import cli, wire
from kmerjs_b200.db import TemplateDB
from kmerjs_b200.matching import NoHitsError

pytestmark = pytest.mark.gpu


def _setup(seed=3):
    golden = json.loads(read_golden("test_long.json"))
    qmap = OrderedDict((k.encode("latin-1"), v) for k, v in golden.items())
    rng = random.Random(seed)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=18, decoys=40)
    return golden, qmap, lists, attrs, summary


def test_first_match_exchange_and_post_kmers():
    golden, qmap, lists, attrs, summary = _setup()
    tdb = TemplateDB.from_lists(lists, attrs, summary)
    body = wire.first_match_request(golden, "Kmers", "genomes")
    doc = json.loads(body)
    assert list(doc)[-2:] == ["db", "collection"] and list(doc)[:-2] == list(golden)       # Map order, bookkeeping keys last
    status, reply = wire.first_match_reply(body, tdb)
    assert status == 200
    winner = wire.parse_first_match_reply(status, reply)
    o_templates, o_hits = ko.first_match(OrderedDict(qmap), ko.TemplateDB(lists, attrs, summary))
    assert winner["hits"] == o_hits and winner["summary"] == summary
    assert list(winner["templates"]) == list(o_templates)
    for name, t in o_templates.items():
        w = winner["templates"][name]
        assert (w["uScore"], w["tScore"], w["lengths"], w["ulength"], w["species"]) == \
            (t["uScore"], t["tScore"], t["lengths"], t["ulength"], t["species"])
        assert [k.encode("latin-1") for k in w["kmers"]] == list(t["kmers"])
    # nothing hits: 204 -> 'No hits were found!'
    status, reply = wire.first_match_reply(json.dumps({"ATGACAAAAAAAAAAT": 2}), tdb)
    assert status == 204
    with pytest.raises(NoHitsError, match=r"^No hits were found!$"):
        wire.parse_first_match_reply(status, reply)
    # POST /kmers: the rows of the winner-takes-all loop, the reference handler's fields
    status, reply = wire.post_kmers(json.dumps(golden), tdb)
    rows = json.loads(reply)
    e_rows = []
    try:
        for r in ko.find_matches(o_templates, summary, OrderedDict(qmap), len(qmap)):
            e_rows.append(r)
    except RuntimeError:
        pass
    assert status == 200 and len(rows) == len(e_rows) >= 2
    for g, e in zip(rows, e_rows):
        assert list(g) == ["template", "score", "expected", "z", "probability", "frac-q", "frac-d", "species"]
        assert all(g[f] == e[f] for f in g if f != "probability") and g["probability"] == pytest.approx(e["probability"], rel=1e-9)
    assert wire.post_kmers(b"", tdb)[0] == 400


def test_cli_flags(tmp_path, capsys):
    rng = random.Random(12)
    data = random_fastq(rng, 200, plant=(b"ATGAC", 0.8), min_len=40)
    qmap, _ = ko.count_fastq(data)
    lists, attrs, summary = synthetic_db(list(qmap.keys()), rng, n_templates=9, decoys=20, share=0.6)
    fq = tmp_path / "reads.fastq"
    fq.write_bytes(data)
    dbp = tmp_path / "db.kjdb"
    TemplateDB.from_lists(lists, attrs, summary).save(str(dbp))
    assert cli.main(["-f", str(fq), "-P", "findKmers", "-o", "0"]) == 0
    assert capsys.readouterr().out.strip().endswith(f"Kmers:  {len(qmap)}")
    for score in ("winner", "standard"):
        assert cli.main(["-f", str(fq), "-p", "ATGAC", "-l", "16", "-s", "1", "-P", "findMatches", "-S", score,
                         "-d", str(dbp), "-o", "0"]) == 0
        out = capsys.readouterr().out.splitlines()
        assert out[0] == f"Kmers:  {len(qmap)}" and out[1].startswith("Template\tScore\tExpected\tz\tp_value")
        templates, hits = ko.first_match(OrderedDict(qmap), ko.TemplateDB(lists, attrs, summary))
        if score == "winner":
            exp = []
            try:
                for r in ko.find_matches(templates, summary, OrderedDict(qmap), len(qmap)):
                    exp.append(r)
            except RuntimeError:
                pass
        else:
            exp = [r for r in (ko.match_summary(len(qmap), templates, n, t, hits, summary) for n, t in templates.items()) if r]
            exp.sort(key=lambda r: -r["score"])
        assert [ln.split("\t")[0] for ln in out[2:]] == [r["template"] for r in exp]
        assert [int(ln.split("\t")[1]) for ln in out[2:]] == [r["score"] for r in exp]
