"""CPU-side checks: the C ABI library loads and exports every symbol the header declares (no GPU
compute is called), and the host logic around it (DB loaders, range planning, JSON helpers)."""
import json
import os
import random
import re

import numpy as np
import pytest

import kmer_oracle as ko
from conftest import ROOT, read_golden

import kmerjs_b200
from kmerjs_b200 import _abi
from kmerjs_b200 import dist as kdist
from kmerjs_b200.db import TemplateDB, load as load_db


def header_symbols():
    text = open(os.path.join(ROOT, "include", "kmerjs_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kj_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _abi.lib()                      # builds with nvcc when the sources are newer
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(L, s), f"{s} is declared in include/kmerjs_b200.h but not exported"
        assert s in _abi.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_abi.SIGNATURES) == syms
    assert L.kj_abi_version() == _abi.KJ_ABI_VERSION


def test_no_cpu_fallback_without_device():
    """Without an sm_100 device kj_init must fail loudly (KJ_E_NO_SM100), never fall back."""
    import ctypes as C
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    if os.environ.get("KMERJS_B200_EMU") == "1":
        pytest.skip("developer emulation harness")
    L = _abi.lib()
    h = C.c_void_p()
    rc = L.kj_init(0, None, C.byref(h))
    assert rc == _abi.KJ_E_NO_SM100 and not h.value
    assert b"no CPU fallback" in L.kj_last_error(None)
    with pytest.raises(_abi.KjError):
        kmerjs_b200.KmerJS("x.fastq").readFile().promise.result(timeout=60)


def test_host_only_exact_decimal_entry_points(known):
    """kj_stats_* are host-only (bignumber.js restatement); pinned by the reference's one stats
    known-answer and by the oracle."""
    ka = known["KA7_best_match"]
    summary = json.loads(read_golden("summary.json"))
    z = kmerjs_b200.zScore(ka["score"], ka["kmers-template"], ka["hits"], summary["uniqueLens"])
    assert round(float(z), 2) == ka["z"]
    ez = ko.z_score(ka["score"], ka["kmers-template"], ka["hits"], summary["uniqueLens"])
    from decimal import Decimal
    assert z == Decimal(ez.n).scaleb(-ez.e)
    assert float(kmerjs_b200.fastp(z)) * summary["templates"] == pytest.approx(ka["probability"], rel=1e-12)
    assert kmerjs_b200.etta == Decimal("0.00000001")
    for zt, p in (("10.7016", 1e-25), ("10.70160000000000000001", 1e-26), ("1.64485", 1.0), ("1.95997", 0.05),
                  ("-3", 1.0)):
        assert float(kmerjs_b200.fastp(zt)) == p          # strict '>' (lib/stats.js:56-112)


def test_json_helpers_and_complement(known):
    m = {"b": 2, "a": 1}
    assert list(kmerjs_b200.mapToJSON(m).items()) == [("b", 2), ("a", 1)]            # Map order kept
    assert kmerjs_b200.stringToMap('{"x": 1, "y": 2}') == {"x": 1, "y": 2}
    assert kmerjs_b200.objectToMap({"k": 3}) == kmerjs_b200.jsonToStrMap({"k": 3}) == {"k": 3}
    ka = known["KA1_complement"]
    assert kmerjs_b200.complement(ka["in"]) == ka["out"]
    assert kmerjs_b200.complement("ANxaT") == "AaxNT"
    assert kmerjs_b200.output_file_text({"ATGACGATTTTCATTG": 1171, "ATGACCGCACAGCAAG": 325}) == \
        known["KA9_out_head"]["head"][:47] + "}\n"
    kj = kmerjs_b200.KmerJS()
    assert (kj.preffix, kj.kmerLength, kj.step, kj.coverage, kj.progress, kj.env) == ("ATGAC", 16, 1, 1, True, "node")
    c = kmerjs_b200.KmerFinderClient("f.fastq", "node")
    assert c.maxHits == 100 and c.progress is True and c.dbName == "Kmers" and c.collection == "genomes"


def test_db_loaders_agree(tmp_path):
    attrs = {"T1": {"lengths": 100, "ulength": 3, "species": "one"},
             "T2": {"lengths": 200, "ulength": 2, "species": "two"}}
    lists = {"AAAA": ["T1", "T2"], "CCCC": ["T1"], "GGGN": ["T2"]}
    summary = {"templates": 2, "uniqueLens": 5, "totalLen": 300}
    a = TemplateDB.from_lists(lists, attrs, summary)
    kmer_docs = [{"kmer": k, "templates": [{"sequence": t, "lengths": attrs[t]["lengths"],
                                            "ulengths": attrs[t]["ulength"], "species": attrs[t]["species"]}
                                           for t in lst]} for k, lst in lists.items()]
    tmpl_docs = [{"sequence": t, "lengths": a_["lengths"], "ulenght": a_["ulength"], "species": a_["species"],
                  "reads": [k for k, lst in lists.items() if t in lst]} for t, a_ in attrs.items()]
    csv_map = {k: ",".join(lst) for k, lst in lists.items()}
    p1 = tmp_path / "kmerdocs.json"; p1.write_text(json.dumps(kmer_docs))
    p2 = tmp_path / "tmpldocs.json"; p2.write_text(json.dumps(tmpl_docs))
    p3 = tmp_path / "map.json"; p3.write_text(json.dumps(csv_map))
    ps = tmp_path / "summary.json"; ps.write_text(json.dumps(summary))
    p4 = tmp_path / "packed.npz"; a.save(str(p4))
    p5 = tmp_path / "packed.kjdb"; a.save(str(p5))              # the versioned binary (kj_db_save_packed): no pickle anywhere
    assert p5.read_bytes()[:8] == b"KJDBv001"
    import numpy as np
    assert not any(v.dtype == object for v in np.load(str(p4), allow_pickle=False).values())
    dbs = [load_db(str(p1), str(ps)), load_db(str(p2), summary), load_db(str(p4)), load_db(str(p5)),
           load_db(str(p3), summary, lengths={t: v["lengths"] for t, v in attrs.items()},
                   ulengths={t: v["ulength"] for t, v in attrs.items()},
                   descriptions={t: v["species"] for t, v in attrs.items()})]
    ref_lists, ref_attrs = a.to_lists()
    for d in dbs:
        l, at = d.to_lists()
        assert l == ref_lists and at == ref_attrs and d.summary == summary
    with pytest.raises(ValueError):
        load_db(str(p1))                    # per-k-mer documents need the Summary record


def test_plan_ranges_and_phase():
    data = read_golden("test_long.kmer.fastq")
    n = len(data)
    for world in (1, 2, 3, 8):
        ranges = kdist.plan_ranges(n, world, halo=64)
        assert ranges[0][0] == 0 and sum(r[1] for r in ranges) == n
        assert all(lo % 16 == 0 for lo, _, _ in ranges)
        assert all(rd >= own and lo + rd <= n for lo, own, rd in ranges)
        assert ranges[-1][1] == ranges[-1][2]
        cnt = [data[lo:lo + own].count(b"\n") for lo, own, _ in ranges]
        last = [data[lo:lo + own].rfind(b"\n") + 1 for lo, own, _ in ranges]
        bl, bc = kdist.phase_of_ranges(cnt, last, [own for _, own, _ in ranges])
        for (lo, own, _), l, c in zip(ranges, bl, bc):
            assert l == data[:lo].count(b"\n")
            assert c == lo - (data[:lo].rfind(b"\n") + 1)
    # a range without any newline passes the column on
    bl, bc = kdist.phase_of_ranges([1, 0, 2], [5, 0, 9], [10, 7, 20])
    assert bl == [0, 1, 1] and bc == [0, 5, 12]


def test_owner_function_matches_its_restatement():
    """kj_owner (host-side; the device uses the same two functions): full-length ACGT k-mers by their 2-bit
    key, byte-string k-mers (N, lower case, ...) by the padded bytes and the length.  Restated here so that a
    change of either hash is a visible change of the exchange format."""
    M = (1 << 64) - 1

    def mix64(x):
        x ^= x >> 30; x = (x * 0xBF58476D1CE4E5B9) & M
        x ^= x >> 27; x = (x * 0x94D049BB133111EB) & M
        x ^= x >> 31
        return x

    def owner(kmer: bytes, n_parts: int) -> int:
        if all(b in b"ACGT" for b in kmer) and 1 <= len(kmer) <= 32:
            key = 0
            for b in kmer:
                key = (key << 2) | ((b >> 1) & 3)
            return (mix64(key ^ 0x9E3779B97F4A7C15) >> 32) % n_parts
        pad = kmer + bytes(32 - len(kmer))
        h = 0x243F6A8885A308D3 ^ len(kmer)
        for i in range(4):
            h = mix64(h ^ int.from_bytes(pad[8 * i:8 * i + 8], "little"))
        return (h >> 32) % n_parts

    L = _abi.lib()
    rng = random.Random(77)
    seen = set()
    for _ in range(400):
        n = rng.choice([1, 5, 16, 31, 32])
        alphabet = rng.choice([b"ACGT", b"ACGTN", b"ACGTacgtN-"])
        kmer = bytes(rng.choice(alphabet) for _ in range(n))
        for parts in (1, 2, 3, 8):
            got = L.kj_owner(kmer, len(kmer), parts)
            assert got == owner(kmer, parts), (kmer, parts)
            assert 0 <= got < parts
        seen.add(L.kj_owner(kmer, len(kmer), 8))
    assert seen == set(range(8))                       # both kinds of k-mer spread over the parts


def test_wire_host_side_shapes():
    """The parts of the findFirstMatch exchange that need no device (lib/kmerFinderClient.js:132-161)."""
    import json
    from kmerjs_b200 import wire
    from kmerjs_b200.matching import NoHitsError
    body = wire.first_match_request({"ATGACAAAAAAAAAAA": 3, "ATGACCCCCCCCCCCC": 1}, "Kmers", "genomes")
    doc = json.loads(body)
    assert list(doc) == ["ATGACAAAAAAAAAAA", "ATGACCCCCCCCCCCC", "db", "collection"]
    assert wire._query_of(body) == {"ATGACAAAAAAAAAAA": 3, "ATGACCCCCCCCCCCC": 1}
    with pytest.raises(NoHitsError, match="No hits were found!"):
        wire.parse_first_match_reply(204, b"")
    with pytest.raises(RuntimeError):
        wire.parse_first_match_reply(500, b"")
    reply = json.dumps({"templates": {"T1": {"tScore": 4, "uScore": 2, "kmers": ["a", "b", "a"]}}, "hits": 2, "summary": {}})
    w = wire.parse_first_match_reply(200, reply.encode())
    assert list(w["templates"]["T1"]["kmers"]) == ["a", "b"] and w["hits"] == 2
    assert wire.post_kmers(b"", None)[0] == 400


def test_js_number_prints_like_a_template_string():
    """Number#toString of JavaScript, which lib/kmerFinderClient.js:195-208 uses for the TSV rows: shortest digits that
    round-trip, positional notation for 1e-7 <= |x| < 1e21, exponent form outside."""
    from kmerjs_b200.kmer_finder_client import js_number
    cases = [(5.0, "5"), (5, "5"), (0.5, "0.5"), (3e-05, "0.00003"), (1e-6, "0.000001"), (1e-7, "1e-7"), (1.5e-7, "1.5e-7"),
             (123456789.125, "123456789.125"), (1e21, "1e+21"), (1e20, "100000000000000000000"), (-2.5, "-2.5"),
             (0.0, "0"), (100.0, "100"), (1234.5678, "1234.5678"), (2.220446049250313e-16, "2.220446049250313e-16"),
             (float("nan"), "NaN"), (float("inf"), "Infinity"), (float("-inf"), "-Infinity"), (99.99, "99.99"),
             (0.1 + 0.2, "0.30000000000000004"), (1e-10, "1e-10"), (12e22, "1.2e+23")]
    for x, want in cases:
        assert js_number(x) == want, (x, js_number(x), want)
    assert js_number("NC_017625") == "NC_017625" and js_number(True) == "True"
