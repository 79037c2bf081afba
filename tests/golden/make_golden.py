"""Regenerates tests/golden/ from the reference checkout (run in the build container, where
/root/reference exists; the GPU box only sees the committed outputs).

  * copies the reference's own fixtures for this path (test_data/*: inputs and golden maps);
  * writes known_answers.json: the values the reference's tests pin, each with
    the file:line it comes from;
  * writes oracle_digests.json: sha256 digests of the oracle's output on the FASTQ fixtures for a
    grid of (prefix, k, step) -- regression pins of the restatement, NOT reference-pinned.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("KMERJS_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle"))

COPY = ["test_short.fastq", "test_long.kmer.fastq", "test_kmers.fastq", "kmers_long.json", "test_long.json",
        "db_long_results.json", "db_short_results.json", "summary.json"]

KNOWN = {
    "KA1_complement": {"in": "ATGACCTGAGAGCCTT", "out": "AAGGCTCTCAGGTCAT", "src": "test/kmers.js:21-26"},
    "KA2_first_key": {"line": None, "first": "ATGACGCAATACTCCT", "src": "test/kmers.js:12-19"},
    "KA3_test_short": {"map": [["ATGACGCAATACTCCT", 1], ["ATGACCTGAGAGCCTT", 1]],
                       "src": "test/kmers.js:28-35, test/kmerFinderClient.js:16-17"},
    "KA4_test_long_kmer_size": {"size": 401, "src": "test/kmers.js:45-52"},
    "KA6_kmers_long": {"size": 6191, "sum": 9301, "n_keys_with_N": 9, "src": "test/kmers.js:37-44 + test_data/kmers_long.json"},
    "KA7_best_match": {"template": "NC_017625", "score": 2295, "expected": 108, "z": 211.00,
                       "probability": 5.03e-23, "frac-q": 74.14, "frac-d": 47.02, "depth": 0.36,
                       "total-frac-q": 74.14, "total-frac-d": 47.02, "total-temp-cover": 0.36,
                       "kmers-template": 4881, "species": "Escherichia coli DH1",
                       "tScore": 3596, "hits": 179108, "kmerMapSize": 6191,
                       "src": "test/kmerFinderServer.js:70-82 + test_data/db_long_results.json + test_data/summary.json"},
    "KA8_hits": {"db_long": 179108, "db_short": 158, "src": "test_data/db_*_results.json: sum(templateentries) == hits"},
    "KA9_out_head": {"head": None, "src": "lib/index.js:381-388 + test_data/out.json"},
}


def digest(counts):
    h = hashlib.sha256()
    for k in sorted(counts):
        h.update(k + b"\t" + str(counts[k]).encode() + b"\n")
    return h.hexdigest()[:16]


def main():
    import kmer_oracle as ko
    for f in COPY:
        shutil.copyfile(os.path.join(REF, "test_data", f), os.path.join(HERE, f))
        os.chmod(os.path.join(HERE, f), 0o644)
    # the template literal of test/kmers.js:14-15 (it contains a newline and the indentation)
    src = open(os.path.join(REF, "test", "kmers.js")).read()
    a = src.index("`") + 1
    KNOWN["KA2_first_key"]["line"] = src[a:src.index("`", a)]
    KNOWN["KA9_out_head"]["head"] = open(os.path.join(REF, "test_data", "out.json")).read(120)
    json.dump(KNOWN, open(os.path.join(HERE, "known_answers.json"), "w"), indent=1)
    dig = {}
    for f in ["test_short.fastq", "test_long.kmer.fastq", "test_kmers.fastq"]:
        data = open(os.path.join(HERE, f), "rb").read()
        for prefix, k, step in [(b"ATGAC", 16, 1), (b"", 31, 1), (b"ATGAC", 16, 3), (b"A", 5, 2), (b"GT", 32, 1),
                                (b"", 1, 1), (b"ATGACG", 6, 1)]:
            counts, lines = ko.count_fastq(data, prefix, k, step)
            dig[f"{f}|{prefix.decode()}|{k}|{step}"] = {"unique": len(counts), "total": sum(counts.values()),
                                                        "lines": lines, "sha": digest(counts)}
    json.dump(dig, open(os.path.join(HERE, "oracle_digests.json"), "w"), indent=1)
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
