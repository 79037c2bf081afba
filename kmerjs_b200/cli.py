"""Command line of the reference (lib/cli.js:9-20), same flags:

    python -m kmerjs_b200.cli -f reads.fastq -p ATGAC -l 16 -s 1 -P findMatches -S winner -d db.json [-u summary.json]

-d names the template database: where the reference takes a Mongo collection name (and -u a mongodb:// URL), this takes a
database file in any layout kj_db_load reads (and -u the Summary JSON, when the file does not carry one)."""
from __future__ import annotations

import argparse
import sys


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="kmerjs", description=__doc__)
    ap.add_argument("-f", "--fastq", default="test_data/test_long.fastq", help="FASTQ file to parse")
    ap.add_argument("-p", "--preffix", default="ATGAC", help="Kmer preffix")
    ap.add_argument("-l", "--length", type=int, default=16, help="Kmer lenght")
    ap.add_argument("-s", "--step", type=int, default=1, help="Kmer step")
    ap.add_argument("-c", "--coverage", type=int, default=1, help="Min coverage")
    ap.add_argument("-o", "--output", type=int, default=1, help="Print info")
    ap.add_argument("-P", "--program", default="findMatches", choices=["findKmers", "findMatches"], help="Program to execute")
    ap.add_argument("-S", "--score", default="winner", choices=["standard", "winner"], help="Score to execute")
    ap.add_argument("-d", "--database", default="KmerMap", help="Database to query: a template DB file")
    ap.add_argument("-u", "--url", default=None, help="Summary JSON of the database (the reference: its URL)")
    o = ap.parse_args(argv)

    from .db import load_native
    from .kmer_finder_client import TSV_HEADER, KmerFinderClient, js_number
    from .matching import Match, NoHitsError

    client = KmerFinderClient(o.fastq, "node", o.preffix, o.length, o.step, o.coverage, bool(o.output), o.database)
    kmers = client.findKmers().promise.result()
    print("Kmers: ", len(kmers))
    if o.program == "findKmers":
        return 0
    db = load_native(o.database, o.url)
    client.dbLocation = client._db = db
    cols = ("template", "score", "expected", "z", "probability", "frac-q", "frac-d", "depth", "kmers-template", "species")
    try:
        if o.score == "standard":
            m = Match(kmers.counts, db)
            rows = m.standard_scoring()
            m.free()
        else:
            client.progress = False
            rows = []
            try:
                for r in client.findMatches(client.findFirstMatch(kmers).result(), kmers):
                    rows.append(r)
            except NoHitsError:
                if not rows:                    # with rows: the generator's way of saying the query is used up
                    raise
    except NoHitsError as exc:
        print(str(exc), file=sys.stderr)
        return 1
    sys.stdout.write(TSV_HEADER)
    for r in rows:
        print("\t".join(js_number(r[k]) for k in cols))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
