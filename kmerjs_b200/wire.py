"""Wire formats around the path, so that existing front ends can talk to the GPU service unchanged.

  first_match_request / first_match_reply    the exchange of KmerFinderClient#findFirstMatch with the reduced-DB service:
      request body  = JSON.stringify(mapToJSON(kmerQuery)) with the two bookkeeping keys ``db`` and ``collection``
                      (lib/kmerFinderClient.js:132-136)
      reply         = {"templates": {name: {tScore, uScore, lengths, ulength, species, kmers: [..]}}, "hits", "summary"}
                      with status 200 (lib/kmerFinderClient.js:139-157; producer lib/kmerFinderServer.js:171-226), or
                      status 204 when nothing hit (:158-161)
  post_kmers                                   Express ``POST /kmers`` (server/app.js:22-54): body = a k-mer map as JSON,
      reply = JSON array of rows {template, score, expected, z, probability, frac-q, frac-d, species}; the two fields the
      handler also asks for (``coverage``, ``ulength``) are undefined in the reference's row Map and vanish in res.json.

No HTTP server lives here (out of scope: SURVEY.md 2); these are the pure functions a handler calls."""
from __future__ import annotations

import json

from .kmer_finder_client import counts_from_map
from .matching import Match, NoHitsError


def first_match_request(kmerMap: dict, dbName: str = "Kmers", collection: str = "genomes") -> bytes:
    body = dict(kmerMap)                       # Map order = insertion order
    body["db"] = dbName
    body["collection"] = collection
    return json.dumps(body, separators=(",", ":")).encode()


def _query_of(body) -> dict:
    doc = json.loads(body) if isinstance(body, (bytes, bytearray, str)) else dict(body)
    return {k: v for k, v in doc.items() if k not in ("db", "collection") and isinstance(v, (int, float)) and not isinstance(v, bool)}


def first_match_reply(body, db, preffix: str = "ATGAC", length: int = 16, step: int = 1):
    """(status, reply bytes): the service side of findFirstMatch on the GPU-resident DB."""
    q = _query_of(body)
    counts = counts_from_map(q, preffix, length, step)
    try:
        m = Match(counts, db)
    except NoHitsError:
        counts.free()
        return 204, b""
    try:
        templates = m.templates(with_kmers=True, keys=list(q.keys()))
        reply = {"templates": templates, "hits": m.hits, "summary": dict(db.summary)}
        return 200, json.dumps(reply, separators=(",", ":")).encode()
    finally:
        m.free()
        counts.free()


def parse_first_match_reply(status: int, body: bytes) -> dict:
    """The client side (lib/kmerFinderClient.js:139-161): ``kmers`` arrays become sets; 204 rejects."""
    if status == 204:
        raise NoHitsError("No hits were found!")
    if status not in (200, 201, 202):
        raise RuntimeError("error")
    winner = json.loads(body)
    for hit in winner["templates"].values():
        hit["kmers"] = dict.fromkeys(hit["kmers"])          # a Set in insertion order
    return winner


def post_kmers(body, db, preffix: str = "ATGAC", length: int = 16, step: int = 1, max_hits: int = 100):
    """(status, reply bytes) of ``POST /kmers``: winner-takes-all rows of the posted k-mer map."""
    if not body:
        return 400, b""
    q = _query_of(body)
    counts = counts_from_map(q, preffix, length, step)
    rows = []
    try:
        m = Match(counts, db)
        try:
            rows, _err = m.all_rows(max_hits)
        finally:
            m.free()
    except NoHitsError:
        rows = []
    finally:
        counts.free()
    out = [{k: r[k] for k in ("template", "score", "expected", "z", "probability", "frac-q", "frac-d", "species")} for r in rows]
    return 200, json.dumps(out, separators=(",", ":")).encode()
