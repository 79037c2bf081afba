"""Host-side mirror of lib/kmerFinderClient.js: class KmerFinderClient with findKmers /
findFirstMatch / findMatches (winner-takes-all), same constructor positions and error texts.

What changes underneath: ``findFirstMatch`` no longer POSTs the whole k-mer map to a server that
asks Redis (lib/kmerFinderClient.js:128-173, lib/kmerFinderServer.js:171-226); it probes the
template DB resident in GPU memory (kj_first_match).  ``findMatches`` drives kj_wta_next, one
generator step per winner (lib/kmerFinderClient.js:174-290).  ``db`` may be a
:class:`kmerjs_b200.db.TemplateDB`, or a path understood by :func:`kmerjs_b200.db.load`."""
from __future__ import annotations

import numpy as np

from . import _abi
from .counts import Counts
from .db import TemplateDB, load as load_db
from .kmers import KmerJS, KmerMap, Promise, mapToJSON  # noqa: F401  (mapToJSON re-exported like the JS module)
from .matching import Match, NoHitsError

def js_number(x) -> str:
    """Number#toString of JavaScript: what a template string prints (lib/kmerFinderClient.js:195-208).  Shortest digits
    that round-trip; positional notation for 1e-7 <= |x| < 1e21, exponent form outside (3e-05 prints as 0.00003, 5.0 as 5)."""
    from decimal import Decimal
    if isinstance(x, bool) or not isinstance(x, (int, float)):
        return str(x)
    if isinstance(x, int):
        return str(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "0"
    sign, digits, exp = Decimal(repr(float(x))).as_tuple()
    ds = "".join(map(str, digits)).lstrip("0") or "0"
    trail = len(ds) - len(ds.rstrip("0"))
    ds, exp = ds.rstrip("0") or "0", exp + trail
    k, n = len(ds), len(ds) + exp                     # value = 0.ds x 10^n
    if k <= n <= 21:
        body = ds + "0" * (n - k)
    elif 0 < n <= 21:
        body = ds[:n] + "." + ds[n:]
    elif -6 < n <= 0:
        body = "0." + "0" * (-n) + ds
    else:
        e = n - 1
        body = (ds[0] + ("." + ds[1:] if k > 1 else "")) + "e" + ("+" if e >= 0 else "-") + str(abs(e))
    return ("-" if sign else "") + body


TSV_HEADER = ("Template\tScore\tExpected\tz\tp_value\tquery\tcoverage [%]\ttemplate coverage [%]"
              "\tdepth\tKmers in Template\tDescription\n")      # lib/kmerFinderClient.js:185


def counts_from_map(kmerMap: dict, preffix: str, length: int, step: int) -> Counts:
    """Device table for a k-mer map that did not come from findKmers (e.g. parsed from JSON):
    keys in Map order get ordinals 0..n-1."""
    c = Counts(preffix.encode("latin-1"), length, step)
    regular, irregular = [], []
    code = np.full(256, 255, dtype=np.uint8)
    for ch, v in (("A", 0), ("C", 1), ("T", 2), ("G", 3)):
        code[ord(ch)] = v
    items = [(k, v) for k, v in kmerMap.items() if isinstance(v, (int, float)) and not isinstance(v, bool)]
    for i, (k, v) in enumerate(items):
        b = k.encode("latin-1")
        cs = code[np.frombuffer(b, dtype=np.uint8)] if b else np.zeros(0, np.uint8)
        key = 0
        if len(b) == length and len(b) <= 32 and (cs != 255).all():
            for x in cs.tolist():
                key = (key << 2) | x
        if len(b) == length and (cs != 255).all() and key != 0xFFFFFFFFFFFFFFFF:
            regular.append((key, int(v), i))
        else:
            if len(b) > 32:
                raise ValueError("k-mers longer than 32 bytes are not supported")
            irregular.append((b, int(v), i))
    if regular:
        c.merge_host_records(np.array(regular, dtype=np.uint64).reshape(-1, 3))
    if irregular:
        raw = np.zeros((len(irregular), 56), dtype=np.uint8)
        for j, (b, v, i) in enumerate(irregular):
            raw[j, :len(b)] = np.frombuffer(b, dtype=np.uint8)
            raw[j, 32:56] = np.array([len(b), v, i], dtype=np.uint64).view(np.uint8)
        c.merge_irregular(raw.reshape(-1))
    c.finish()
    return c


class KmerFinderClient(KmerJS):
    def __init__(self, fastq, env="node", preffix="ATGAC", length=16, step=1, coverage=1, out=True,
                 db="server", url="http://localhost:3000/kmers", summary=None, collection="genomes",
                 dbName="Kmers"):
        super().__init__(fastq, preffix, length, step, coverage, out, env)   # :118
        self.dbLocation = db
        self.dbURL = url
        self.collection = collection
        self.dbName = dbName
        self.maxHits = 100                                                    # :123
        self._summary_arg = summary
        self._db = None
        self._match = None
        self.firstMatches = None
        self.summary = None

    def _template_db(self) -> TemplateDB:
        if self._db is None:
            if isinstance(self.dbLocation, TemplateDB):
                self._db = self.dbLocation
            else:
                self._db = load_db(self.dbLocation, self._summary_arg)
        return self._db

    def findKmers(self):                                                      # :125-127
        return self.readFile()

    def findFirstMatch(self, kmerQuery):
        """Promise of {templates, summary, hits} (lib/kmerFinderClient.js:128-173 reply contract;
        producer semantics lib/kmerFinderServer.js:171-226).  Rejects with 'No hits were found!'."""
        promise = Promise()
        try:
            db = self._template_db()
            # the reference adds its two bookkeeping keys to the caller's map (:132-133)
            kmerQuery["db"] = self.dbName
            kmerQuery["collection"] = self.collection
            counts = getattr(kmerQuery, "counts", None)
            if counts is None or counts.handle is None:
                counts = counts_from_map(kmerQuery, self.preffix, int(self.kmerLength), int(self.step))
                if self.kmerMapSize == 0:
                    self.kmerMapSize = counts.size
            self._counts = counts
            m = Match(counts, db)
            if self.kmerMapSize:
                m.set_query_size(int(self.kmerMapSize))
            self._match = m
            promise.set_result({"templates": m.templates(), "summary": dict(db.summary), "hits": m.hits})
        except NoHitsError:
            promise.set_exception(NoHitsError("No hits were found!"))          # :159-161
        except BaseException as exc:  # noqa: BLE001
            promise.set_exception(exc)
        return promise

    def findMatches(self, winner, kmerMap):
        """Generator of row dicts, one per winner (lib/kmerFinderClient.js:174-290).  ``kmerMap`` is
        mutated like the reference's Map: the winner's k-mers are deleted after every row."""
        self.summary = winner["summary"]                                       # :287
        self.firstMatches = winner["templates"]                               # :288
        m = self._match
        if m is None:
            raise RuntimeError("findMatches needs the result of findFirstMatch")
        m.set_max_hits(self.maxHits)
        db = self._template_db()
        keys = [k for k in kmerMap.keys() if k not in ("db", "collection")]

        def loop():
            first = True
            alive_prev = None
            while True:
                row = m.next_row()             # raises NoHitsError with the reference's texts
                if row is None:
                    return
                if first:
                    first = False
                    self.firstMatches = m.templates()                        # :182-184
                    if self.progress:
                        print(TSV_HEADER, end="")
                if self.progress:                                              # :195-208
                    print("\t".join(js_number(row[k]) for k in ("template", "score", "expected", "z", "probability",
                                                          "frac-q", "frac-d", "depth", "kmers-template",
                                                          "species")))
                # removeWinnerKmers on the caller's map (:220-230)
                alive = self._counts.alive()
                if len(alive) == len(keys):
                    gone = np.nonzero((alive == 0) & ((alive_prev if alive_prev is not None else 1) != 0))[0]
                    for i in gone.tolist():
                        kmerMap.pop(keys[i], None)
                    alive_prev = alive
                yield row

        return loop()

    def close(self):
        if self._match is not None:
            self._match.free()
            self._match = None
