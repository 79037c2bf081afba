"""Object wrapper over kj_first_match / kj_wta_next / kj_standard_scoring."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .counts import Counts
from .db import TemplateDB

# lib/kmerFinderClient.js:75-89 -- key order is part of the output JSON layout
ROW_KEYS = ["template", "score", "expected", "z", "probability", "frac-q", "frac-d", "depth",
            "kmers-template", "total-frac-q", "total-frac-d", "total-temp-cover", "species"]


class NoHitsError(RuntimeError):
    """'No hits were found!' family (lib/kmerFinderClient.js:161,265,284)."""


def matched_segment_bytes(cap_entries: int, cap_pairs: int) -> int:
    return int(_abi.lib().kj_matched_segment_bytes(cap_entries, cap_pairs))


def row_to_dict(r: _abi.kj_row, db: TemplateDB) -> dict:
    def num(x):                      # JS numbers: 108 prints as 108, not 108.0
        return int(x) if float(x).is_integer() and abs(x) < 2 ** 53 else float(x)
    t = int(r.template_id)
    return {"template": db.names[t], "score": int(r.score), "expected": num(r.expected), "z": num(r.z),
            "probability": float(r.probability), "frac-q": num(r.frac_q), "frac-d": num(r.frac_d),
            "depth": num(r.depth), "kmers-template": int(r.kmers_template),
            "total-frac-q": num(r.total_frac_q), "total-frac-d": num(r.total_frac_d),
            "total-temp-cover": num(r.total_temp_cover), "species": db.species[t]}


class Match:
    def __init__(self, counts: Counts, db: TemplateDB, *, local_only: bool = False, part: int = 0,
                 n_parts: int = 1):
        self.counts, self.db = counts, db
        self.ctx = counts.ctx
        self._L = _abi.lib()
        self._dbh = db.device(self.ctx, part, n_parts)
        h = C.c_void_p()
        fn = self._L.kj_first_match_local if local_only else self._L.kj_first_match
        rc = fn(self.ctx.handle, counts.handle, self._dbh.handle, C.byref(h))
        if rc == _abi.KJ_E_NO_HITS:
            raise NoHitsError("No hits were found!")
        _abi.check(rc, self.ctx.handle)
        self.handle = h
        self.last_row = None

    @classmethod
    def from_matched(cls, ctx, db: TemplateDB, segments, query_size: int, *, part: int = 0, n_parts: int = 1):
        """A match over the matched entries of all ranks (kj_match_from_matched).  segments: one
        (entries_dev_ptr, n_entries, tmpl_dev_ptr, n_pairs) per rank, as written by export_matched."""
        self = cls.__new__(cls)
        self.counts, self.db, self.ctx = None, db, ctx
        self._L = _abi.lib()
        self._dbh = db.device(ctx, part, n_parts)
        n = len(segments)
        arr = lambda i: (C.c_uint64 * max(n, 1))(*[int(s[i]) for s in segments])
        h = C.c_void_p()
        _abi.check(self._L.kj_match_from_matched(ctx.handle, self._dbh.handle, n, arr(0), arr(1), arr(2), arr(3),
                                                 int(query_size), C.byref(h)), ctx.handle)
        self.handle = h
        self.last_row = None
        return self

    @classmethod
    def from_segments(cls, ctx, db: TemplateDB, n_segments: int, dev_ptr: int, cap_entries: int, cap_pairs: int, *,
                      part: int = 0, n_parts: int = 1):
        """A match over the fixed-capacity segments of all ranks (kj_match_from_segments): the gathered buffer at
        dev_ptr must outlive the match.  Sizes and flags are checked by commit()."""
        self = cls.__new__(cls)
        self.counts, self.db, self.ctx = None, db, ctx
        self._L = _abi.lib()
        self._dbh = db.device(ctx, part, n_parts)
        h = C.c_void_p()
        _abi.check(self._L.kj_match_from_segments(ctx.handle, self._dbh.handle, n_segments, C.c_void_p(dev_ptr),
                                                  cap_entries, cap_pairs, C.byref(h)), ctx.handle)
        self.handle = h
        self.last_row = None
        return self

    def export_segment(self, dev_ptr: int, cap_entries: int, cap_pairs: int, query_size: int, flags: int = 0):
        """This rank's matched entries as one fixed-capacity segment (stream-ordered, kj_match_export_segment)."""
        _abi.check(self._L.kj_match_export_segment(self.handle, C.c_void_p(dev_ptr), cap_entries, cap_pairs,
                                                   int(query_size), int(flags)), self.ctx.handle)

    @property
    def query_size(self) -> int:
        return int(self._L.kj_match_query_size(self.handle))

    def segment_sizes(self):
        ne, np_ = C.c_uint64(), C.c_uint64()
        _abi.check(self._L.kj_match_segment_sizes(self.handle, C.byref(ne), C.byref(np_)), self.ctx.handle)
        return int(ne.value), int(np_.value)

    # -- distributed protocol ---------------------------------------------------------------------
    def matched_size(self):
        """(entries that hit the DB, sum of their template-list lengths) on this rank."""
        ne, np_ = C.c_uint64(), C.c_uint64()
        _abi.check(self._L.kj_match_matched_size(self.handle, C.byref(ne), C.byref(np_)), self.ctx.handle)
        return int(ne.value), int(np_.value)

    def export_matched(self, entries_ptr: int, cap_entries: int, tmpl_ptr: int, cap_pairs: int):
        """Stream-ordered on the context's stream (kj_match_export_matched)."""
        _abi.check(self._L.kj_match_export_matched(self.handle, C.c_void_p(entries_ptr), cap_entries,
                                                   C.c_void_p(tmpl_ptr), cap_pairs), self.ctx.handle)

    def vec_len(self, which: int) -> int:
        return int(self._L.kj_match_vec_len(self.handle, which))

    def get(self, which: int, dev_ptr: int, sync: bool = True):
        fn = self._L.kj_match_get if sync else self._L.kj_match_get_async
        _abi.check(fn(self.handle, which, C.c_void_p(dev_ptr)), self.ctx.handle)

    def set(self, which: int, dev_ptr: int, sync: bool = True):
        fn = self._L.kj_match_set if sync else self._L.kj_match_set_async
        _abi.check(fn(self.handle, which, C.c_void_p(dev_ptr)), self.ctx.handle)

    def commit(self):
        _abi.check(self._L.kj_match_commit(self.handle), self.ctx.handle)

    def set_query_size(self, n: int):
        _abi.check(self._L.kj_match_set_query_size(self.handle, n), self.ctx.handle)

    # -- results ----------------------------------------------------------------------------------
    @property
    def hits(self) -> int:
        return int(self._L.kj_match_hits(self.handle))

    def scores(self):
        """(uScore[T], tScore[T], order[n_matched]) of the first match."""
        T = self.db.n_templates
        u = np.zeros(max(T, 1), dtype=np.uint64)
        t = np.zeros(max(T, 1), dtype=np.uint64)
        n = int(self._L.kj_match_n_matched(self.handle))
        o = np.zeros(max(n, 1), dtype=np.uint32)
        _abi.check(self._L.kj_match_scores(self.handle, u.ctypes.data, t.ctypes.data, o.ctypes.data),
                   self.ctx.handle)
        return u[:T], t[:T], o[:n]

    def template_kmers(self, template_id: int) -> np.ndarray:
        """Export-order positions of the query k-mers that list the template (its ``kmers`` Set)."""
        n = C.c_uint64()
        _abi.check(self._L.kj_match_template_kmers(self.handle, template_id, None, 0, C.byref(n)), self.ctx.handle)
        idx = np.zeros(max(int(n.value), 1), dtype=np.uint64)
        _abi.check(self._L.kj_match_template_kmers(self.handle, template_id, idx.ctypes.data, idx.size, C.byref(n)),
                   self.ctx.handle)
        return idx[:int(n.value)]

    def templates(self, with_kmers: bool = False, keys=None) -> dict:
        """name -> {tScore,uScore,lengths,ulength,species[,kmers]} in first-encounter order: the
        ``templates`` Map of the findFirstMatch reply (lib/kmerFinderClient.js:150-157).  ``kmers`` (the
        Set of matched k-mers, in insertion order) is materialised on request: ``keys`` is the list of
        query k-mers in Map order."""
        u, t, order = self.scores()
        out = {}
        for i in order.tolist():
            out[self.db.names[i]] = {"tScore": int(t[i]), "uScore": int(u[i]),
                                     "lengths": int(self.db.lengths[i]), "ulength": int(self.db.ulengths[i]),
                                     "species": self.db.species[i]}
            if with_kmers:
                idx = self.template_kmers(i).tolist()
                out[self.db.names[i]]["kmers"] = [keys[j] for j in idx] if keys is not None else idx
        return out

    def set_max_hits(self, n: int):
        _abi.check(self._L.kj_match_set_max_hits(self.handle, int(n)), self.ctx.handle)

    def next_row(self):
        """One step of the findMatches generator: row dict, or None when the loop ended."""
        r = _abi.kj_row()
        rc = self._L.kj_wta_next(self.handle, C.byref(r))
        if rc == _abi.KJ_E_NO_HITS:
            raise NoHitsError("No hits were found! (nHits === 0)")
        if rc == _abi.KJ_E_NO_WINNER:
            raise NoHitsError("No hits were found! (kmerResults.length === 0)")
        _abi.check(rc, self.ctx.handle)
        self.last_row = r
        return row_to_dict(r, self.db) if rc == 1 else None

    def all_rows(self, max_hits: int = 100):
        """The whole findMatches generator in one call (kj_wta_all): (rows, error) where error is the
        NoHitsError the generator would have raised after yielding the rows, or None."""
        self.set_max_hits(max_hits)
        buf = (_abi.kj_row * max(max_hits, 1))()
        n, end = C.c_uint32(), C.c_int()
        _abi.check(self._L.kj_wta_all(self.handle, buf, max(max_hits, 1), C.byref(n), C.byref(end)), self.ctx.handle)
        rows = [row_to_dict(buf[i], self.db) for i in range(n.value)]
        err = None
        if end.value == _abi.KJ_E_NO_HITS:
            err = NoHitsError("No hits were found! (nHits === 0)")
        elif end.value == _abi.KJ_E_NO_WINNER:
            err = NoHitsError("No hits were found! (kmerResults.length === 0)")
        return rows, err

    def rows(self, max_hits: int = 100):
        """findMatches as a generator over all_rows(): the rows, then the error the reference throws (if any)."""
        rows, err = self.all_rows(max_hits)
        yield from rows
        if err is not None:
            raise err

    def defer_rows(self, on: bool = True):
        _abi.check(self._L.kj_match_defer_rows(self.handle, 1 if on else 0), self.ctx.handle)

    def next_row_begin(self):
        """(state, row): state 0 = loop ended, 1 = row complete, 2 = winner accepted, row pending (finish_row)."""
        r = _abi.kj_row()
        rc = self._L.kj_wta_next(self.handle, C.byref(r))
        if rc == _abi.KJ_E_NO_HITS:
            raise NoHitsError("No hits were found! (nHits === 0)")
        if rc == _abi.KJ_E_NO_WINNER:
            raise NoHitsError("No hits were found! (kmerResults.length === 0)")
        _abi.check(rc, self.ctx.handle)
        self.last_row = r
        return rc, (row_to_dict(r, self.db) if rc == 1 else None)

    def finish_row(self):
        r = _abi.kj_row()
        _abi.check(self._L.kj_wta_row(self.handle, C.byref(r)), self.ctx.handle)
        self.last_row = r
        return row_to_dict(r, self.db)

    def standard_scoring(self) -> list:
        n = C.c_uint32()
        cap = max(self.db.n_templates, 1)
        rows = (_abi.kj_row * cap)()
        _abi.check(self._L.kj_standard_scoring(self.handle, rows, cap, C.byref(n)), self.ctx.handle)
        return [row_to_dict(rows[i], self.db) for i in range(n.value)]

    def free(self):
        if getattr(self, "handle", None):
            self._L.kj_match_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
