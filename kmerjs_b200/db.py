"""Template database: the k-mer -> ordered template list store the reference keeps in Redis/Mongo
(lib/kmerFinderServer.js:68-92,171-226,712-728), loaded into GPU memory as a hash index plus CSR
template lists (kj_db_create).

On-disk formats accepted by :func:`load` (all defined by the reference):
  * per-k-mer documents  [{"kmer": K, "templates": [{"sequence","lengths","ulengths","species"}, ..]}, ..]
    (lib/kmerFinderServer.js:68-92; the Redis lists of :184-199 hold the same records)
  * per-template documents [{"sequence","lengths","ulenght","species","reads":[K, ..]}, ..]
    (src/kmerPyToMongo.py:35-42 -- the field really is spelled ``ulenght``)
  * the original KmerFinder map {K: "T1,T2,.."} with side tables for lengths / ulengths /
    descriptions (lib/index.js:184-192, src/kmerPyToMongo.py:15-24)
  * ``.kjdb`` written by :meth:`TemplateDB.save` / kj_db_save_packed: a versioned binary of plain arrays (the fast path;
    nothing in it is executable -- no pickle); ``.npz`` of the same arrays is still read and written (allow_pickle=False)
:func:`load` parses on the host (numpy arrays, e.g. to hand the DB to the CPU oracle in tests); :func:`load_native` goes
through the C ABI (kj_db_load): file -> GPU without Python touching the records.
The Summary record {"templates","uniqueLens","totalLen"} (lib/kmerFinderServer.js:716-724,
test_data/summary.json) travels with the DB."""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _abi
from .context import Context, default_context


class TemplateDB:
    """Host description (numpy arrays) + lazily created device handle."""

    def __init__(self, kmer_bytes: np.ndarray, kmer_len: np.ndarray, list_off: np.ndarray,
                 tmpl_ids: np.ndarray, names: list, lengths: np.ndarray, ulengths: np.ndarray,
                 species: list, summary: dict):
        self.kmer_bytes = np.ascontiguousarray(kmer_bytes, dtype=np.uint8)
        self.kmer_len = np.ascontiguousarray(kmer_len, dtype=np.uint32)
        self.list_off = np.ascontiguousarray(list_off, dtype=np.uint64)
        self.tmpl_ids = np.ascontiguousarray(tmpl_ids, dtype=np.uint32)
        self.names = list(names)
        self.lengths = np.ascontiguousarray(lengths, dtype=np.uint64)
        self.ulengths = np.ascontiguousarray(ulengths, dtype=np.uint64)
        self.species = list(species)
        self.summary = {"templates": int(summary["templates"]), "uniqueLens": int(summary["uniqueLens"]),
                        "totalLen": int(summary.get("totalLen", 0))}
        self._dev = {}

    # ------------------------------------------------------------------ constructors
    @classmethod
    def from_lists(cls, kmer_lists: dict, attrs: dict, summary: dict):
        """kmer_lists: {kmer(str|bytes): [template name, ..]} in DB order;
        attrs: {name: {"lengths", "ulength", "species"}} (insertion order = template ids)."""
        names = list(attrs.keys())
        tid = {n: i for i, n in enumerate(names)}
        kb, kl, off, tm = [], [], [0], []
        for kmer, lst in kmer_lists.items():
            b = kmer if isinstance(kmer, bytes) else kmer.encode("latin-1")
            kb.append(b)
            kl.append(len(b))
            tm.extend(tid[t] for t in lst)
            off.append(len(tm))
        return cls(np.frombuffer(b"".join(kb), dtype=np.uint8), np.array(kl, dtype=np.uint32),
                   np.array(off, dtype=np.uint64), np.array(tm, dtype=np.uint32), names,
                   np.array([int(attrs[n]["lengths"]) for n in names], dtype=np.uint64),
                   np.array([int(attrs[n]["ulength"]) for n in names], dtype=np.uint64),
                   [attrs[n].get("species", "") for n in names], summary)

    @classmethod
    def from_kmer_docs(cls, docs: list, summary: dict):
        """[{kmer, templates:[{sequence,lengths,ulengths,species}]}]  lib/kmerFinderServer.js:68-92."""
        attrs, lists = {}, {}
        for d in docs:
            lst = []
            for t in d["templates"]:
                if isinstance(t, str):
                    t = json.loads(t)          # Redis list entries are JSON strings (:184-186)
                name = t["sequence"]
                if name not in attrs:
                    attrs[name] = {"lengths": t["lengths"], "ulength": t["ulengths"],
                                   "species": t.get("species", "")}
                lst.append(name)
            lists[d["kmer"]] = lst
        return cls.from_lists(lists, attrs, summary)

    @classmethod
    def from_template_docs(cls, docs: list, summary: dict | None = None):
        """[{sequence,lengths,ulenght,species,reads:[..]}]  src/kmerPyToMongo.py:35-42.  The k-mer
        lists come out in document order (what the Mongo unwind/group of
        lib/kmerFinderServer.js:70-92 produces for an ordered collection)."""
        attrs, lists = {}, {}
        for d in docs:
            name = d["sequence"]
            attrs[name] = {"lengths": d["lengths"], "ulength": d.get("ulenght", d.get("ulengths", 0)),
                           "species": d.get("species", "")}
            for k in d.get("reads", []):
                lists.setdefault(k, []).append(name)
        if summary is None:
            summary = {"templates": len(attrs),
                       "uniqueLens": sum(int(a["ulength"]) for a in attrs.values()),
                       "totalLen": sum(int(a["lengths"]) for a in attrs.values())}
        return cls.from_lists(lists, attrs, summary)

    @classmethod
    def from_kmerfinder_map(cls, kmer_to_csv: dict, lengths: dict, ulengths: dict, descriptions: dict,
                            summary: dict | None = None):
        """{kmer: "T1,T2,.."} + side tables (lib/index.js:184-192, src/kmerPyToMongo.py:15-24)."""
        attrs = {}
        lists = {}
        for k, csv in kmer_to_csv.items():
            lst = [t for t in csv.split(",") if t]
            for t in lst:
                if t not in attrs:
                    attrs[t] = {"lengths": lengths.get(t, 0), "ulength": ulengths.get(t, 0),
                                "species": descriptions.get(t, "")}
            lists[k] = lst
        if summary is None:
            summary = {"templates": len(attrs),
                       "uniqueLens": sum(int(a["ulength"]) for a in attrs.values()),
                       "totalLen": sum(int(a["lengths"]) for a in attrs.values())}
        return cls.from_lists(lists, attrs, summary)

    def _desc(self, part: int = 0, n_parts: int = 1):
        return _abi.kj_db_desc(self.n_kmers, self.kmer_bytes.ctypes.data, self.kmer_len.ctypes.data,
                               self.list_off.ctypes.data, self.tmpl_ids.ctypes.data, self.n_templates,
                               self.lengths.ctypes.data, self.ulengths.ctypes.data,
                               self.summary["templates"], self.summary["uniqueLens"],
                               self.summary["totalLen"], part, n_parts)

    def save(self, path: str):
        """``.npz``: numpy archive of plain arrays; anything else: the versioned packed binary (kj_db_save_packed)."""
        if str(path).endswith(".npz"):
            np.savez(path, kmer_bytes=self.kmer_bytes, kmer_len=self.kmer_len, list_off=self.list_off,
                     tmpl_ids=self.tmpl_ids, lengths=self.lengths, ulengths=self.ulengths,
                     names=np.array(self.names, dtype=np.str_), species=np.array(self.species, dtype=np.str_),
                     summary=np.array(json.dumps(self.summary)))
            return
        d = self._desc()
        T = self.n_templates
        names = (C.c_char_p * max(T, 1))(*[n.encode("utf-8") for n in self.names])
        species = (C.c_char_p * max(T, 1))(*[s.encode("utf-8") for s in self.species])
        _abi.check(_abi.lib().kj_db_save_packed(str(path).encode(), C.byref(d), names, species))

    # ------------------------------------------------------------------ views
    @property
    def n_kmers(self) -> int:
        return int(self.kmer_len.size)

    @property
    def n_templates(self) -> int:
        return len(self.names)

    def to_lists(self):
        """({kmer(bytes): [name, ..]}, {name: attrs}) -- the form the CPU oracle takes (tests)."""
        lists = {}
        boff = 0
        raw = self.kmer_bytes.tobytes()
        for i in range(self.n_kmers):
            ln = int(self.kmer_len[i])
            k = raw[boff:boff + ln]
            boff += ln
            lists[k] = [self.names[t] for t in self.tmpl_ids[int(self.list_off[i]):int(self.list_off[i + 1])]]
        attrs = {n: {"lengths": int(self.lengths[i]), "ulength": int(self.ulengths[i]),
                     "species": self.species[i]} for i, n in enumerate(self.names)}
        return lists, attrs

    # ------------------------------------------------------------------ device
    def device(self, ctx: Context | None = None, part: int = 0, n_parts: int = 1):
        """kj_db handle holding the k-mers that `part` owns (all of them when n_parts == 1)."""
        ctx = ctx or default_context()
        key = (id(ctx), part, n_parts)
        h = self._dev.get(key)
        if h is None:
            L = _abi.lib()
            d = self._desc(part, n_parts)
            out = C.c_void_p()
            _abi.check(L.kj_db_create(ctx.handle, C.byref(d), C.byref(out)), ctx.handle)
            h = _DbHandle(out, ctx)
            self._dev[key] = h
        return h


class _DbHandle:
    def __init__(self, handle, ctx):
        self.handle, self.ctx = handle, ctx

    @property
    def n_kmers(self):
        return int(_abi.lib().kj_db_n_kmers(self.handle))

    @property
    def n_pairs(self):
        return int(_abi.lib().kj_db_n_pairs(self.handle))

    def __del__(self):
        try:
            if self.handle:
                _abi.lib().kj_db_free(self.handle)
                self.handle = None
        except Exception:
            pass


def load(path: str, summary=None, **side) -> TemplateDB:
    """Load a template DB from ``.npz`` (packed) or ``.json`` (any of the reference's layouts).
    ``summary``: dict or path of a Summary JSON; ``side``: lengths= ulengths= descriptions= (dicts
    or JSON paths) for the original KmerFinder ``{kmer: "T1,T2"}`` map."""
    def _json(x):
        if isinstance(x, (str, bytes)):
            with open(x) as f:
                return json.load(f)
        return x

    if str(path).endswith(".npz"):
        z = np.load(path, allow_pickle=False)
        return TemplateDB(z["kmer_bytes"], z["kmer_len"], z["list_off"], z["tmpl_ids"], [str(x) for x in z["names"]],
                          z["lengths"], z["ulengths"], [str(x) for x in z["species"]], json.loads(str(z["summary"])))
    with open(path, "rb") as f:
        magic = f.read(8)
    if magic[:4] == b"KJDB":
        return _load_packed(path)
    doc = _json(path)
    summary = _json(summary) if summary is not None else None
    if isinstance(doc, dict):
        return TemplateDB.from_kmerfinder_map(doc, _json(side.get("lengths", {})), _json(side.get("ulengths", {})),
                                              _json(side.get("descriptions", {})), summary)
    if doc and "reads" in doc[0]:
        return TemplateDB.from_template_docs(doc, summary)
    if summary is None:
        raise ValueError("per-k-mer documents need a Summary record")
    return TemplateDB.from_kmer_docs(doc, summary)


def _load_packed(path: str) -> TemplateDB:
    """The packed binary (kj_dbio.cu PackedHeader) read with numpy: header of 8 + 9 * 8 bytes, then plain arrays."""
    with open(path, "rb") as f:
        if f.read(8) != b"KJDBv001":
            raise ValueError(f"{path}: not a kmerjs_b200 packed database (magic / version)")
        n_kmers, kbytes, n_pairs, T, nb, sb, s_t, s_u, s_l = np.fromfile(f, dtype=np.uint64, count=9).tolist()
        kmer_len = np.fromfile(f, dtype=np.uint32, count=n_kmers)
        kmer_bytes = np.fromfile(f, dtype=np.uint8, count=kbytes)
        list_off = np.fromfile(f, dtype=np.uint64, count=n_kmers + 1)
        tmpl_ids = np.fromfile(f, dtype=np.uint32, count=n_pairs)
        lengths = np.fromfile(f, dtype=np.uint64, count=T)
        ulengths = np.fromfile(f, dtype=np.uint64, count=T)
        names = f.read(nb).split(b"\0")[:T]
        species = f.read(sb).split(b"\0")[:T]
    return TemplateDB(kmer_bytes, kmer_len, list_off, tmpl_ids, [x.decode("utf-8") for x in names], lengths, ulengths,
                      [x.decode("utf-8") for x in species], {"templates": s_t, "uniqueLens": s_u, "totalLen": s_l})


class NativeTemplateDB:
    """A template DB loaded by the library itself (kj_db_load): file -> host arrays -> HBM inside the C ABI.  Same
    surface as TemplateDB for scoring (names, species, lengths, ulengths, summary, device())."""

    def __init__(self, path: str, summary: str | None = None, fmt: int = _abi.KJ_DB_AUTO, ctx: Context | None = None,
                 part: int = 0, n_parts: int = 1):
        self.ctx = ctx or default_context()
        L = _abi.lib()
        out = C.c_void_p()
        _abi.check(L.kj_db_load(self.ctx.handle, str(path).encode(), fmt, summary.encode() if summary else None,
                                part, n_parts, C.byref(out)), self.ctx.handle)
        self._h = _DbHandle(out, self.ctx)
        self._key = (part, n_parts)
        T = int(L.kj_db_n_templates(out))
        self.names, self.species = [], []
        self.lengths = np.zeros(T, dtype=np.uint64)
        self.ulengths = np.zeros(T, dtype=np.uint64)
        nm, sp, ln, ul = C.c_char_p(), C.c_char_p(), C.c_uint64(), C.c_uint64()
        for t in range(T):
            _abi.check(L.kj_db_template(out, t, C.byref(nm), C.byref(sp), C.byref(ln), C.byref(ul)))
            self.names.append((nm.value or b"").decode("utf-8"))
            self.species.append((sp.value or b"").decode("utf-8"))
            self.lengths[t], self.ulengths[t] = ln.value, ul.value
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _abi.check(L.kj_db_summary(out, C.byref(a), C.byref(b), C.byref(c)))
        self.summary = {"templates": int(a.value), "uniqueLens": int(b.value), "totalLen": int(c.value)}

    @property
    def n_templates(self) -> int:
        return len(self.names)

    @property
    def n_kmers(self) -> int:
        return self._h.n_kmers

    def device(self, ctx: Context | None = None, part: int = 0, n_parts: int = 1):
        if (part, n_parts) != self._key or (ctx is not None and ctx is not self.ctx):
            raise ValueError("a natively loaded DB lives on the context / part it was loaded for")
        return self._h


def load_native(path: str, summary: str | None = None, fmt: int = _abi.KJ_DB_AUTO, ctx: Context | None = None,
                part: int = 0, n_parts: int = 1) -> NativeTemplateDB:
    """kj_db_load: any of the reference's JSON layouts or the packed binary, straight into GPU memory."""
    return NativeTemplateDB(path, summary, fmt, ctx, part, n_parts)
