"""ctypes wrapper over oracle/libkmer_oracle.so (kmer_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmer_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kmer_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "libkmer_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.ko_count.restype = ctypes.c_void_p
        L.ko_count.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_int,
                               ctypes.c_int, ctypes.c_int]
        for f in ("ko_size", "ko_lines", "ko_key_bytes"):
            getattr(L, f).restype = ctypes.c_uint64
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.ko_export.restype = None
        L.ko_export.argtypes = [ctypes.c_void_p] * 4
        L.ko_free.restype = None
        L.ko_free.argtypes = [ctypes.c_void_p]
        L.ko_complement.restype = None
        L.ko_complement.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
        _lib = L
    return _lib


def count_fastq(data, prefix: bytes = b"ATGAC", k: int = 16, step: int = 1):
    """Same contract as kmer_oracle.count_fastq: (insertion-ordered {bytes: int}, n_lines).
    ``data`` may be bytes or a numpy uint8 array."""
    import numpy as np
    L = lib()
    arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) \
        else np.ascontiguousarray(data, dtype=np.uint8)
    h = L.ko_count(arr.ctypes.data, arr.size, prefix, len(prefix), k, step)
    try:
        n = L.ko_size(h)
        kb = L.ko_key_bytes(h)
        keys = np.empty(max(kb, 1), dtype=np.uint8)
        lens = np.empty(max(n, 1), dtype=np.uint32)
        cnts = np.empty(max(n, 1), dtype=np.uint64)
        L.ko_export(h, keys.ctypes.data, lens.ctypes.data, cnts.ctypes.data)
        lines = L.ko_lines(h)
    finally:
        L.ko_free(h)
    out = {}
    raw = keys.tobytes()
    off = 0
    for i in range(n):
        ln = int(lens[i])
        out[raw[off:off + ln]] = int(cnts[i])
        off += ln
    return out, int(lines)


def count_only(data, prefix: bytes = b"ATGAC", k: int = 16, step: int = 1):
    """Run the count and return (n_unique, n_lines) without materialising the map (timing leg)."""
    import numpy as np
    L = lib()
    arr = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) \
        else np.ascontiguousarray(data, dtype=np.uint8)
    h = L.ko_count(arr.ctypes.data, arr.size, prefix, len(prefix), k, step)
    try:
        return int(L.ko_size(h)), int(L.ko_lines(h))
    finally:
        L.ko_free(h)


def complement(s: bytes) -> bytes:
    buf = ctypes.create_string_buffer(len(s))
    lib().ko_complement(s, len(s), buf)
    return buf.raw


def find_matches_fast(qmap, db, max_hits: int = 100):
    """The scoring path with the integer loop in C (ko_wta) and the exact-decimal gate + rows in
    Python (kmer_oracle.match_summary): same results as kmer_oracle.first_match + find_matches
    (checked in tests/test_oracle_golden.py), for databases the Python loop cannot follow in seconds.
    qmap: insertion-ordered {kmer bytes: count}; db: kmer_oracle.TemplateDB.
    Returns (first OrderedDict name -> {uScore, tScore}, hits, rows, error text or None)."""
    import numpy as np
    from collections import OrderedDict
    import kmer_oracle as kp
    L = lib()
    if not hasattr(L, "_wta_ready"):
        L.ko_wta.restype = ctypes.c_uint32
        L._wta_ready = True
    names = list(db.attrs.keys())
    tid = {n: i for i, n in enumerate(names)}
    T = len(names)
    qcount = np.fromiter(qmap.values(), dtype=np.uint64, count=len(qmap))
    qoff = np.zeros(len(qmap) + 1, dtype=np.uint64)
    flat = []
    for i, kmer in enumerate(qmap):
        lst = db.kmer_lists.get(kmer, ())
        flat.extend(tid[n] for n in lst)
        qoff[i + 1] = len(flat)
    qt = np.asarray(flat, dtype=np.uint32) if flat else np.zeros(1, dtype=np.uint32)
    return wta_arrays(qcount, qoff, qt, names, db.attrs, db.summary, len(qmap), max_hits)


def wta_arrays(qcount, qoff, qt, names, attrs, summary, kmer_map_size, max_hits: int = 100):
    """find_matches_fast on arrays: query entries in Map order with their template-id lists (CSR)."""
    import numpy as np
    from collections import OrderedDict
    import kmer_oracle as kp
    L = lib()
    L.ko_wta.restype = ctypes.c_uint32
    T = len(names)
    rows = []
    evalue = kp.BN(0.05)

    def gate(rnd, t, u, tau, hits, u1, t1):
        name = names[t]
        a = attrs[name]
        match = {"uScore": int(u), "tScore": int(tau), "lengths": a["lengths"], "ulength": a["ulength"],
                 "species": a["species"]}
        first = {name: {"uScore": int(u1), "tScore": int(t1)}}
        row = kp.match_summary(kmer_map_size, first, name, match, int(hits), summary, evalue)
        if row is not None and evalue.cmp(kp.BN(row["probability"])) >= 0:
            rows.append(row)
            return 1
        return 0

    GATE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                            ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64)
    out = np.zeros(4 * max(max_hits, 1), dtype=np.uint64)
    order = np.zeros(max(T, 1), dtype=np.uint32)
    u0 = np.zeros(max(T, 1), dtype=np.uint64)
    t0 = np.zeros(max(T, 1), dtype=np.uint64)
    status, n_order, hits0 = ctypes.c_int(0), ctypes.c_uint32(0), ctypes.c_uint64(0)
    qcount = np.ascontiguousarray(qcount, dtype=np.uint64)
    qoff = np.ascontiguousarray(qoff, dtype=np.uint64)
    qt = np.ascontiguousarray(qt, dtype=np.uint32)
    L.ko_wta(ctypes.c_uint64(len(qcount)), ctypes.c_void_p(qcount.ctypes.data), ctypes.c_void_p(qoff.ctypes.data),
             ctypes.c_void_p(qt.ctypes.data), ctypes.c_uint32(T), ctypes.c_uint32(max_hits), GATE(gate),
             ctypes.c_void_p(out.ctypes.data), ctypes.byref(status), ctypes.c_void_p(order.ctypes.data),
             ctypes.byref(n_order), ctypes.c_void_p(u0.ctypes.data), ctypes.c_void_p(t0.ctypes.data),
             ctypes.byref(hits0))
    first = OrderedDict((names[int(t)], {"uScore": int(u0[int(t)]), "tScore": int(t0[int(t)])})
                        for t in order[: n_order.value])
    err = None
    if hits0.value == 0:
        err = "No hits were found!"
    elif status.value == 1:
        err = "No hits were found! (nHits === 0)"
    elif status.value == 2:
        err = "No hits were found! (kmerResults.length === 0)"
    return first, int(hits0.value), rows, err
