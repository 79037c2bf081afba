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
