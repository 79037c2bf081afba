"""Thin object wrapper over the kj_counts_* entry points (include/kmerjs_b200.h)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from .context import Context, default_context


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Counts:
    """The k-mer count table of one job (the reference's ``kmerMap``, lib/kmers.js:76), resident
    in HBM.  add_* calls feed the FASTQ byte stream; finish() freezes it."""

    def __init__(self, prefix: bytes = b"ATGAC", k: int = 16, step: int = 1, *, flags: int = 0,
                 base_line: int = 0, base_col: int = 0, capacity_hint: int = 0,
                 ctx: Context | None = None):
        self.ctx = ctx or default_context()
        self._L = _abi.lib()
        if isinstance(prefix, str):
            prefix = prefix.encode("latin-1")
        self.prefix, self.k, self.step, self.flags = bytes(prefix), int(k), int(step), int(flags)
        p = _abi.kj_count_params(self.prefix, len(self.prefix), self.k, self.step, self.flags,
                                 base_line, base_col, capacity_hint)
        self._keep = p
        h = C.c_void_p()
        _abi.check(self._L.kj_counts_create(self.ctx.handle, C.byref(p), C.byref(h)), self.ctx.handle)
        self.handle = h
        self.finished = False

    # -- feeding -------------------------------------------------------------------------------
    def add_host(self, data, own_n: int | None = None, final: bool = True):
        """data: bytes / bytearray / numpy uint8 / pinned torch uint8 tensor (host memory)."""
        ptr, n, keep = _host_pointer(data)
        own = n if own_n is None else own_n
        _abi.check(self._L.kj_counts_add_buffer(self.handle, ptr, n, own, _abi.KJ_MEM_HOST,
                                                1 if final else 0), self.ctx.handle)
        del keep
        return self

    def add_device(self, dev_ptr: int, n: int, own_n: int | None = None, final: bool = True):
        """dev_ptr: 16-byte aligned device address (e.g. torch tensor .data_ptr())."""
        own = n if own_n is None else own_n
        _abi.check(self._L.kj_counts_add_buffer(self.handle, C.c_void_p(dev_ptr), n, own,
                                                _abi.KJ_MEM_DEVICE, 1 if final else 0),
                   self.ctx.handle)
        return self

    def add_file(self, path: str):
        _abi.check(self._L.kj_counts_add_file(self.handle, str(path).encode()), self.ctx.handle)
        return self

    def finish(self):
        _abi.check(self._L.kj_counts_finish(self.handle), self.ctx.handle)
        self.finished = True
        return self

    # -- results -------------------------------------------------------------------------------
    @property
    def size(self) -> int:
        return int(self._L.kj_counts_size(self.handle))

    @property
    def lines(self) -> int:
        return int(self._L.kj_counts_lines(self.handle))

    @property
    def bases(self) -> int:
        return int(self._L.kj_counts_bases(self.handle))

    @property
    def bytes_read(self) -> int:
        return int(self._L.kj_counts_bytes_read(self.handle))

    @property
    def occurrences(self) -> int:
        return int(self._L.kj_counts_occurrences(self.handle))

    def export_arrays(self):
        """(keys uint8[n,32], key_len uint32[n], counts uint64[n]) in first-insertion order."""
        n = self.size
        keys = np.zeros((max(n, 1), 32), dtype=np.uint8)
        lens = np.zeros(max(n, 1), dtype=np.uint32)
        cnts = np.zeros(max(n, 1), dtype=np.uint64)
        _abi.check(self._L.kj_counts_export(self.handle, _ptr(keys), _ptr(lens), _ptr(cnts)),
                   self.ctx.handle)
        return keys[:n], lens[:n], cnts[:n]

    def to_dict(self) -> dict:
        """{kmer(str): count} in Map insertion order (mapToJSON layout, lib/kmers.js:46-54)."""
        keys, lens, cnts = self.export_arrays()
        n = len(lens)
        if n == 0:
            return {}
        raw = keys.tobytes()
        if int(lens.min()) == int(lens.max()):
            ln = int(lens[0])
            ks = [raw[32 * i:32 * i + ln].decode("latin-1") for i in range(n)]
        else:
            ks = [raw[32 * i:32 * i + int(lens[i])].decode("latin-1") for i in range(n)]
        return dict(zip(ks, cnts.tolist()))

    def alive(self) -> np.ndarray:
        n = self.size
        a = np.zeros(max(n, 1), dtype=np.uint8)
        _abi.check(self._L.kj_counts_alive(self.handle, _ptr(a)), self.ctx.handle)
        return a[:n]

    # -- exchange (multi-GPU) -------------------------------------------------------------------
    def partition(self, n_parts: int):
        """(device pointer of the owner-grouped 24-byte records, sizes per part)."""
        ptr = C.c_void_p()
        sizes = (C.c_uint64 * n_parts)()
        _abi.check(self._L.kj_counts_partition(self.handle, n_parts, C.byref(ptr), sizes),
                   self.ctx.handle)
        return int(ptr.value or 0), [int(x) for x in sizes]

    def merge_records(self, dev_ptr: int, n: int):
        _abi.check(self._L.kj_counts_merge_records(self.handle, C.c_void_p(dev_ptr), n), self.ctx.handle)
        self.finished = False

    def merge_host_records(self, records: np.ndarray):
        """records: uint64[n, 3] = (2-bit key, count, ordinal)."""
        records = np.ascontiguousarray(records, dtype=np.uint64)
        n = records.size // 3
        if n:
            _abi.check(self._L.kj_counts_merge_host_records(self.handle, _ptr(records), n), self.ctx.handle)
            self.finished = False

    def irregular_records(self) -> np.ndarray:
        n = int(self._L.kj_counts_irregular_size(self.handle))
        rec = np.zeros(max(n, 1) * 56, dtype=np.uint8)
        _abi.check(self._L.kj_counts_irregular_export(self.handle, _ptr(rec)), self.ctx.handle)
        return rec[:n * 56]

    def merge_irregular(self, records: np.ndarray, part: int = 0, n_parts: int = 1):
        """Merge byte-string k-mer records; with n_parts > 1 only those this part owns."""
        n = records.size // 56
        if n:
            records = np.ascontiguousarray(records, dtype=np.uint8)
            _abi.check(self._L.kj_counts_irregular_merge_part(self.handle, _ptr(records), n, part, n_parts),
                       self.ctx.handle)
            self.finished = False

    def partition_segments(self, n_parts: int, dev_ptr: int, cap_reg: int, cap_irr: int):
        """Fixed-capacity exchange, sender side (stream-ordered, no finish() needed): n_parts segments of
        segment_bytes(cap_reg, cap_irr) bytes at dev_ptr."""
        _abi.check(self._L.kj_counts_partition_segments(self.handle, n_parts, C.c_void_p(dev_ptr), cap_reg, cap_irr),
                   self.ctx.handle)

    def merge_segments(self, dev_ptr: int, n_parts: int, cap_reg: int, cap_irr: int):
        """Receiver side: merge the n_parts segments this rank got; totals and overflow surface in finish()."""
        _abi.check(self._L.kj_counts_merge_segments(self.handle, C.c_void_p(dev_ptr), n_parts, cap_reg, cap_irr),
                   self.ctx.handle)
        self.finished = False

    def export_matched_segment(self, dbh, dev_ptr: int, cap_entries: int, cap_pairs: int) -> bool:
        """The matched entries of this (unfinished) handle against the device DB `dbh` (TemplateDB.device(...)) as one
        fixed-capacity segment, straight from the hash table (kj_counts_export_matched_segment).  False: the short cut does
        not apply, finish() first."""
        rc = self._L.kj_counts_export_matched_segment(self.handle, dbh.handle, C.c_void_p(dev_ptr), cap_entries, cap_pairs)
        if rc == 1:
            return False
        _abi.check(rc, self.ctx.handle)
        return True

    def set_totals(self, lines: int, bases: int, occurrences: int, bytes_read: int):
        _abi.check(self._L.kj_counts_set_totals(self.handle, lines, bases, occurrences, bytes_read),
                   self.ctx.handle)

    def free(self):
        if getattr(self, "handle", None):
            self._L.kj_counts_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def segment_bytes(cap_reg: int, cap_irr: int) -> int:
    return int(_abi.lib().kj_segment_bytes(cap_reg, cap_irr))


def _host_pointer(data):
    """(c_void_p, nbytes, keepalive) of a host byte container."""
    try:
        import torch
        if isinstance(data, torch.Tensor):
            if data.is_cuda:
                raise TypeError("device tensor passed to add_host; use add_device")
            t = data.contiguous().view(torch.uint8)
            return C.c_void_p(t.data_ptr()), t.numel(), t
    except ImportError:
        pass
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        return C.c_void_p(a.ctypes.data), a.size, a
    if isinstance(data, (bytes, bytearray, memoryview)):
        a = np.frombuffer(data, dtype=np.uint8)
        return C.c_void_p(a.ctypes.data), a.size, (a, data)
    raise TypeError(f"unsupported host buffer type {type(data)!r}")


def count_newlines_device(dev_ptr: int, n: int, ctx: Context | None = None):
    """(number of '\\n', offset + 1 of the last one) of a device buffer."""
    ctx = ctx or default_context()
    cnt, last = C.c_uint64(), C.c_uint64()
    _abi.check(_abi.lib().kj_count_newlines(ctx.handle, C.c_void_p(dev_ptr), n, _abi.KJ_MEM_DEVICE,
                                            C.byref(cnt), C.byref(last)), ctx.handle)
    return int(cnt.value), int(last.value)
