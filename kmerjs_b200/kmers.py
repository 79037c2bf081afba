"""Host-side mirror of lib/kmers.js (josl/kmerjs): same names, argument meaning and results; the
work runs on the GPU through libkmerjs_b200.so.

Reference interface replaced (paths relative to the kmerjs repository):
  complementMap / complement                  lib/kmers.js:12-17,31-38
  jsonToStrMap / stringToMap / objectToMap /
  mapToJSON                                   lib/kmers.js:19-29,40-54
  class KmerJS (ctor, kmersInLine, readFile)  lib/kmers.js:56-186

JavaScript ``Map<string, number>`` becomes an insertion-ordered ``dict``; a Promise becomes a
:class:`Promise` (a ``concurrent.futures.Future`` with ``then``); the progress-stream emitter
becomes :class:`ProgressEvent` (``on('progress', cb)``).
"""
from __future__ import annotations

import json
import os
import threading
from concurrent.futures import Future
from decimal import Decimal

from . import _abi
from .counts import Counts

complementMap = {"A": "T", "T": "A", "G": "C", "C": "G"}      # lib/kmers.js:12-17
_COMP = str.maketrans("ATGC", "TACG")


def complement(string: str) -> str:
    """Reverse complement; only upper-case A,T,G,C are mapped (lib/kmers.js:31-38)."""
    return string.translate(_COMP)[::-1]


def objToStrMap(obj) -> dict:
    return dict(obj)


def jsonToStrMap(jsonStr) -> dict:           # lib/kmers.js:27-29 (takes an object, despite the name)
    return objToStrMap(jsonStr)


def stringToMap(string: str) -> dict:        # lib/kmers.js:40-42
    return objToStrMap(json.loads(string))


def objectToMap(obj) -> dict:                # lib/kmers.js:43-45
    return objToStrMap(obj)


def mapToJSON(strMap) -> dict:               # lib/kmers.js:46-54: plain object, Map order
    return dict(strMap)


class Promise(Future):
    """Future with the two Promise methods the reference's callers use."""

    def then(self, on_ok, on_err=None):
        out = Promise()

        def done(f):
            try:
                v = f.result()
            except BaseException as exc:  # noqa: BLE001 - mirrors promise rejection
                if on_err is None:
                    out.set_exception(exc)
                else:
                    try:
                        out.set_result(on_err(exc))
                    except BaseException as e2:  # noqa: BLE001
                        out.set_exception(e2)
                return
            try:
                out.set_result(on_ok(v))
            except BaseException as exc:  # noqa: BLE001
                out.set_exception(exc)

        self.add_done_callback(done)
        return out

    def catch(self, on_err):
        return self.then(lambda v: v, on_err)


class ProgressEvent:
    """Stand-in for the progress-stream object returned as ``event`` (lib/kmers.js:108-110,181-184)."""

    def __init__(self):
        self._cbs = {}

    def on(self, name, cb):
        self._cbs.setdefault(name, []).append(cb)
        return self

    def emit(self, name, *args):
        for cb in self._cbs.get(name, []):
            cb(*args)


class KmerMap(dict):
    """The resolved k-mer map: a dict (Map insertion order) that remembers the device table it
    was exported from, so findFirstMatch can score it without a host round trip."""

    counts: Counts | None = None


class KmerJS:
    """lib/kmers.js:56-186."""

    def __init__(self, fastq="", preffix="ATGAC", length=16, step=1, coverage=1, progress=True,
                 env="node"):
        self.fastq = fastq
        self.preffix = preffix
        self.kmerLength = length
        self.step = step
        self.progress = progress
        self.coverage = coverage            # stored, never read on this path (as in the reference)
        self.evalue = Decimal("0.05")
        self.kmerMap = KmerMap()
        self.kmerMapSize = 0
        self.env = env
        if env == "browser":
            self.fileDataRead = 0
        self.lines = 0
        self.bytesRead = 0

    def _params(self):
        return dict(prefix=self.preffix.encode("latin-1"), k=int(self.kmerLength), step=int(self.step))

    def kmersInLine(self, line: str) -> None:
        """lib/kmers.js:88-100: count the windows of ``line`` (this strand only) into kmerMap."""
        prefix = self.preffix
        sentinel = None
        if "\n" in line or "\n" in prefix:
            # a JS string may hold '\n' as an ordinary character (test/kmers.js:14-15 does); the byte
            # stream API splits on it, so it travels as a byte that occurs nowhere else
            used = set(line) | set(prefix)
            sentinel = next(chr(i) for i in range(1, 256) if chr(i) not in used and chr(i) not in "ATGC\n")
            line = line.replace("\n", sentinel)
            prefix = prefix.replace("\n", sentinel)
        # base_line=1: the buffer starts on line index 1, i.e. it is a sequence line of the FSM;
        # the length > 1 gate belongs to readFile (lib/kmers.js:151), not to kmersInLine
        c = Counts(prefix.encode("latin-1"), int(self.kmerLength), int(self.step),
                   flags=_abi.KJ_F_FORWARD_ONLY | _abi.KJ_F_NO_LINE_GATE, base_line=1)
        try:
            c.add_host(line.encode("latin-1"), final=True)
            c.finish()
            for kmer, n in c.to_dict().items():
                if sentinel is not None:
                    kmer = kmer.replace(sentinel, "\n")
                self.kmerMap[kmer] = self.kmerMap.get(kmer, 0) + n
        finally:
            c.free()

    def _count_file(self) -> Counts:
        c = Counts(**self._params())
        if isinstance(self.fastq, (bytes, bytearray, memoryview)):
            c.add_host(self.fastq, final=True)       # env 'browser': a File/Blob's bytes
        else:
            c.add_file(os.fspath(self.fastq))
        c.finish()
        return c

    def readFile(self):
        """lib/kmers.js:106-185.  Returns an object with ``promise`` (resolves to the k-mer map)
        and ``event`` (progress emitter)."""
        promise = Promise()
        event = ProgressEvent()

        def work():
            try:
                c = self._count_file()
                m = KmerMap(c.to_dict())
                m.counts = c
                self.kmerMap = m
                self.kmerMapSize = len(m)              # lib/kmers.js:177
                self.lines = c.lines                   # lib/kmers.js:164-165
                self.bytesRead = c.bytes_read
                if self.env == "node" and self.progress:
                    # one line instead of one write per FASTQ line (lib/kmers.js:166-169,174-176)
                    print(f"Lines: {self.lines} / Kmers: {len(m)}\r\n                               ")
                event.emit("progress", {"percentage": 100.0, "transferred": self.bytesRead})
                promise.set_result(m)
            except BaseException as exc:  # noqa: BLE001
                promise.set_exception(exc)

        threading.Thread(target=work, daemon=True).start()
        return _ReadHandle(promise, event)


class _ReadHandle:
    def __init__(self, promise, event):
        self.promise = promise
        self.event = event

    def __getitem__(self, key):      # allow handle['promise'] like the JS object literal
        return getattr(self, key)
