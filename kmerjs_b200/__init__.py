"""kmerjs_b200 -- the kmerjs hot path (FASTQ -> prefix-filtered k-mer counts -> KmerFinder template
scoring) on NVIDIA B200 (sm_100a), behind the reference's own interface.

    from kmerjs_b200 import kmerjs, KmerJS, KmerFinderClient, complement

Everything that computes runs in libkmerjs_b200.so (hand-written CUDA, C ABI in
include/kmerjs_b200.h).  There is no CPU fallback."""
from __future__ import annotations

import os

from .kmers import (KmerJS, KmerMap, Promise, complement, complementMap, jsonToStrMap, mapToJSON,  # noqa: F401
                    objectToMap, stringToMap)
from .kmer_finder_client import KmerFinderClient  # noqa: F401
from .stats import etta, fastp, zScore  # noqa: F401
from .db import TemplateDB, load as load_db  # noqa: F401
from .matching import NoHitsError  # noqa: F401

__all__ = ["kmerjs", "KmerJS", "KmerFinderClient", "complement", "complementMap", "jsonToStrMap",
           "stringToMap", "objectToMap", "mapToJSON", "zScore", "fastp", "etta", "TemplateDB", "load_db",
           "NoHitsError", "output_file_text"]


def output_file_text(kmerMap: dict) -> str:
    """The ``out`` pseudo-JSON of lib/index.js:381-388: ``{\\nK: V,K: V,}\\n``."""
    return "{\n" + "".join(f"{k}: {v}," for k, v in kmerMap.items()) + "}\n"


class _KmerjsResult(Promise):
    """What ``kmerjs(...)`` returns: thenable (README.md:12-16 shows it assigned directly), with
    ``.map`` once done."""

    @property
    def map(self):
        return self.result()


def kmerjs(fastqPath, prefix="ATGAC", k=16, step=1, output=None):
    """README.md:12-16 entry point ``kmerjs(fastqPath, prefix, k, step, output)`` (documented by the
    reference, not implemented there).  Resolves to the k-mer map; a truthy
    ``output`` also writes it in the lib/index.js:381-388 format."""
    res = _KmerjsResult()
    job = KmerJS(fastqPath, prefix, k, step, 1, False)

    def done(m):
        if output:
            with open(os.fspath(output), "w") as f:
                f.write(output_file_text(m))
        res.set_result(m)

    def fail(exc):
        res.set_exception(exc)

    job.readFile().promise.then(done, fail)
    res.job = job
    return res
