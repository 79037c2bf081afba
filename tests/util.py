"""Helpers shared by the tests: seeded FASTQ fuzz, synthetic template DBs, device buffers."""
import os
import random

import numpy as np


def emulated() -> bool:
    return os.environ.get("KMERJS_B200_EMU") == "1"


class DevBuf:
    """Bytes in device memory (torch CUDA tensor; host memory under the tools/cuemu harness)."""

    def __init__(self, data):
        a = np.frombuffer(bytes(data), dtype=np.uint8)
        if emulated():
            pad = np.zeros(a.size + 64, dtype=np.uint8)
            off = (-pad.ctypes.data) % 16
            self._keep = pad
            pad[off:off + a.size] = a
            self.ptr = pad.ctypes.data + off
        else:
            import torch
            self._keep = torch.from_numpy(a.copy()).cuda() if a.size else torch.zeros(16, dtype=torch.uint8, device="cuda")
            self.ptr = self._keep.data_ptr()
        self.n = a.size


def dev_u64(arr):
    """u64 numpy array -> (ptr, keepalive) in device memory; .back() reads it."""
    a = np.ascontiguousarray(arr, dtype=np.uint64)
    if emulated():
        b = a.copy()
        return b.ctypes.data, b, (lambda: b.copy())
    import torch
    t = torch.from_numpy(a.view(np.int64).copy()).cuda()
    return t.data_ptr(), t, (lambda: t.cpu().numpy().view(np.uint64))


def random_fastq(rng: random.Random, n_reads: int, *, min_len=0, max_len=140, p_n=0.01, p_lower=0.0,
                 crlf=False, blank_lines=0.0, trailing_newline=True, plant=(b"ATGAC", 0.3),
                 alphabet=b"ACGT") -> bytes:
    """Illumina-shaped records with the irregularities the reference tolerates silently."""
    out = []
    for r in range(n_reads):
        L = rng.randint(min_len, max_len)
        seq = bytearray(rng.choice(alphabet) for _ in range(L))
        if plant and L >= 24:
            motif, p = plant
            for _ in range(3):
                if rng.random() < p:
                    pos = rng.randrange(0, L - len(motif))
                    m = motif if rng.random() < 0.5 else bytes(reversed(motif.translate(bytes.maketrans(b"ATGC", b"TACG"))))
                    seq[pos:pos + len(m)] = m
        for i in range(L):
            x = rng.random()
            if x < p_n:
                seq[i] = ord("N")
            elif x < p_n + p_lower:
                seq[i] = seq[i] | 0x20
        qual = bytes(rng.randint(35, 73) for _ in range(L))       # '@' and '+' included
        eol = b"\r\n" if crlf else b"\n"
        out.append(b"@r%d" % r + eol + bytes(seq) + eol + b"+" + eol + qual + eol)
        if rng.random() < blank_lines:
            out.append(b"\n")
    data = b"".join(out)
    if not trailing_newline and data.endswith(b"\n"):
        data = data[:-1]
    return data


def synthetic_db(query_keys, rng: random.Random, n_templates=40, *, decoys=200, k=16, share=0.35,
                 summary=None):
    """Seeded template DB over (a subset of) the query keys plus decoy k-mers: per-template k-mer
    sets of very different sizes, shared k-mers, ties.  Returns (kmer_lists, attrs, summary) in the
    form oracle.kmer_oracle.TemplateDB takes."""
    names = [f"T{idx:04d}" for idx in range(n_templates)]
    keys = list(query_keys)
    rng.shuffle(keys)
    kmer_lists = {}
    weights = [rng.random() ** 3 for _ in names]
    for key in keys:
        members = [n for n, w in zip(names, weights) if rng.random() < w * share]
        if members:
            rng.shuffle(members)
            kmer_lists[key] = members
    for _ in range(decoys):
        key = bytes(rng.choice(b"ACGT") for _ in range(k))
        if key not in kmer_lists:
            kmer_lists[key] = rng.sample(names, rng.randint(1, min(4, n_templates)))
    # a pair of templates with identical k-mer sets -> uScore ties, decided by first-encounter order
    if n_templates >= 2:
        for key, lst in kmer_lists.items():
            if names[0] in lst and names[1] not in lst:
                lst.append(names[1])
            elif names[1] in lst and names[0] not in lst:
                lst.append(names[0])
    items = list(kmer_lists.items())
    rng.shuffle(items)
    kmer_lists = dict(items)
    per_t = {n: 0 for n in names}
    for lst in kmer_lists.values():
        for n in lst:
            per_t[n] += 1
    attrs = {n: {"lengths": 1000 + 50 * per_t[n] + rng.randint(0, 500), "ulength": per_t[n] + rng.randint(1, 50),
                 "species": f"Species {n}"} for n in names}
    if summary is None:
        summary = {"templates": n_templates, "uniqueLens": sum(a["ulength"] for a in attrs.values()) * 20,
                   "totalLen": sum(a["lengths"] for a in attrs.values())}
    return kmer_lists, attrs, summary
