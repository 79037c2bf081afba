"""Synthetic workloads for the bench and the large-size tests: a random genome,
Illumina-shaped reads sampled from it on the device (kj_synth_*), and a template DB whose first
template is that genome.  Deterministic in (seed, sizes)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from .context import Context, default_context

_EMU = os.environ.get("KMERJS_B200_EMU") == "1"       # tools/cuemu developer harness: "device" = host


class _DeviceBytes:
    """n bytes of device memory (torch owns it: PyTorch is the allocator, not the product)."""

    def __init__(self, n: int, device: int):
        self.n = n
        if _EMU:
            self._t = np.zeros(n + 64, dtype=np.uint8)
            self.ptr = self._t.ctypes.data + ((-self._t.ctypes.data) % 16)
        else:
            import torch
            self._t = torch.empty(max(n, 16), dtype=torch.uint8, device=f"cuda:{device}")
            self.ptr = self._t.data_ptr()

    def to_host(self, n: int | None = None) -> bytes:
        n = self.n if n is None else n
        if _EMU:
            return C.string_at(self.ptr, n)
        return self._t[:n].cpu().numpy().tobytes()


class Workload:
    def __init__(self, n_reads: int, genome_len: int = 5_000_000, read_len: int = 150, seed: int = 0x6B6D6572,
                 sub_rate: float = 0.005, n_rate: float = 1e-4, lead_n_rate: float = 0.02,
                 first_read: int = 0, ctx: Context | None = None):
        self.ctx = ctx or default_context()
        L = _abi.lib()
        self.n_reads, self.read_len, self.seed, self.genome_len = n_reads, read_len, seed, genome_len
        self.genome = _DeviceBytes(genome_len, self.ctx.device)
        _abi.check(L.kj_synth_genome(self.ctx.handle, seed, C.c_void_p(self.genome.ptr), genome_len), self.ctx.handle)
        p = _abi.kj_synth_params(seed, n_reads, read_len, first_read, self.genome.ptr, genome_len,
                                 sub_rate, n_rate, lead_n_rate)
        nb = C.c_uint64()
        _abi.check(L.kj_synth_size(self.ctx.handle, C.byref(p), C.byref(nb)), self.ctx.handle)
        self.n_bytes = int(nb.value)
        self.record_bytes = self.n_bytes // max(n_reads, 1)
        self.fastq = _DeviceBytes(self.n_bytes, self.ctx.device)
        _abi.check(L.kj_synth_generate(self.ctx.handle, C.byref(p), C.c_void_p(self.fastq.ptr), self.n_bytes),
                   self.ctx.handle)
        self.fastq_ptr = self.fastq.ptr
        self.bases = n_reads * read_len

    def host_bytes(self, n_reads: int | None = None) -> bytes:
        """The first n_reads records as host bytes (CPU-baseline sample, oracle checks)."""
        n = self.n_reads if n_reads is None else min(n_reads, self.n_reads)
        return self.fastq.to_host(n * self.record_bytes)

    def genome_host(self) -> bytes:
        return self.genome.to_host()


def template_db_from_genome(genome: bytes, n_templates: int = 64, prefix: bytes = b"ATGAC", k: int = 16,
                            seed: int = 7, divergence: float = 0.03):
    """KmerFinder-style DB over one genome: template 0 is the genome itself, templates 1.. are
    point-mutated relatives (shared k-mers, decreasing similarity), the rest random decoys.  k-mer
    sets are the prefix-filtered k-mers of the template's forward strand.  Returns a TemplateDB."""
    from .db import TemplateDB
    rng = np.random.default_rng(seed)
    g = np.frombuffer(genome, dtype=np.uint8)
    code = np.zeros(256, dtype=np.uint8)
    for ch, v in ((65, 0), (67, 1), (84, 2), (71, 3)):
        code[ch] = v
    letters = np.array([65, 67, 84, 71], dtype=np.uint8)

    def kmers_of(seq: np.ndarray) -> np.ndarray:
        c = code[seq].astype(np.uint64)
        n = c.size - k + 1
        if n <= 0:
            return np.zeros(0, dtype=np.uint64)
        key = np.zeros(n, dtype=np.uint64)
        for i in range(k):
            key = (key << np.uint64(2)) | c[i:i + n]
        pk = np.uint64(0)
        for b in prefix:
            pk = (pk << np.uint64(2)) | np.uint64(code[b])
        sel = (key >> np.uint64(2 * (k - len(prefix)))) == pk
        return np.unique(key[sel])

    names, lengths, ulens, species, sets = [], [], [], [], []
    n_rel = max(1, n_templates // 2)
    for t in range(n_templates):
        if t == 0:
            seq = g
        elif t < n_rel:
            seq = g.copy()
            rate = divergence * t
            pos = np.nonzero(rng.random(seq.size) < rate)[0]
            seq[pos] = letters[rng.integers(0, 4, pos.size)]
        else:
            seq = letters[rng.integers(0, 4, g.size)]
        ks = kmers_of(seq)
        sets.append(ks)
        names.append(f"SYN_{t:05d}")
        lengths.append(int(seq.size))
        ulens.append(int(ks.size))
        species.append("Synthetic genome" if t == 0 else (f"relative {t}" if t < n_rel else f"decoy {t}"))
    # invert: k-mer -> template list (ascending template id = DB list order)
    allk = np.concatenate(sets)
    tid = np.concatenate([np.full(s.size, i, dtype=np.uint32) for i, s in enumerate(sets)])
    order = np.lexsort((tid, allk))
    allk, tid = allk[order], tid[order]
    uniq, start = np.unique(allk, return_index=True)
    list_off = np.concatenate([start, [allk.size]]).astype(np.uint64)
    # shuffle the k-mer order so DB order is unrelated to key order
    perm = rng.permutation(uniq.size)
    lens = (list_off[1:] - list_off[:-1])[perm]
    new_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    tm = np.concatenate([tid[int(list_off[i]):int(list_off[i + 1])] for i in perm]) if uniq.size else np.zeros(0, np.uint32)
    uk = uniq[perm]
    kb = np.zeros((uk.size, k), dtype=np.uint8)
    for i in range(k):
        kb[:, i] = letters[((uk >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)).astype(np.int64)]
    summary = {"templates": n_templates, "uniqueLens": int(sum(ulens)), "totalLen": int(sum(lengths))}
    return TemplateDB(kb.reshape(-1), np.full(uk.size, k, dtype=np.uint32), new_off, tm, names,
                      np.array(lengths, dtype=np.uint64), np.array(ulens, dtype=np.uint64), species, summary)


def genus_template_db(sample_keys: np.ndarray, n_templates: int = 10_000, per_template: int = 10_000,
                      genus_size: int = 20, identity: float = 0.7, overlap: float = 0.4,
                      prefix: bytes = b"ATGAC", k: int = 16, seed: int = 11, relabel: bool = True,
                      sample_share: float = 0.9):
    """BASELINE config 4's database: ``n_templates`` templates of about ``per_template``
    prefix-filtered k-mers each, in genera of ``genus_size`` templates that share a k-mer pool (a template
    holds each k-mer of its genus pool with probability ``identity``, so two templates of a genus are
    about ``identity`` alike); neighbouring genera overlap in a fraction ``overlap`` of their pools, so
    k-mer lists mix genera.  Template 0 holds ``sample_share`` of ``sample_keys`` (2-bit packed keys, first base most
    significant: the sample genome's own k-mers -- the sample is a strain of template 0, not template 0 itself, so
    the winner-takes-all loop goes on to the relatives) and its genus the relatives of the sample.  Built in k-mer
    space, deterministic in ``seed``.  ``relabel`` permutes the template ids so that list order (ascending
    original id) is unrelated to the id.  Returns a TemplateDB."""
    from .db import TemplateDB
    rng = np.random.default_rng(seed)
    m = len(prefix)
    code = {65: 0, 67: 1, 84: 2, 71: 3}
    pk = 0
    for b in prefix:
        pk = (pk << 2) | code[b]
    sample_keys = np.unique(np.asarray(sample_keys, dtype=np.uint64))
    n_genera = (n_templates + genus_size - 1) // genus_size
    pool = max(int(round(per_template / identity)), int(sample_keys.size), 1)
    stride = max(1, int(round(pool * (1.0 - overlap))))
    sfx_bits = 2 * (k - m)
    space = 1 << sfx_bits if sfx_bits < 62 else None
    if space is not None and n_genera > 1 and stride * (n_genera - 1) + pool > space:
        stride = max(1, (space - pool) // (n_genera - 1))           # the k-mer space is small: genera overlap more
    n_pos = stride * (n_genera - 1) + pool                       # k-mers of the DB, in "position" order
    # position -> key: the sample's k-mers first, then distinct random keys with the prefix
    need = n_pos - sample_keys.size
    if space is not None and space <= (1 << 26):
        if need > space - sample_keys.size:
            raise ValueError("the k-mer space is smaller than the database asked for")
        allk = (np.uint64(pk) << np.uint64(sfx_bits)) | rng.permutation(space).astype(np.uint64)
        rest = allk[~np.isin(allk, sample_keys)][:need]
    else:
        rest = np.zeros(0, dtype=np.uint64)
        while rest.size < need:
            cand = rng.integers(0, 1 << min(sfx_bits, 62), size=2 * (need - rest.size) + 16, dtype=np.uint64)
            cand = (np.uint64(pk) << np.uint64(sfx_bits)) | cand
            cand = cand[~np.isin(cand, sample_keys)]
            rest = np.unique(np.concatenate([rest, cand]))
        rest = rng.permutation(rest)[:need]
    keys = np.concatenate([sample_keys, rest])
    # per genus: membership matrix [template in genus, pool column]; column j of genus g is position g*stride + j
    col_cnt = np.zeros((n_genera, pool), dtype=np.uint32)
    members = []
    per_t = np.zeros(n_templates, dtype=np.uint64)
    for g in range(n_genera):
        gs = min(genus_size, n_templates - g * genus_size)
        M = rng.random((gs, pool)) < identity
        if g == 0:
            M[0, :] = (np.arange(pool) < sample_keys.size) & (rng.random(pool) < sample_share)   # the sample's closest template
        cols, ts = np.nonzero(M.T)                                 # sorted by column, then template
        members.append((cols.astype(np.int64), (ts + g * genus_size).astype(np.uint32)))
        col_cnt[g] = M.sum(axis=0)
        per_t[g * genus_size:g * genus_size + gs] = M.sum(axis=1)
    # list length of every position = its column in the genus that starts at or before it + columns of
    # earlier genera that still reach it
    reach = (pool + stride - 1) // stride                          # genera covering one position (at most)
    cnt = np.zeros(n_pos, dtype=np.uint64)
    for g in range(n_genera):
        cnt[g * stride:g * stride + pool] += col_cnt[g]
    list_off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.uint64)
    tmpl = np.zeros(int(list_off[-1]), dtype=np.uint32)
    fill = np.zeros(n_pos, dtype=np.uint64)                        # entries already written per position
    for g in range(n_genera):                                      # ascending genus = ascending template id
        cols, ts = members[g]
        if cols.size == 0:
            continue
        pos = cols + g * stride
        first = np.concatenate([[0], np.nonzero(np.diff(cols))[0] + 1])
        start_of = np.repeat(first, np.diff(np.concatenate([first, [cols.size]])))
        rank = np.arange(cols.size) - start_of                     # rank inside the column
        tmpl[(list_off[pos] + fill[pos] + rank.astype(np.uint64)).astype(np.int64)] = ts
        fill[g * stride:g * stride + pool] += col_cnt[g]
    del members, reach
    if relabel:
        perm = rng.permutation(n_templates).astype(np.uint32)
        perm[perm == 0], perm[0] = perm[0], 0                      # the sample genome stays template 0
        tmpl = perm[tmpl]
        inv = np.empty_like(perm)
        inv[perm] = np.arange(n_templates, dtype=np.uint32)
        per_t = per_t[inv]
    letters = np.array([65, 67, 84, 71], dtype=np.uint8)
    kb = np.zeros((n_pos, k), dtype=np.uint8)
    for i in range(k):
        kb[:, i] = letters[((keys >> np.uint64(2 * (k - 1 - i))) & np.uint64(3)).astype(np.int64)]
    ulens = per_t + rng.integers(1, 50, size=n_templates).astype(np.uint64)
    lengths = np.uint64(1000) + np.uint64(50) * per_t
    names = [f"SYN_{t:05d}" for t in range(n_templates)]
    species = ["Synthetic sample genome" if t == 0 else f"synthetic template {t}" for t in range(n_templates)]
    summary = {"templates": n_templates, "uniqueLens": int(ulens.sum()), "totalLen": int(lengths.sum())}
    db = TemplateDB(kb.reshape(-1), np.full(n_pos, k, dtype=np.uint32), list_off, tmpl, names, lengths, ulens,
                    species, summary)
    db.keys_u64 = keys                                             # position -> 2-bit key (tests)
    return db
