"""numpy restatement of the synthetic FASTQ generator (kmerjs_b200/csrc/kj_synth.cu) -- TEST INFRASTRUCTURE.

The CPU legs of bench.py (`cpu_baseline`, `--impl reference`) need the same reads as the GPU arm without
loading libkmerjs_b200.so; tests/test_synth_ref.py pins this module byte for byte to the device generator.
Only tests/ and bench.py's CPU legs import it."""
from __future__ import annotations

import numpy as np

_M1, _M2 = np.uint64(0xBF58476D1CE4E5B9), np.uint64(0x94D049BB133111EB)
_G1, _G2, _G3 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xD1B54A32D192ED03), np.uint64(0x632BE59BD9B4E019)
HDR = 42


def _mix64(x):
    x = x ^ (x >> np.uint64(30))
    x = x * _M1
    x = x ^ (x >> np.uint64(27))
    x = x * _M2
    return x ^ (x >> np.uint64(31))


def _rng(seed, read, j):
    with np.errstate(over="ignore"):
        return _mix64(_mix64(np.uint64(seed) ^ (np.asarray(read, dtype=np.uint64) * _G1)) + np.asarray(j, dtype=np.uint64) * _G2 + _G3)


def genome(seed: int, n: int) -> np.ndarray:
    out = np.empty(n, dtype=np.uint8)
    letters = np.frombuffer(b"ACGT", dtype=np.uint8)
    for lo in range(0, n, 1 << 22):
        hi = min(n, lo + (1 << 22))
        h = _rng(seed, np.uint64(0x67656E6F6D65), np.arange(lo, hi, dtype=np.uint64))
        out[lo:hi] = letters[(h >> np.uint64(62)).astype(np.int64)]
    return out


def fastq(seed: int, n_reads: int, g: np.ndarray, read_len: int = 150, first_read: int = 0, sub_rate: float = 0.005,
          n_rate: float = 1e-4, lead_n_rate: float = 0.02) -> np.ndarray:
    """The records of reads [first_read, first_read + n_reads) as one uint8 array."""
    L = read_len
    rec = HDR + 2 * L + 4
    out = np.empty((n_reads, rec), dtype=np.uint8)

    def thr(p):
        return np.uint64(min(max(p * 16777216.0, 0.0), 16777216.0))

    sub_thr, n_thr, lead_thr = thr(sub_rate), thr(n_rate), thr(lead_n_rate)
    comp = np.frombuffer(b"TGAC", dtype=np.uint8)              # complement by code (A C T G)
    base_of = np.frombuffer(b"ACGT", dtype=np.uint8)
    cur_of = np.zeros(256, dtype=np.int64)
    for ch, v in ((65, 0), (67, 1), (71, 2), (84, 3)):
        cur_of[ch] = v
    cur_of[ord("N")] = 3
    jj = np.arange(L, dtype=np.uint64)
    m24 = np.uint64(0xFFFFFF)
    for lo in range(0, n_reads, 50000):
        hi = min(n_reads, lo + 50000)
        r = np.arange(first_read + lo, first_read + hi, dtype=np.uint64)
        o = out[lo:hi]
        h0 = _rng(seed, r, np.uint64(0xFFFFFFF0))
        h1 = _rng(seed, r, np.uint64(0xFFFFFFF1))
        start = (h0 % np.uint64(g.size - L + 1)).astype(np.int64)
        strand = (h1 & np.uint64(1)).astype(bool)
        lead_n = ((h1 >> np.uint64(8)) & m24) < lead_thr
        tile = 1101 + ((h1 >> np.uint64(32)) % np.uint64(1000)).astype(np.int64)
        x = ((h0 >> np.uint64(20)) % np.uint64(100000)).astype(np.int64)
        y = ((h0 >> np.uint64(40)) % np.uint64(100000)).astype(np.int64)
        o[:, :12] = np.frombuffer(b"@SIM:1:FC:1:", dtype=np.uint8)
        for d in range(4):
            o[:, 12 + d] = 48 + (tile // 10 ** (3 - d)) % 10
        o[:, 16] = ord(":")
        for d in range(5):
            o[:, 17 + d] = 48 + (x // 10 ** (4 - d)) % 10
            o[:, 23 + d] = 48 + (y // 10 ** (4 - d)) % 10
        o[:, 22] = ord(":")
        o[:, 28:42] = np.frombuffer(b" 1:N:0:CGATGT\n", dtype=np.uint8)
        # bases
        j = np.arange(L, dtype=np.int64)
        idx = np.where(strand[:, None], start[:, None] + (L - 1 - j)[None, :], start[:, None] + j[None, :])
        gb = g[idx]
        c = np.where(strand[:, None], comp[(gb >> 1) & 3], gb)
        h = _rng(seed, r[:, None], jj[None, :])
        sub = (h & m24) < sub_thr
        cur = cur_of[c]
        alt = base_of[(cur + 1 + ((h >> np.uint64(24)) % np.uint64(3)).astype(np.int64)) & 3]
        c = np.where(sub, alt, c)
        c = np.where(((h >> np.uint64(32)) & m24) < n_thr, np.uint8(ord("N")), c)
        c[:, 0] = np.where(lead_n, np.uint8(ord("N")), c[:, 0])
        o[:, HDR:HDR + L] = c
        o[:, HDR + L] = 10
        o[:, HDR + L + 1] = ord("+")
        o[:, HDR + L + 2] = 10
        hq = _rng(seed, r[:, None], (np.uint64(0x10000) + jj)[None, :])
        o[:, HDR + L + 3:HDR + 2 * L + 3] = (35 + (hq % np.uint64(39))).astype(np.uint8)
        o[:, HDR + 2 * L + 3] = 10
    return out.reshape(-1)
