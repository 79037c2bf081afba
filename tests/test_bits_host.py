"""The SIMD-in-register primitives of kmerjs_b200/csrc/kj_bits.cuh, host build (g++), checked exhaustively per byte value and
byte position against byte-at-a-time definitions -- before any GPU time is spent.  The device build of the same header swaps a
few of them for PRMT / IDP.4A / BREV / SHF forms of the same functions; those run in the GPU parity tests."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include <stdint.h>
#include "kj_bits.cuh"
extern "C" {
void b_pack4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_pack4(w[i]); }
void b_nl4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_nl4(w[i]); }
void b_nl_flags4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_nl_flags4(w[i]); }
void b_nl_msb4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_nl_msb4(w[i]); }
void b_bad4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_nz4(kj_not_acgt4(w[i])); }
void b_nz4(const uint32_t *w, uint64_t n, uint32_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_nz4(w[i]); }
void b_16(const uint32_t *w, uint64_t n, uint32_t *pack, uint32_t *nl, uint32_t *bad) {
    for (uint64_t i = 0; i < n; ++i) {
        pack[i] = kj_pack16(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        nl[i] = kj_nl16(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        bad[i] = kj_bad16(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
}
void b_pairrev(const uint64_t *x, uint64_t n, uint64_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_pairrev64(x[i]); }
void b_mix(const uint64_t *x, uint64_t n, uint64_t *o) { for (uint64_t i = 0; i < n; ++i) o[i] = kj_mix64(x[i]); }
uint32_t b_funnel(uint32_t lo, uint32_t hi, uint32_t s) { return kj_funnel_r(lo, hi, s); }
uint32_t b_lanes(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t d) { return kj_lanes(c0, c1, c2, d); }
uint32_t b_zero_lanes(uint32_t a) { return kj_zero_lanes(a); }
uint64_t b_ordinal(uint64_t r, uint32_t s, uint64_t p) { return kj_ordinal(r, s, p); }
uint32_t b_code(uint32_t b) { return kj_code(b); }
uint32_t b_comp(uint32_t b) { return kj_comp_byte((uint8_t)b); }
int b_is_acgt(uint32_t b) { return kj_is_acgt(b) ? 1 : 0; }
}
'''


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    d = tmp_path_factory.mktemp("bits")
    src, so = d / "bits.cpp", d / "bits.so"
    src.write_text(SRC)
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "kmerjs_b200", "csrc"),
                    "-o", str(so), str(src)], check=True)
    L = C.CDLL(str(so))
    for name in ("b_funnel", "b_lanes", "b_zero_lanes", "b_code", "b_comp"):
        getattr(L, name).restype = C.c_uint32
    L.b_ordinal.restype = C.c_uint64
    L.b_ordinal.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64]
    return L


def _words():
    """every byte value in every byte position, the other three bytes random; plus the all-equal words"""
    rng = random.Random(1)
    out = []
    for pos in range(4):
        for b in range(256):
            for _ in range(3):
                bs = [rng.randrange(256) for _ in range(4)]
                bs[pos] = b
                out.append(bs)
    for b in range(256):
        out.append([b] * 4)
    for special in (b"ACGT", b"\n\n\n\n", b"acgt", b"NNNN", b"A\nC\n", b"\x00\x00\x00\x00", b"\xff\xff\xff\xff", b"\x0a\x8a\x0b\x1a"):
        out.append(list(special))
    return np.array(out, dtype=np.uint8)


def _call(L, name, arr_u32):
    out = np.zeros(len(arr_u32), dtype=np.uint32)
    getattr(L, name)(arr_u32.ctypes.data_as(C.c_void_p), C.c_uint64(len(arr_u32)), out.ctypes.data_as(C.c_void_p))
    return out


def test_four_byte_primitives_exhaustively(lib):
    by = _words()
    w = by.copy().view("<u4").reshape(-1)
    exp_pack = sum((((by[:, i].astype(np.uint32) >> 1) & 3) << (2 * i)) for i in range(4))
    exp_nl = sum(((by[:, i] == 0x0A).astype(np.uint32) << i) for i in range(4))
    exp_msb = sum(((by[:, i] == 0x0A).astype(np.uint32) << (8 * i + 7)) for i in range(4))
    acgt = np.isin(by, np.frombuffer(b"ACGT", dtype=np.uint8))
    exp_bad = sum(((~acgt[:, i]).astype(np.uint32) << i) for i in range(4))
    exp_nz = sum(((by[:, i] != 0).astype(np.uint32) << i) for i in range(4))
    assert (_call(lib, "b_pack4", w) == exp_pack).all()
    assert (_call(lib, "b_nl4", w) == exp_nl).all()
    assert (_call(lib, "b_nl_flags4", w) == exp_msb).all()
    assert (_call(lib, "b_nl_msb4", w) == exp_msb).all()
    assert (_call(lib, "b_bad4", w) == exp_bad).all()
    assert (_call(lib, "b_nz4", w) == exp_nz).all()


def test_sixteen_byte_primitives(lib):
    rng = np.random.default_rng(2)
    alphabet = np.frombuffer(b"ACGTNacgt\n\r@+!I5F~\x00\xff", dtype=np.uint8)
    by = alphabet[rng.integers(0, len(alphabet), size=(20000, 16))]
    by[:256, 5] = np.arange(256, dtype=np.uint8)              # every value at an odd position too
    w = by.copy().view("<u4").reshape(-1)
    n = len(by)
    pack, nl, bad = (np.zeros(n, dtype=np.uint32) for _ in range(3))
    lib.b_16(w.ctypes.data_as(C.c_void_p), C.c_uint64(n), pack.ctypes.data_as(C.c_void_p), nl.ctypes.data_as(C.c_void_p),
             bad.ctypes.data_as(C.c_void_p))
    exp_pack = sum((((by[:, p].astype(np.uint32) >> 1) & 3) << (2 * p)) for p in range(16))
    exp_nl = sum(((by[:, p] == 0x0A).astype(np.uint32) << p) for p in range(16))
    acgt = np.isin(by, np.frombuffer(b"ACGT", dtype=np.uint8))
    exp_bad = sum(((~acgt[:, p]).astype(np.uint32) << p) for p in range(16))
    assert (pack == exp_pack).all() and (nl == exp_nl).all() and (bad == exp_bad).all()


def test_word_primitives(lib):
    rng = random.Random(3)
    xs = np.array([rng.getrandbits(64) for _ in range(5000)] + [0, 1, (1 << 64) - 1, 0xAAAAAAAAAAAAAAAA, 3], dtype=np.uint64)
    out = np.zeros_like(xs)
    lib.b_pairrev(xs.ctypes.data_as(C.c_void_p), C.c_uint64(len(xs)), out.ctypes.data_as(C.c_void_p))
    for x, r in zip(xs.tolist(), out.tolist()):
        exp = 0
        for f in range(32):
            exp |= ((x >> (2 * f)) & 3) << (2 * (31 - f))
        assert r == exp
    lib.b_mix(xs.ctypes.data_as(C.c_void_p), C.c_uint64(len(xs)), out.ctypes.data_as(C.c_void_p))
    M = (1 << 64) - 1
    for x, r in zip(xs.tolist(), out.tolist()):
        x ^= x >> 30; x = x * 0xBF58476D1CE4E5B9 & M
        x ^= x >> 27; x = x * 0x94D049BB133111EB & M
        x ^= x >> 31
        assert r == x
    for _ in range(3000):
        lo, hi, c2 = rng.getrandbits(32), rng.getrandbits(32), rng.getrandbits(32)
        s = rng.randrange(32)
        assert lib.b_funnel(lo, hi, s) == (((hi << 32) | lo) >> s) & 0xFFFFFFFF
        d = rng.randrange(32)
        big = (c2 << 64) | (hi << 32) | lo
        assert lib.b_lanes(lo, hi, c2, d) == (big >> (2 * d)) & 0xFFFFFFFF
        acc = rng.getrandbits(32) & rng.getrandbits(32)
        assert lib.b_zero_lanes(acc) == sum(1 << (2 * p) for p in range(16) if (acc >> (2 * p)) & 3 == 0)
    assert lib.b_ordinal(5, 1, 7) == (5 << 28) | (1 << 27) | 7
    assert lib.b_ordinal((1 << 36) - 1, 0, (1 << 27) - 1) == (((1 << 36) - 1) << 28) | ((1 << 27) - 1)
    for b in range(256):
        assert lib.b_code(b) == (b >> 1) & 3
        assert lib.b_comp(b) == {65: 84, 84: 65, 71: 67, 67: 71}.get(b, b)          # lib/kmers.js:12-17
        assert lib.b_is_acgt(b) == (1 if b in b"ACGT" else 0)
    # codes: A=0 C=1 T=2 G=3, complement = code ^ 2, both cases alike
    assert [lib.b_code(c) for c in b"ACTGactg"] == [0, 1, 2, 3, 0, 1, 2, 3]
