"""Parity of the CUDA extraction + count path (through the C ABI) with the CPU oracle.
Bit-exact: key sets, counts, line counts, first-insertion order."""
import json
import os
import random

import numpy as np
import pytest

import kmer_oracle as ko_py
import ko as ko_c
from conftest import read_golden
from util import DevBuf, random_fastq

from kmerjs_b200 import _abi
from kmerjs_b200.counts import Counts

pytestmark = pytest.mark.gpu


def oracle(data, prefix=b"ATGAC", k=16, step=1):
    counts, lines = ko_c.count_fastq(data, prefix, k, step)
    return [(kk.decode("latin-1"), v) for kk, v in counts.items()], lines


def gpu(data, prefix=b"ATGAC", k=16, step=1, flags=0, device=False, pieces=None, halo=64, **kw):
    c = Counts(prefix, k, step, flags=flags | _abi.KJ_F_COUNT_BASES, **kw)
    if pieces is None:
        if device:
            b = DevBuf(data)
            c.add_device(b.ptr, b.n, final=True)
        else:
            c.add_host(data, final=True)
    else:
        # feed the stream in pieces: each non-final piece carries a halo the next one presents again
        keep = []
        lo = 0
        cuts = list(pieces) + [len(data)]
        for hi in cuts:
            final = hi == len(data)
            end = len(data) if final else min(len(data), hi + halo)
            if not final and end - hi < 32:          # ABI: non-final pieces need >= 32 halo bytes
                continue
            chunk = data[lo:end]
            if device:
                b = DevBuf(chunk)
                keep.append(b)
                c.add_device(b.ptr, b.n, own_n=hi - lo, final=final)
            else:
                c.add_host(chunk, own_n=hi - lo, final=final)
            lo = hi
    c.finish()
    out = list(c.to_dict().items()), c.lines
    stats = (c.occurrences, c.bases, c.size)
    c.free()
    return out, stats


def seq_bases(data):
    lines = ko_py.split_lines(data)
    return sum(len(l) for i, l in enumerate(lines) if i % 4 == 1)      # every sequence line, gated or not


@pytest.mark.parametrize("name", ["test_short.fastq", "test_long.kmer.fastq", "test_kmers.fastq"])
@pytest.mark.parametrize("flags", [0, _abi.KJ_F_FORCE_GENERIC])
def test_reference_fixtures(name, flags):
    data = read_golden(name)
    exp = oracle(data)
    got, (occ, bases, size) = gpu(data, flags=flags)
    assert got == exp
    assert occ == sum(v for _, v in exp[0]) and size == len(exp[0])
    assert bases == seq_bases(data)


def test_known_answers(known):
    got, _ = gpu(read_golden("test_short.fastq"))
    assert got[0] == [tuple(x) for x in known["KA3_test_short"]["map"]] and got[1] == 40
    got, _ = gpu(read_golden("test_long.kmer.fastq"), device=True)
    assert len(got[0]) == known["KA4_test_long_kmer_size"]["size"]
    golden = json.loads(read_golden("kmers_long.json"))
    assert all(k in golden and v <= golden[k] for k, v in got[0])


GRID = [(b"ATGAC", 16, 1), (b"", 31, 1), (b"ATGAC", 16, 3), (b"A", 5, 2), (b"GT", 32, 1), (b"", 1, 1),
        (b"ATGACG", 6, 1), (b"N", 3, 1), (b"ATGACATGACATGACATG", 16, 1), (b"G", 1, 1), (b"AC", 2, 5)]


@pytest.mark.parametrize("prefix,k,step", GRID)
def test_parameter_grid_small_fixture(prefix, k, step):
    data = read_golden("test_kmers.fastq")          # alphabet with X W Z E, ragged read lengths
    assert gpu(data, prefix, k, step)[0] == oracle(data, prefix, k, step)
    assert gpu(data, prefix, k, step, flags=_abi.KJ_F_FORCE_GENERIC, device=True)[0] == oracle(data, prefix, k, step)


def test_digests_long_fixture():
    dig = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "oracle_digests.json")))
    data = read_golden("test_long.kmer.fastq")
    for key in ("test_long.kmer.fastq|ATGAC|16|3", "test_long.kmer.fastq|GT|32|1", "test_long.kmer.fastq|ATGACG|6|1"):
        _, prefix, k, step = key.split("|")
        (m, lines), _ = gpu(data, prefix.encode(), int(k), int(step))
        assert (len(m), sum(v for _, v in m), lines) == (dig[key]["unique"], dig[key]["total"], dig[key]["lines"])
        assert m == oracle(data, prefix.encode(), int(k), int(step))[0]


FUZZ = [
    dict(),                                             # plain
    dict(crlf=True),                                    # '\r' stays in the line (lib/kmers.js:121)
    dict(blank_lines=0.15),                             # blank lines shift the mod-4 phase
    dict(trailing_newline=False),                       # unterminated last line is flushed (:130-136)
    dict(p_n=0.08, p_lower=0.05),                       # irregular k-mers
    dict(min_len=0, max_len=20),                        # reads shorter than k, empty and 1-char lines
    dict(alphabet=b"AAAAAAAC", plant=None),             # low complexity: heavy duplicates, palindromes
    dict(alphabet=b"ATGAC", plant=(b"ATGAC", 0.9)),     # dense candidates
]


@pytest.mark.parametrize("case", range(len(FUZZ)))
@pytest.mark.parametrize("flags", [0, _abi.KJ_F_FORCE_GENERIC])
def test_fuzz(case, flags):
    rng = random.Random(1000 + case)
    data = random_fastq(rng, 120, **FUZZ[case])
    for prefix, k, step in [(b"ATGAC", 16, 1), (b"AT", 4, 1), (b"", 9, 1), (b"CA", 7, 3)]:
        exp = oracle(data, prefix, k, step)
        got, (occ, bases, _) = gpu(data, prefix, k, step, flags=flags)
        assert got == exp, (case, prefix, k, step)
        assert occ == sum(v for _, v in exp[0])
        assert bases == seq_bases(data)


@pytest.mark.parametrize("k", [2, 5, 16, 31, 32])
def test_dense_kernel_empty_prefix(k):
    """Empty prefix, step 1, k >= 2 runs the dense kernel (every window is an emission, both strands);
    the line kernel (KJ_F_FORCE_GENERIC) and the oracle must agree with it, irregular bytes included."""
    rng = random.Random(900 + k)
    for case in (dict(p_n=0.03, p_lower=0.02), dict(crlf=True), dict(blank_lines=0.1, trailing_newline=False),
                 dict(min_len=0, max_len=40), dict(alphabet=b"AAAAAAAC", plant=None)):
        data = random_fastq(rng, 80, **case)
        exp = oracle(data, b"", k, 1)
        got, (occ, bases, _) = gpu(data, b"", k, 1)
        assert got == exp, (k, case)
        assert occ == sum(v for _, v in exp[0]) and bases == seq_bases(data)
        assert gpu(data, b"", k, 1, flags=_abi.KJ_F_FORCE_GENERIC, device=True)[0] == exp
    data = read_golden("test_kmers.fastq")
    cuts = [1000, 3000, 5555]
    assert gpu(data, b"", k, 1, pieces=cuts, halo=64)[0] == oracle(data, b"", k, 1)      # window kernel: 64-byte halo


def test_empty_and_degenerate_inputs():
    for data in (b"", b"\n", b"\n\n\n\n\n", b"@r\nA\n+\n#\n", b"@r\nATGACATGACATGACAT", b"ATGACATGACATGACATG",
                 b"@r\n\n+\n\n", b"@\nATGACATGACATGACATGAC\n+"):
        for flags in (0, _abi.KJ_F_FORCE_GENERIC):
            assert gpu(data, flags=flags)[0] == oracle(data), data
            assert gpu(data, b"", 1, 1, flags=flags)[0] == oracle(data, b"", 1, 1), data


def test_tile_boundaries():
    """Sequence lines, windows and newlines that straddle the 32 KiB tiles of the scan kernel."""
    rng = random.Random(7)
    body = random_fastq(rng, 900, min_len=60, max_len=130)
    for pad in (0, 1, 15, 16, 17, 31):
        data = b"@" + b"x" * pad + b"\n" + body[body.index(b"\n") + 1:]
        for flags in (0, _abi.KJ_F_FORCE_GENERIC):
            assert gpu(data, flags=flags, device=True)[0] == oracle(data)


def test_long_lines():
    rng = random.Random(11)
    seq = bytes(rng.choice(b"ACGT") for _ in range(100000))
    seq = seq[:40000] + b"ATGACATGACATGACATGAC" + seq[40000:]
    data = b"@long\n" + seq + b"\n+\n" + b"I" * len(seq) + b"\n"
    assert gpu(data, device=True)[0] == oracle(data)
    assert gpu(data, b"ATGAC", 16, 1, flags=_abi.KJ_F_FORCE_GENERIC)[0] == oracle(data)


@pytest.mark.parametrize("device", [False, True])
def test_streaming_pieces_equal_whole(device):
    rng = random.Random(21)
    data = random_fastq(rng, 400, p_n=0.02, blank_lines=0.05)
    exp = oracle(data)
    cuts = sorted(rng.sample(range(1, len(data) - 1), 7))
    assert gpu(data, pieces=cuts, device=device)[0] == exp
    # line kernel: the halo must cover whole lines
    assert gpu(data, b"", 12, 1, pieces=cuts, halo=400, device=device)[0] == oracle(data, b"", 12, 1)
    assert gpu(data, b"AC", 8, 2, pieces=cuts, halo=400, device=device)[0] == oracle(data, b"AC", 8, 2)


def test_sharded_base_line():
    """A rank that starts in the middle of the stream: base_line = '\\n' before it, base_col = bytes
    of the current line before it (multi-GPU ingest)."""
    rng = random.Random(5)
    data = random_fastq(rng, 300)
    whole = dict(oracle(data)[0])
    for cut in (1, len(data) // 3, len(data) // 2 + 7):
        nl = data[:cut].count(b"\n")
        col = cut - (data[:cut].rfind(b"\n") + 1)
        a = Counts(b"ATGAC", 16, 1, flags=_abi.KJ_F_COUNT_BASES)
        a.add_host(data[:cut + 64], own_n=cut, final=False)     # left rank: halo of 64 bytes, not final
        # a left rank never sees the end of stream; finish() is still valid
        a.finish()
        b = Counts(b"ATGAC", 16, 1, flags=_abi.KJ_F_COUNT_BASES, base_line=nl, base_col=col)
        b.add_host(data[cut:], final=True)
        b.finish()
        merged = {}
        for d in (a.to_dict(), b.to_dict()):
            for kk, v in d.items():
                merged[kk] = merged.get(kk, 0) + v
        assert merged == whole, cut
        assert a.bases + b.bases == seq_bases(data), cut      # the cut may fall inside a sequence line
        assert b.lines == oracle(data)[1]
        a.free(); b.free()


def test_no_order_flag_same_counts():
    data = read_golden("test_long.kmer.fastq")
    exp = dict(oracle(data)[0])
    got, _ = gpu(data, flags=_abi.KJ_F_NO_ORDER)
    assert dict(got[0]) == exp


def test_table_growth_from_tiny_hint():
    rng = random.Random(3)
    data = random_fastq(rng, 3000, min_len=100, max_len=100, plant=None)
    exp = oracle(data, b"", 11, 1)
    got, _ = gpu(data, b"", 11, 1, capacity_hint=16)
    assert got == exp
    exp = oracle(data, b"A", 11, 1)
    got, _ = gpu(data, b"A", 11, 1, capacity_hint=16, device=True)
    assert got == exp


def test_file_api_and_kmerjs_entry(tmp_path, known):
    import kmerjs_b200
    path = tmp_path / "short.fastq"
    path.write_bytes(read_golden("test_short.fastq"))
    out = tmp_path / "out.json"
    res = kmerjs_b200.kmerjs(str(path), "ATGAC", 16, 1, str(out))
    m = res.result(timeout=120)
    assert list(m.items()) == [tuple(x) for x in known["KA3_test_short"]["map"]]
    assert out.read_text() == "{\nATGACGCAATACTCCT: 1,ATGACCTGAGAGCCTT: 1,}\n"      # lib/index.js:381-388
    job = kmerjs_b200.KmerJS(str(path), progress=False)
    h = job.readFile()
    assert h.promise.result(timeout=120) == m and job.kmerMapSize == 2 and job.lines == 40
    assert job.bytesRead == len(read_golden("test_short.fastq"))


def test_kmers_in_line(known):
    import kmerjs_b200
    job = kmerjs_b200.KmerJS()
    ka = known["KA2_first_key"]
    job.kmersInLine(ka["line"])                                   # test/kmers.js:12-19 (the literal holds '\n')
    assert next(iter(job.kmerMap)) == ka["first"]
    exp = {}
    ko_py.kmers_in_line(ka["line"].encode(), exp)
    assert {k.encode(): v for k, v in job.kmerMap.items()} == exp
    job.kmersInLine(ka["line"])                                   # counts accumulate in the same Map
    assert all(job.kmerMap[k.decode()] == 2 * v for k, v in exp.items())
    one = kmerjs_b200.KmerJS("", "A", 1)
    one.kmersInLine("A")                                          # no length gate in kmersInLine itself
    assert dict(one.kmerMap) == {"A": 1}
    assert kmerjs_b200.complement(known["KA1_complement"]["in"]) == known["KA1_complement"]["out"]


def test_size_independent_properties():
    """A larger synthetic input (generated on the device): totals, strand symmetry and split
    invariance hold at sizes the oracle is not asked to follow."""
    import ctypes as C
    from kmerjs_b200 import synth
    n_reads = 4000 if os.environ.get("KMERJS_B200_EMU") == "1" else 400000
    w = synth.Workload(n_reads=n_reads, genome_len=200000, seed=99)
    c = Counts(b"ATGAC", 16, 1, flags=_abi.KJ_F_COUNT_BASES)
    c.add_device(w.fastq_ptr, w.n_bytes, final=True).finish()
    m = c.to_dict()
    assert c.lines == 4 * n_reads and c.bases == 150 * n_reads
    assert sum(m.values()) == c.occurrences
    # strand symmetry: the reverse-complement prefix job sees every occurrence from the other side
    c2 = Counts(b"ATGAC", 16, 1)
    half = (n_reads // 2) * w.record_bytes
    c2.add_device(w.fastq_ptr, w.n_bytes, own_n=half, final=False)
    c2.add_device(w.fastq_ptr + half, w.n_bytes - half, final=True)
    c2.finish()
    assert c2.to_dict() == m and c2.lines == c.lines
    # oracle on a prefix of the same reads
    sample = w.host_bytes(200)
    exp = oracle(sample)
    cs = Counts(b"ATGAC", 16, 1)
    cs.add_host(sample).finish()
    assert (list(cs.to_dict().items()), cs.lines) == exp


def test_n_dense_input_is_counted_without_limit():
    """Reads over the alphabet ACGTN (20 % N): most prefix-passing k-mers are irregular (byte-string keys), far more
    than any fixed side-table or spill-list size.  The reference counts them without a limit; so does every kernel
    here, whatever the capacity hint (the filter path retries marked entries after growing the tables, the dense
    and line kernels keep a piece's worst case of spills)."""
    from util import emulated
    rng = random.Random(77)
    data = random_fastq(rng, 120 if emulated() else 3400, min_len=150, max_len=150, alphabet=b"ACGTN", plant=None, p_n=0.0)
    for prefix, k in [(b"A", 20), (b"AT", 16)]:
        exp = oracle(data, prefix, k, 1)
        assert emulated() or len(exp[0]) > (60000 if prefix == b"A" else 15000)
        # (device + a hint that fits: the single-wait finish, whose first copy brings 8192 irregular records back)
        for kw in (dict(), dict(capacity_hint=1 << 22), dict(device=True, capacity_hint=1 << 10), dict(device=True, capacity_hint=1 << 22)):
            got, (occ, _, _) = gpu(data, prefix, k, 1, **kw)
            assert got == exp, (prefix, k, kw)
            assert occ == sum(v for _, v in exp[0])
    exp = oracle(data, b"", 31, 1)
    assert gpu(data, b"", 31, 1, capacity_hint=1 << 10)[0] == exp                               # dense kernel
    assert gpu(data, b"A", 20, 1, flags=_abi.KJ_F_FORCE_GENERIC, capacity_hint=1 << 10)[0] == oracle(data, b"A", 20, 1)


def test_irregular_kmers_sorted_on_the_device(monkeypatch):
    """Large irregular sets are put into first-seen order by a device radix sort in kj_counts_finish (config 5 has some
    10^8 of them); the threshold is lowered so that the small case takes that path: same map, same Map order."""
    from util import emulated
    rng = random.Random(78)
    data = random_fastq(rng, 60 if emulated() else 600, min_len=150, max_len=150, alphabet=b"ACGTN", plant=None, p_n=0.0)
    exp_dense, exp_filter = oracle(data, b"", 31, 1), oracle(data, b"A", 20, 1)
    monkeypatch.setenv("KJ_IRR_DEVICE_SORT_MIN", "8")
    assert gpu(data, b"", 31, 1)[0] == exp_dense
    assert gpu(data, b"A", 20, 1)[0] == exp_filter
    assert gpu(data, b"A", 20, 1, flags=_abi.KJ_F_FORCE_GENERIC)[0] == exp_filter


@pytest.mark.parametrize("extra", [1, 31, 32, 33, 100])
def test_file_tail_shorter_than_the_halo(tmp_path, ctx, extra):
    """readFile() on a file whose last bytes fall just past a staging-chunk boundary: a tail shorter than the
    32-byte halo belongs to the piece before it (1 MiB staging chunks, file = 2 chunks + extra bytes)."""
    from util import emulated
    chunk = (1 << 18) if emulated() else (1 << 20)
    ctx.set_stage_chunk(chunk)
    try:
        rng = random.Random(5000 + extra)
        size = 2 * chunk + extra
        data = random_fastq(rng, size // 250 + 50, min_len=100, max_len=140)[:size]
        assert len(data) == size
        path = tmp_path / "tail.fastq"
        path.write_bytes(data)
        c = Counts(b"ATGAC", 16, 1)
        c.add_file(str(path)).finish()
        assert (list(c.to_dict().items()), c.lines) == oracle(data)
        assert c.bytes_read == size
        c.free()
        c = Counts(b"ATGAC", 16, 1)                  # the same bytes as a pageable host buffer
        c.add_host(data).finish()
        assert (list(c.to_dict().items()), c.lines) == oracle(data)
        c.free()
    finally:
        ctx.set_stage_chunk(64 << 20)


@pytest.mark.parametrize("extra", [0, 1, 31, 33, 64, 65, 5000])
def test_gzip_file_is_inflated_into_the_staging_buffers(tmp_path, ctx, extra):
    """readFile() on a gzip'd FASTQ (kj_counts_add_file sees the magic and inflates on the reader thread): several staging
    pieces, the end of the stream falling at / just behind / inside the halo of a piece boundary, two gzip members in one
    file; the same map, Map order, line and byte counts as the oracle on the plain bytes.  A truncated stream is an error."""
    import gzip
    from util import emulated
    chunk = (1 << 18) if emulated() else (1 << 20)
    ctx.set_stage_chunk(chunk)
    try:
        rng = random.Random(7000 + extra)
        size = 2 * chunk + extra
        data = random_fastq(rng, size // 250 + 50, min_len=100, max_len=140)[:size]
        assert len(data) == size
        path = tmp_path / "reads.fastq.gz"
        cut = size // 3
        path.write_bytes(gzip.compress(data[:cut], 1) + gzip.compress(data[cut:], 6))       # two members
        for prefix, k in ((b"ATGAC", 16), (b"", 31)) if extra in (0, 33) else ((b"ATGAC", 16),):
            c = Counts(prefix, k, 1)
            c.add_file(str(path)).finish()
            assert (list(c.to_dict().items()), c.lines) == oracle(data, prefix, k, 1)
            assert c.bytes_read == size
            c.free()
        if extra == 5000:
            blob = path.read_bytes()
            bad = tmp_path / "truncated.fastq.gz"
            bad.write_bytes(blob[: len(blob) // 2])
            c = Counts(b"ATGAC", 16, 1)
            with pytest.raises(_abi.KjError):
                c.add_file(str(bad))
            c.free()
            empty = tmp_path / "empty.fastq.gz"
            empty.write_bytes(gzip.compress(b""))
            c = Counts(b"ATGAC", 16, 1)
            c.add_file(str(empty)).finish()
            assert c.size == 0 and c.lines == 0
            c.free()
    finally:
        ctx.set_stage_chunk(64 << 20)
