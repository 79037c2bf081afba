"""CPU oracle for the kmerjs hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-Python restatement of the reference algorithm (josl/kmerjs), written from
its behaviour, for checking the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import this module; nothing under
``kmerjs_b200/`` does.

Reference lines followed (paths relative to the kmerjs repository root):

* ``complement``            lib/kmers.js:12-17,31-38
* ``kmers_in_line``         lib/kmers.js:88-100      (incl. the ``step>1`` short-window quirk)
* ``split_lines``           lib/kmers.js:114-136     (split on '\\n', carry, flush non-empty tail)
* ``count_fastq``           lib/kmers.js:143-178     (4-line FSM, ``length > 1`` gate, kmerMapSize)
* ``first_match``           lib/kmerFinderServer.js:171-226
* ``get_matches``           lib/kmerFinderClient.js:232-271
* ``find_winner`` / gate    lib/kmerFinderClient.js:100-109,179-218
* ``match_summary``         lib/kmerFinderClient.js:41-92
* ``find_matches`` (loop)   lib/kmerFinderClient.js:273-289
* ``z_score`` / ``fastp``   lib/stats.js:6,19-45,52-115

Third-party arithmetic: the reference computes its statistics with ``bignumber.js ^2.3.0``
(package.json:61; not vendored in the reference tree).  ``BN`` below restates the subset
used: exact plus/minus/times, ``dividedBy`` and ``sqrt`` correctly rounded to
DECIMAL_PLACES=20 with ROUNDING_MODE (default 4 = ROUND_HALF_UP; 2 = ROUND_CEIL once
lib/kmerFinderServer.js:7 ran), ``round(dp, rm)`` and ``toNumber``.

Parity pinning: checked in tests/test_oracle_golden.py against every known answer the
reference's tests/fixtures hold for this path.
"""
from __future__ import annotations

import math
from collections import OrderedDict

# --------------------------------------------------------------------------------------
# extraction + count
# --------------------------------------------------------------------------------------

_COMP = bytes.maketrans(b"ATGC", b"TACG")  # lib/kmers.js:12-17 -- upper-case ACGT only


def complement(s: bytes) -> bytes:
    """Reverse-complement; bytes outside upper-case ACGT pass through (lib/kmers.js:31-38)."""
    return s.translate(_COMP)[::-1]


def kmers_in_line(line: bytes, counts: dict, k: int = 16, step: int = 1,
                  prefix: bytes = b"ATGAC") -> None:
    """lib/kmers.js:88-100.  ``counts`` is an insertion-ordered dict (JS Map)."""
    ini = 0
    end = k
    stop = len(line) - k
    index = 0
    while index <= stop:
        kmer = line[ini:end]          # String#substring clips to the line length
        if kmer.startswith(prefix):
            counts[kmer] = counts.get(kmer, 0) + 1
        ini += step
        end = ini + k
        index += 1


def split_lines(data: bytes) -> list:
    """lib/kmers.js:114-136: every '\\n'-terminated piece (empty ones too), plus a
    non-empty unterminated tail."""
    parts = data.split(b"\n")
    tail = parts.pop()
    if tail:
        parts.append(tail)
    return parts


def count_fastq(data: bytes, prefix: bytes = b"ATGAC", k: int = 16, step: int = 1):
    """lib/kmers.js:143-178.  Returns (counts, n_lines); ``len(counts)`` is kmerMapSize."""
    counts: dict = {}
    i = 0
    lines = 0
    for line in split_lines(data):
        if i == 1 and len(line) > 1:
            kmers_in_line(line, counts, k, step, prefix)
            kmers_in_line(complement(line), counts, k, step, prefix)
        elif i == 3:
            i = -1
        i += 1
        lines += 1
    return counts, lines


def output_file_text(counts: dict) -> str:
    """The ``out`` pseudo-JSON written by lib/index.js:381-388."""
    return "{\n" + "".join(f"{k.decode('latin-1')}: {v}," for k, v in counts.items()) + "}\n"


# --------------------------------------------------------------------------------------
# bignumber.js subset (exact decimal)
# --------------------------------------------------------------------------------------

ROUND_UP, ROUND_DOWN, ROUND_CEIL, ROUND_FLOOR = 0, 1, 2, 3
ROUND_HALF_UP, ROUND_HALF_DOWN, ROUND_HALF_EVEN = 4, 5, 6
DECIMAL_PLACES = 20


class BNConfig:
    rounding_mode = ROUND_HALF_UP    # bignumber.js default; lib/kmerFinderServer.js:7 sets 2


def _round_div(num: int, den: int, rm: int) -> int:
    """round(num/den) to an integer under bignumber.js rounding mode ``rm`` (den > 0)."""
    neg = num < 0
    a = -num if neg else num
    q, r = divmod(a, den)
    if r:
        twice = 2 * r
        if rm == ROUND_UP:
            inc = True
        elif rm == ROUND_DOWN:
            inc = False
        elif rm == ROUND_CEIL:
            inc = not neg
        elif rm == ROUND_FLOOR:
            inc = neg
        elif rm == ROUND_HALF_UP:
            inc = twice >= den
        elif rm == ROUND_HALF_DOWN:
            inc = twice > den
        elif rm == ROUND_HALF_EVEN:
            inc = twice > den or (twice == den and (q & 1))
        else:
            raise ValueError(rm)
        if inc:
            q += 1
    return -q if neg else q


class BN:
    """value = n / 10**e, exact."""
    __slots__ = ("n", "e")

    def __init__(self, v, e: int | None = None):
        if e is not None:
            self.n, self.e = int(v), e
        elif isinstance(v, BN):
            self.n, self.e = v.n, v.e
        elif isinstance(v, int):
            self.n, self.e = v, 0
        elif isinstance(v, float):
            # JS Number -> shortest round-trip string -> exact decimal (what new BN(number) sees)
            from decimal import Decimal
            d = Decimal(repr(v))
            sign, digits, exp = d.as_tuple()
            if len(str(int("".join(map(str, digits))))) > 15:
                raise ValueError("new BigNumber() number type has more than 15 significant digits")
            n = int("".join(map(str, digits)))
            if sign:
                n = -n
            if exp >= 0:
                self.n, self.e = n * 10 ** exp, 0
            else:
                self.n, self.e = n, -exp
        else:
            raise TypeError(type(v))
        self._norm()

    def _norm(self):
        while self.e > 0 and self.n % 10 == 0:
            self.n //= 10
            self.e -= 1

    @staticmethod
    def _align(a: "BN", b: "BN"):
        e = max(a.e, b.e)
        return a.n * 10 ** (e - a.e), b.n * 10 ** (e - b.e), e

    def plus(self, o):
        o = BN(o)
        x, y, e = BN._align(self, o)
        return BN(x + y, e)

    def minus(self, o):
        o = BN(o)
        x, y, e = BN._align(self, o)
        return BN(x - y, e)

    def times(self, o):
        o = BN(o)
        return BN(self.n * o.n, self.e + o.e)

    def dividedBy(self, o):
        o = BN(o)
        if o.n == 0:
            raise ZeroDivisionError("BN division by zero")
        # (n1/10^e1)/(n2/10^e2) * 10^DP, rounded
        num = self.n * 10 ** (o.e + DECIMAL_PLACES)
        den = o.n * 10 ** self.e
        if den < 0:
            num, den = -num, -den
        return BN(_round_div(num, den, BNConfig.rounding_mode), DECIMAL_PLACES)

    def sqrt(self):
        if self.n < 0:
            raise ValueError("sqrt of negative")
        rm = BNConfig.rounding_mode
        # X = value * 10^(2*DP); result = round(sqrt(X)) / 10^DP
        num = self.n * 10 ** (2 * DECIMAL_PLACES)
        den = 10 ** self.e
        fl = math.isqrt(num // den)
        exact = (fl * fl * den == num)
        if exact:
            r = fl
        elif rm in (ROUND_UP, ROUND_CEIL):
            r = fl + 1
        elif rm in (ROUND_DOWN, ROUND_FLOOR):
            r = fl
        else:  # nearest; ties impossible for a non-square rational with this scaling
            r = (math.isqrt((4 * num) // den) + 1) // 2
        return BN(r, DECIMAL_PLACES)

    def round(self, dp: int = 0, rm: int | None = None):
        if rm is None:
            rm = BNConfig.rounding_mode
        if self.e <= dp:
            return BN(self)
        return BN(_round_div(self.n, 10 ** (self.e - dp), rm), dp)

    def cmp(self, o) -> int:
        o = BN(o)
        x, y, _ = BN._align(self, o)
        return (x > y) - (x < y)

    comparedTo = cmp

    def toNumber(self) -> float:
        from decimal import Decimal
        return float(Decimal(self.n).scaleb(-self.e))

    def __repr__(self):
        from decimal import Decimal
        return f"BN({Decimal(self.n).scaleb(-self.e)})"


ETTA = BN(1.0e-8)          # lib/stats.js:6

_FASTP_TABLE = [            # lib/stats.js:56-112 (strict '>' on each threshold, in this order)
    (10.7016, 1e-26), (10.4862, 1e-25), (10.2663, 1e-24), (10.0416, 1e-23), (9.81197, 1e-22),
    (9.5769, 1e-21), (9.33604, 1e-20), (9.08895, 1e-19), (8.83511, 1e-18), (8.57394, 1e-17),
    (8.30479, 1e-16), (8.02686, 1e-15), (7.73926, 1e-14), (7.4409, 1e-13), (7.13051, 1e-12),
    (6.8065, 1e-11), (6.46695, 1e-10), (6.10941, 1e-9), (5.73073, 1e-8), (5.32672, 1e-7),
    (4.89164, 1e-6), (4.41717, 1e-5), (3.89059, 1e-4), (3.29053, 1e-3), (2.57583, 0.01),
    (1.95996, 0.05), (1.64485, 0.1),
]


def z_score(r1: int, n1: int, r2: int, n2: int) -> BN:
    """lib/stats.js:19-45."""
    p1 = BN(r1).dividedBy(n1).plus(ETTA)
    p2 = BN(r2).dividedBy(n2).plus(ETTA)
    p = BN(r1).plus(r2).dividedBy(BN(n1).plus(n2).plus(ETTA))
    q = BN(1).minus(p)
    square = BN(p).times(q).times(
        BN(1).dividedBy(BN(n1).plus(ETTA)).plus(BN(1).dividedBy(BN(n2).plus(ETTA)))
    ).plus(ETTA).sqrt()
    return BN(p1).minus(p2).dividedBy(square)


def fastp(z: BN) -> BN:
    """lib/stats.js:52-115."""
    for thr, p in _FASTP_TABLE:
        if z.cmp(BN(thr)) > 0:
            return BN(p)
    return BN(1.0)


# --------------------------------------------------------------------------------------
# template scoring
# --------------------------------------------------------------------------------------

class TemplateDB:
    """k-mer -> ordered template list (the Redis LRANGE lists of lib/kmerFinderServer.js:184-199)
    plus per-template attributes and the Summary record (lib/kmerFinderServer.js:716-724)."""

    def __init__(self, kmer_lists: dict, attrs: dict, summary: dict):
        self.kmer_lists = kmer_lists      # {kmer(bytes): [template_name, ...]}  (DB order)
        self.attrs = attrs                # {name: {"lengths": int, "ulength": int, "species": str}}
        self.summary = summary            # {"templates": int, "uniqueLens": int, "totalLen": int}


def first_match(qmap: dict, db: TemplateDB):
    """lib/kmerFinderServer.js:171-226.  Returns (templates OrderedDict, hits)."""
    templates: "OrderedDict[str, dict]" = OrderedDict()
    n_hits = 0
    for kmer, cov in qmap.items():                    # Map insertion order (:175)
        lst = db.kmer_lists.get(kmer, ())
        n_hits += len(lst)                            # :182
        for name in lst:                              # DB list order (:184)
            t = templates.get(name)
            if t is not None:
                t["tScore"] += cov
                t["uScore"] += 1
                t["kmers"][kmer] = None
            else:
                a = db.attrs[name]
                templates[name] = {"tScore": cov, "uScore": 1, "lengths": a["lengths"],
                                   "ulength": a["ulength"], "species": a["species"],
                                   "kmers": OrderedDict([(kmer, None)])}
    if n_hits == 0:
        raise RuntimeError("No hits were found!")
    return templates, n_hits


def get_matches(first_matches: "OrderedDict[str, dict]", qmap: dict):
    """lib/kmerFinderClient.js:232-271 (mutates first_matches: drops dead templates)."""
    templates: "OrderedDict[str, dict]" = OrderedDict()
    n_hits = 0
    for name in list(first_matches.keys()):
        hit = first_matches[name]
        t = None
        for kmer in hit["kmers"]:
            if kmer in qmap:
                cov = qmap[kmer]
                if t is not None:
                    t["tScore"] += cov
                    t["uScore"] += 1
                    t["kmers"][kmer] = None
                else:
                    t = {"tScore": cov, "uScore": 1, "lengths": hit["lengths"],
                         "ulength": hit["ulength"], "species": hit["species"],
                         "kmers": OrderedDict([(kmer, None)])}
                    templates[name] = t
        if t is not None:
            n_hits += len(t["kmers"])
        else:
            del first_matches[name]
    if n_hits == 0:
        raise RuntimeError("No hits were found! (nHits === 0)")
    return templates, n_hits


ROW_KEYS = ["template", "score", "expected", "z", "probability", "frac-q", "frac-d", "depth",
            "kmers-template", "total-frac-q", "total-frac-d", "total-temp-cover", "species"]


def match_summary(kmer_map_size: int, first_matches, sequence: str, match: dict, hits: int,
                  summary: dict, evalue: BN = BN(0.05)):
    """lib/kmerFinderClient.js:41-92.  Returns the 13-field OrderedDict or None."""
    seq_hit = first_matches[sequence]
    orig_u, orig_t = seq_hit["uScore"], seq_hit["tScore"]
    u = match["uScore"]
    if not u > 0:
        return None
    z = z_score(u, match["ulength"], hits, summary["uniqueLens"])
    probability = fastp(z).times(summary["templates"])
    if evalue.cmp(probability) < 0:
        return None
    qden = BN(kmer_map_size).plus(ETTA)
    dden = BN(match["ulength"]).plus(ETTA)
    frac_q = BN(100).times(2).times(u).dividedBy(qden)
    frac_d = BN(100).times(u).dividedBy(dden)
    tot_frac_q = BN(100).times(2).times(orig_u).dividedBy(qden)
    tot_frac_d = BN(100).times(orig_u).dividedBy(dden)
    tot_frac_cov = BN(orig_t).dividedBy(match["lengths"]).round(2, ROUND_HALF_EVEN).toNumber()
    expected = BN(hits).times(match["ulength"]).dividedBy(summary["uniqueLens"])
    return OrderedDict([
        ("template", sequence),
        ("score", u),
        ("expected", expected.round(0, ROUND_HALF_EVEN).toNumber()),
        ("z", z.round(2).toNumber()),
        ("probability", probability.toNumber()),
        ("frac-q", frac_q.round(2, ROUND_HALF_EVEN).toNumber()),
        ("frac-d", frac_d.round(2, ROUND_HALF_EVEN).toNumber()),
        ("depth", BN(match["tScore"]).dividedBy(match["lengths"]).round(2, ROUND_HALF_EVEN).toNumber()),
        ("kmers-template", match["ulength"]),
        ("total-frac-q", tot_frac_q.round(2, ROUND_HALF_EVEN).toNumber()),
        ("total-frac-d", tot_frac_d.round(2, ROUND_HALF_EVEN).toNumber()),
        ("total-temp-cover", tot_frac_cov),
        ("species", match["species"]),
    ])


def find_matches(first_templates, summary: dict, qmap: dict, kmer_map_size: int,
                 max_hits: int = 100):
    """lib/kmerFinderClient.js:174-290 as a generator of row dicts.  ``qmap`` is mutated
    (winner k-mers are deleted, :220-230).  Ties on uScore: stable sort = earliest template."""
    first_matches = first_templates
    hit_counter = 0
    not_found = True
    evalue = BN(0.05)
    while not_found and hit_counter < max_hits:
        templates, hits = get_matches(first_matches, qmap)
        ordered = sorted(templates.items(), key=lambda kv: -kv[1]["uScore"])   # stable
        if hit_counter == 0:
            first_matches = templates                       # :182-184
        sequence, match = ordered[0]
        row = match_summary(kmer_map_size, first_matches, sequence, match, hits, summary, evalue)
        if row is not None and evalue.cmp(BN(row["probability"])) >= 0:
            hit_counter += 1
            for kmer in match["kmers"]:
                qmap.pop(kmer, None)
            yield row
        else:
            not_found = False
    if hit_counter == 0:
        raise RuntimeError("No hits were found! (kmerResults.length === 0)")
