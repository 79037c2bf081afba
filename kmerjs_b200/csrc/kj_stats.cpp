// kj_stats.cpp -- exact-decimal statistics of the scoring path (host side, <= 100 rows per job).
//
// The reference computes zScore / fastp (lib/stats.js:19-45,52-115) and the 13 row fields
// (lib/kmerFinderClient.js:41-92) with bignumber.js ^2.3.0 (package.json:61; third party, not in
// the reference tree).  Its published semantics for the operations used here: plus/minus/times
// are exact; dividedBy and sqrt are correctly rounded to DECIMAL_PLACES = 20 with ROUNDING_MODE
// (default 4 = ROUND_HALF_UP; lib/kmerFinderServer.js:7 sets 2 = ROUND_CEIL); round(dp, rm);
// toNumber() = the double nearest to the decimal.  The GPU produces the integers (uScore, tScore,
// hits) and a double-precision z for the gate; the reported row is finished here so that the
// rounded fields are identical to the reference's, not merely close.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <cmath>
#include <string>
#include <vector>
#include "kj_stats.hpp"

namespace kjstats {

// ------------------------------------------------------------------ unsigned big integer
// fixed-capacity limb array with the few std::vector members the arithmetic below uses: the rows
// are finished on the critical path of kj_wta_next, so no heap traffic here.  1024 bits are far
// more than the widest intermediate (a 60-decimal product scaled by 10^40: ~335 bits).
struct Limbs {
    enum { CAP = 40 };
    uint32_t v[CAP];
    uint32_t n = 0;
    size_t size() const { return n; }
    bool empty() const { return n == 0; }
    uint32_t &operator[](size_t i) { return v[i]; }
    const uint32_t &operator[](size_t i) const { return v[i]; }
    uint32_t &back() { return v[n - 1]; }
    const uint32_t &back() const { return v[n - 1]; }
    void push_back(uint32_t x) { if (n >= CAP) abort(); v[n++] = x; }
    void pop_back() { --n; }
    void resize(size_t m) { if (m > CAP) abort(); for (size_t i = n; i < m; ++i) v[i] = 0; n = (uint32_t)m; }
    void assign(size_t m, uint32_t x) { if (m > CAP) abort(); for (size_t i = 0; i < m; ++i) v[i] = x; n = (uint32_t)m; }
};

struct Big {
    Limbs d;   // little endian, no leading zero limbs; empty = 0
    Big() {}
    Big(uint64_t v) { while (v) { d.push_back((uint32_t)v); v >>= 32; } }
    bool zero() const { return d.empty(); }
    void trim() { while (!d.empty() && d.back() == 0) d.pop_back(); }
};

static int cmp(const Big &a, const Big &b) {
    if (a.d.size() != b.d.size()) return a.d.size() < b.d.size() ? -1 : 1;
    for (size_t i = a.d.size(); i-- > 0;)
        if (a.d[i] != b.d[i]) return a.d[i] < b.d[i] ? -1 : 1;
    return 0;
}
static Big add(const Big &a, const Big &b) {
    Big r;
    uint64_t carry = 0;
    size_t n = std::max(a.d.size(), b.d.size());
    r.d.resize(n);
    for (size_t i = 0; i < n; ++i) {
        uint64_t s = carry + (i < a.d.size() ? a.d[i] : 0) + (i < b.d.size() ? b.d[i] : 0);
        r.d[i] = (uint32_t)s;
        carry = s >> 32;
    }
    if (carry) r.d.push_back((uint32_t)carry);
    return r;
}
static Big sub(const Big &a, const Big &b) {   // a >= b
    Big r;
    r.d.resize(a.d.size());
    int64_t borrow = 0;
    for (size_t i = 0; i < a.d.size(); ++i) {
        int64_t s = (int64_t)a.d[i] - (i < b.d.size() ? b.d[i] : 0) - borrow;
        borrow = s < 0;
        if (s < 0) s += (int64_t)1 << 32;
        r.d[i] = (uint32_t)s;
    }
    r.trim();
    return r;
}
static Big mul(const Big &a, const Big &b) {
    Big r;
    if (a.zero() || b.zero()) return r;
    r.d.assign(a.d.size() + b.d.size(), 0);
    for (size_t i = 0; i < a.d.size(); ++i) {
        uint64_t carry = 0;
        for (size_t j = 0; j < b.d.size(); ++j) {
            uint64_t t = (uint64_t)a.d[i] * b.d[j] + r.d[i + j] + carry;
            r.d[i + j] = (uint32_t)t;
            carry = t >> 32;
        }
        size_t k = i + b.d.size();
        while (carry) {
            uint64_t t = (uint64_t)r.d[k] + carry;
            r.d[k] = (uint32_t)t;
            carry = t >> 32;
            ++k;
        }
    }
    r.trim();
    return r;
}
static Big mul_small(const Big &a, uint32_t m) {
    Big r;
    if (a.zero() || !m) return r;
    r.d.resize(a.d.size());
    uint64_t carry = 0;
    for (size_t i = 0; i < a.d.size(); ++i) {
        uint64_t t = (uint64_t)a.d[i] * m + carry;
        r.d[i] = (uint32_t)t;
        carry = t >> 32;
    }
    if (carry) r.d.push_back((uint32_t)carry);
    return r;
}
static uint32_t divmod_small(Big &a, uint32_t m) {   // a /= m, returns remainder
    uint64_t rem = 0;
    for (size_t i = a.d.size(); i-- > 0;) {
        uint64_t cur = (rem << 32) | a.d[i];
        a.d[i] = (uint32_t)(cur / m);
        rem = cur % m;
    }
    a.trim();
    return (uint32_t)rem;
}
static size_t bits(const Big &a) {
    if (a.zero()) return 0;
    return 32 * (a.d.size() - 1) + (32 - __builtin_clz(a.d.back()));
}
static bool bit(const Big &a, size_t i) {
    size_t w = i >> 5;
    return w < a.d.size() && ((a.d[w] >> (i & 31)) & 1u);
}
static void shl1_or(Big &a, bool b) {   // a = a*2 + b
    uint32_t carry = b ? 1u : 0u;
    for (size_t i = 0; i < a.d.size(); ++i) {
        uint32_t nc = a.d[i] >> 31;
        a.d[i] = (a.d[i] << 1) | carry;
        carry = nc;
    }
    if (carry) a.d.push_back(carry);
}
// q = a / b, r = a % b  (b != 0): schoolbook long division in base 2^32 (Knuth, algorithm D)
static void divmod(const Big &a, const Big &b, Big &q, Big &r) {
    q = Big();
    r = Big();
    const size_t n = b.d.size(), la = a.d.size();
    if (la < n) { r = a; return; }
    if (n == 1) {
        q = a;
        uint32_t rem = divmod_small(q, b.d[0]);
        if (rem) r.d.push_back(rem);
        return;
    }
    const int sh = __builtin_clz(b.d[n - 1]);
    // normalised copies: v = b << sh (n limbs), u = a << sh (la + 1 limbs)
    uint32_t v[Limbs::CAP], u[Limbs::CAP + 1];
    for (size_t i = n; i-- > 0;)
        v[i] = sh ? (b.d[i] << sh) | (i ? b.d[i - 1] >> (32 - sh) : 0) : b.d[i];
    u[la] = sh ? a.d[la - 1] >> (32 - sh) : 0;
    for (size_t i = la; i-- > 0;)
        u[i] = sh ? (a.d[i] << sh) | (i ? a.d[i - 1] >> (32 - sh) : 0) : a.d[i];
    const size_t m = la - n;
    q.d.assign(m + 1, 0);
    for (size_t j = m + 1; j-- > 0;) {
        const uint64_t num = ((uint64_t)u[j + n] << 32) | u[j + n - 1];
        uint64_t qhat = num / v[n - 1], rhat = num % v[n - 1];
        while (qhat >> 32 || qhat * v[n - 2] > ((rhat << 32) | u[j + n - 2])) {
            --qhat;
            rhat += v[n - 1];
            if (rhat >> 32) break;
        }
        // u[j .. j+n] -= qhat * v
        int64_t borrow = 0;
        uint64_t carry = 0;
        for (size_t i = 0; i < n; ++i) {
            const uint64_t p = qhat * v[i] + carry;
            carry = p >> 32;
            const int64_t t = (int64_t)u[i + j] - borrow - (int64_t)(p & 0xFFFFFFFFull);
            u[i + j] = (uint32_t)t;
            borrow = t < 0 ? 1 : 0;
        }
        const int64_t t = (int64_t)u[j + n] - borrow - (int64_t)carry;
        u[j + n] = (uint32_t)t;
        if (t < 0) {   // qhat was one too large: add v back
            --qhat;
            uint64_t c = 0;
            for (size_t i = 0; i < n; ++i) {
                const uint64_t sum = (uint64_t)u[i + j] + v[i] + c;
                u[i + j] = (uint32_t)sum;
                c = sum >> 32;
            }
            u[j + n] += (uint32_t)c;
        }
        q.d[j] = (uint32_t)qhat;
    }
    q.trim();
    r.d.assign(n, 0);
    for (size_t i = 0; i < n; ++i)
        r.d[i] = sh ? (u[i] >> sh) | ((uint64_t)u[i + 1] << (32 - sh)) : u[i];
    r.trim();
}
static Big pow10_slow(unsigned e) {
    Big r(1);
    while (e >= 9) { r = mul_small(r, 1000000000u); e -= 9; }
    static const uint32_t p[9] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000};
    if (e) r = mul_small(r, p[e]);
    return r;
}
// the exponents the rows use (<= 2 DP + a few) come from a table built once
struct Pow10Table {
    enum { N = 96 };
    Big p[N];
    Pow10Table() { for (unsigned e = 0; e < N; ++e) p[e] = pow10_slow(e); }
};
static const Big &pow10(unsigned e) {
    static const Pow10Table t;
    static thread_local Big big;
    if (e < Pow10Table::N) return t.p[e];
    big = pow10_slow(e);
    return big;
}
// floor(sqrt(a)): Newton's iteration from above, x <- (x + a / x) / 2 until it stops decreasing
static Big isqrt(const Big &a) {
    if (a.zero()) return Big();
    // start from above: the top 62 bits of a give sqrt(a) to ~30 bits (hardware sqrt, rounded up with
    // a margin), so Newton's iteration needs 2-3 divisions instead of one per bit-doubling from 2^nb
    const size_t na = bits(a);
    size_t s = na > 62 ? na - 62 : 0;
    s += s & 1;                                // even shift
    uint64_t top = 0;
    for (size_t i = 0; i < 64 && s + i < na; ++i)
        if (bit(a, s + i)) top |= 1ull << i;
    uint64_t r0 = (uint64_t)std::sqrt((double)top) + 2;      // >= floor(sqrt(top)) + 1 despite rounding
    while (r0 * r0 <= top) ++r0;                              // r0 < 2^32 here, the product cannot wrap
    Big x(r0);                                 // x = r0 << (s / 2) >= sqrt((top + 1) << s) > sqrt(a)
    for (size_t i = 0; i < s / 2; ++i) shl1_or(x, false);
    for (;;) {
        Big qd, rd;
        divmod(a, x, qd, rd);
        Big y = add(x, qd);
        // y >>= 1
        uint32_t carry = 0;
        for (size_t i = y.d.size(); i-- > 0;) {
            const uint32_t nc = y.d[i] & 1u;
            y.d[i] = (y.d[i] >> 1) | (carry << 31);
            carry = nc;
        }
        y.trim();
        if (cmp(y, x) >= 0) return x;
        x = y;
    }
}
static std::string to_string(Big a) {
    if (a.zero()) return "0";
    std::string s;
    while (!a.zero()) {
        uint32_t rem = divmod_small(a, 1000000000u);
        for (int i = 0; i < 9; ++i) {
            s.push_back((char)('0' + rem % 10));
            rem /= 10;
            if (a.zero() && rem == 0) break;
        }
    }
    std::reverse(s.begin(), s.end());
    return s;
}

// ------------------------------------------------------------------ decimal = (-1)^neg * n / 10^e
struct Dec {
    Big n;
    unsigned e = 0;
    bool neg = false;
    Dec() {}
    Dec(uint64_t v) : n(v) {}
    Dec(uint64_t v, unsigned e_) : n(v), e(e_) { norm(); }
    void norm() {
        while (e > 0 && !n.zero()) {
            if (n.d[0] & 1u) break;            // odd: not a multiple of 10
            Big t = n;
            if (divmod_small(t, 10) != 0) break;
            n = t;
            --e;
        }
        if (n.zero()) { e = 0; neg = false; }
    }
};

static void align(const Dec &a, const Dec &b, Big &x, Big &y, unsigned &e) {
    e = std::max(a.e, b.e);
    x = a.e < e ? mul(a.n, pow10(e - a.e)) : a.n;
    y = b.e < e ? mul(b.n, pow10(e - b.e)) : b.n;
}
static Dec plus(const Dec &a, const Dec &b) {
    Big x, y;
    Dec r;
    align(a, b, x, y, r.e);
    if (a.neg == b.neg) { r.n = add(x, y); r.neg = a.neg; }
    else if (cmp(x, y) >= 0) { r.n = sub(x, y); r.neg = a.neg; }
    else { r.n = sub(y, x); r.neg = b.neg; }
    r.norm();
    return r;
}
static Dec minus(const Dec &a, Dec b) { b.neg = !b.neg; if (b.n.zero()) b.neg = false; return plus(a, b); }
static Dec times(const Dec &a, const Dec &b) {
    Dec r;
    r.n = mul(a.n, b.n);
    r.e = a.e + b.e;
    r.neg = a.neg != b.neg;
    r.norm();
    return r;
}
static int cmp(const Dec &a, const Dec &b) {
    if (a.neg != b.neg) return a.neg ? -1 : 1;
    Big x, y;
    unsigned e;
    align(a, b, x, y, e);
    int c = cmp(x, y);
    return a.neg ? -c : c;
}

enum { RM_UP = 0, RM_DOWN, RM_CEIL, RM_FLOOR, RM_HALF_UP, RM_HALF_DOWN, RM_HALF_EVEN };

// round(num/den) of magnitudes; neg is the sign of the quotient
static Big round_div(const Big &num, const Big &den, bool neg, int rm) {
    Big q, r;
    divmod(num, den, q, r);
    if (!r.zero()) {
        Big twice = add(r, r);
        int c = cmp(twice, den);
        bool inc;
        switch (rm) {
            case RM_UP: inc = true; break;
            case RM_DOWN: inc = false; break;
            case RM_CEIL: inc = !neg; break;
            case RM_FLOOR: inc = neg; break;
            case RM_HALF_UP: inc = c >= 0; break;
            case RM_HALF_DOWN: inc = c > 0; break;
            default: inc = c > 0 || (c == 0 && !q.zero() && (q.d[0] & 1u)); break;   // HALF_EVEN
        }
        if (inc) q = add(q, Big(1));
    }
    return q;
}

static const unsigned DP = 20;   // bignumber.js DECIMAL_PLACES

static Dec div(const Dec &a, const Dec &b, int rm) {
    Dec r;
    Big num = mul(a.n, pow10(b.e + DP));
    Big den = mul(b.n, pow10(a.e));
    r.neg = a.neg != b.neg;
    r.n = round_div(num, den, r.neg, rm);
    r.e = DP;
    r.norm();
    return r;
}
static Dec sqrt(const Dec &a, int rm) {   // a >= 0
    Dec r;
    // X = a * 10^(2 DP) = n * 10^(2 DP) / 10^e ; result = round(sqrt(X)) / 10^DP
    Big num = mul(a.n, pow10(2 * DP));
    Big den = pow10(a.e);
    Big fl_arg, rem;
    divmod(num, den, fl_arg, rem);
    Big fl = isqrt(fl_arg);
    bool exact = rem.zero() && cmp(mul(fl, fl), fl_arg) == 0;
    if (exact) r.n = fl;
    else if (rm == RM_UP || rm == RM_CEIL) r.n = add(fl, Big(1));
    else if (rm == RM_DOWN || rm == RM_FLOOR) r.n = fl;
    else {
        // nearest: floor(sqrt(4X)) is 2 fl or 2 fl + 1; a tie would need X = (fl + 1/2)^2 exactly,
        // which has an odd factor pattern no 2*DP-scaled decimal takes when it is not a square
        Big four, rem4;
        divmod(mul_small(num, 4), den, four, rem4);
        Big s2 = isqrt(four);
        Big q2, r2;
        divmod(add(s2, Big(1)), Big(2), q2, r2);
        r.n = q2;
    }
    r.e = DP;
    r.norm();
    return r;
}
static Dec round(const Dec &a, unsigned dp, int rm) {
    if (a.e <= dp) return a;
    Dec r;
    r.neg = a.neg;
    r.n = round_div(a.n, pow10(a.e - dp), a.neg, rm);
    r.e = dp;
    r.norm();
    return r;
}
static std::string text(const Dec &a) {
    std::string s = to_string(a.n);
    if (a.e) {
        if (s.size() <= a.e) s = std::string(a.e - s.size() + 1, '0') + s;
        s.insert(s.size() - a.e, ".");
    }
    return (a.neg ? "-" : "") + s;
}
static double to_number(const Dec &a) {
    // digits into a stack buffer, then glibc strtod (correctly rounded)
    char buf[Limbs::CAP * 10 + 32];
    char *end = buf + sizeof(buf) - 16, *p = end;
    Big t = a.n;
    if (t.zero()) *--p = '0';
    while (!t.zero()) {
        uint32_t rem = divmod_small(t, 1000000000u);
        for (int i = 0; i < 9; ++i) {
            *--p = (char)('0' + rem % 10);
            rem /= 10;
            if (t.zero() && rem == 0) break;
        }
    }
    if (a.neg) *--p = '-';
    snprintf(end, 16, "e-%u", a.e);
    return strtod(p, nullptr);
}
static bool parse(const char *s, Dec &out) {
    Dec r;
    bool neg = false, seen_dot = false, any = false;
    if (*s == '-') { neg = true; ++s; } else if (*s == '+') ++s;
    unsigned e = 0;
    Big n;
    for (; *s; ++s) {
        if (*s == '.') { if (seen_dot) return false; seen_dot = true; continue; }
        if (*s < '0' || *s > '9') break;
        n = add(mul_small(n, 10), Big((uint64_t)(*s - '0')));
        if (seen_dot) ++e;
        any = true;
    }
    if (!any) return false;
    long ex = 0;
    if (*s == 'e' || *s == 'E') { ex = strtol(s + 1, nullptr, 10); }
    else if (*s) return false;
    if (ex > 0) { n = mul(n, pow10((unsigned)ex)); }
    else e += (unsigned)(-ex);
    r.n = n; r.e = e; r.neg = neg;
    r.norm();
    out = r;
    return true;
}

static const Dec ETTA(1, 8);   // lib/stats.js:6

struct Thr { uint64_t n; unsigned e; uint64_t pn; unsigned pe; };
static const Thr FASTP[] = {   // lib/stats.js:56-112: strict '>' in this order
    {107016, 4, 1, 26}, {104862, 4, 1, 25}, {102663, 4, 1, 24}, {100416, 4, 1, 23},
    {981197, 5, 1, 22}, {95769, 4, 1, 21},  {933604, 5, 1, 20}, {908895, 5, 1, 19},
    {883511, 5, 1, 18}, {857394, 5, 1, 17}, {830479, 5, 1, 16}, {802686, 5, 1, 15},
    {773926, 5, 1, 14}, {74409, 4, 1, 13},  {713051, 5, 1, 12}, {68065, 4, 1, 11},
    {646695, 5, 1, 10}, {610941, 5, 1, 9},  {573073, 5, 1, 8},  {532672, 5, 1, 7},
    {489164, 5, 1, 6},  {441717, 5, 1, 5},  {389059, 5, 1, 4},  {329053, 5, 1, 3},
    {257583, 5, 1, 2},  {195996, 5, 5, 2},  {164485, 5, 1, 1},
};

static Dec z_score(int rm, uint64_t r1, uint64_t n1, uint64_t r2, uint64_t n2) {
    // lib/stats.js:21-42
    Dec p1 = plus(div(Dec(r1), Dec(n1), rm), ETTA);
    Dec p2 = plus(div(Dec(r2), Dec(n2), rm), ETTA);
    Dec p = div(plus(Dec(r1), Dec(r2)), plus(plus(Dec(n1), Dec(n2)), ETTA), rm);
    Dec q = minus(Dec(1), p);
    Dec inv = plus(div(Dec(1), plus(Dec(n1), ETTA), rm), div(Dec(1), plus(Dec(n2), ETTA), rm));
    Dec radicand = plus(times(times(p, q), inv), ETTA);
    if (radicand.neg) radicand = Dec(0);   // bignumber.js would give NaN; not reachable with r <= n
    Dec square = sqrt(radicand, rm);
    return div(minus(p1, p2), square, rm);
}
static Dec fastp(const Dec &z) {
    for (const Thr &t : FASTP)
        if (cmp(z, Dec(t.n, t.e)) > 0) return Dec(t.pn, t.pe);
    return Dec(1);
}

}  // namespace kjstats

using namespace kjstats;

bool kj_exact_zscore(int rm, uint64_t r1, uint64_t n1, uint64_t r2, uint64_t n2, double *z,
                     std::string *z_text) {
    if (n1 == 0 || n2 == 0) return false;
    Dec zz = z_score(rm, r1, n1, r2, n2);
    if (z) *z = to_number(zz);
    if (z_text) *z_text = text(zz);
    return true;
}

bool kj_exact_fastp_text(const char *z_text, double *p) {
    Dec z;
    if (!parse(z_text, z)) return false;
    if (p) *p = to_number(fastp(z));
    return true;
}

// lib/kmerFinderClient.js:41-92.  accepted = uScore > 0 && evalue(0.05) >= probability.
bool kj_exact_row(int rm, uint64_t uscore, uint64_t tscore, uint64_t uscore0, uint64_t tscore0,
                  uint64_t lengths, uint64_t ulength, uint64_t hits, uint64_t kmer_map_size,
                  uint64_t summary_templates, uint64_t summary_unique_lens, kj_row *out,
                  int *accepted) {
    *accepted = 0;
    if (!(uscore > 0)) return true;                       // :48  minScore = 0
    if (ulength == 0 || summary_unique_lens == 0 || lengths == 0) return false;
    Dec z = z_score(rm, uscore, ulength, hits, summary_unique_lens);            // :49
    Dec probability = times(fastp(z), Dec(summary_templates));                  // :50
    out->score = uscore;
    out->kmers_template = ulength;
    out->tscore = tscore;
    out->hits = hits;
    out->z = to_number(round(z, 2, rm));                                        // :79 z.round(2): global mode
    out->probability = to_number(probability);
    if (cmp(Dec(5, 2), probability) < 0) return true;                           // :51-52 evalue.cmp(p) >= 0
    Dec qden = plus(Dec(kmer_map_size), ETTA);
    Dec dden = plus(Dec(ulength), ETTA);
    Dec frac_q = div(times(Dec(200), Dec(uscore)), qden, rm);                   // :53-55
    Dec frac_d = div(times(Dec(100), Dec(uscore)), dden, rm);                   // :56-58
    Dec tot_q = div(times(Dec(200), Dec(uscore0)), qden, rm);                   // :59-64
    Dec tot_d = div(times(Dec(100), Dec(uscore0)), dden, rm);                   // :65-69
    Dec tot_cov = div(Dec(tscore0), Dec(lengths), rm);                          // :70-71
    Dec expected = div(times(Dec(hits), Dec(ulength)), Dec(summary_unique_lens), rm);   // :72-74
    Dec depth = div(Dec(tscore), Dec(lengths), rm);                             // :82
    out->expected = to_number(round(expected, 0, RM_HALF_EVEN));
    out->frac_q = to_number(round(frac_q, 2, RM_HALF_EVEN));
    out->frac_d = to_number(round(frac_d, 2, RM_HALF_EVEN));
    out->depth = to_number(round(depth, 2, RM_HALF_EVEN));
    out->total_frac_q = to_number(round(tot_q, 2, RM_HALF_EVEN));
    out->total_frac_d = to_number(round(tot_d, 2, RM_HALF_EVEN));
    out->total_temp_cover = to_number(round(tot_cov, 2, RM_HALF_EVEN));
    *accepted = 1;
    return true;
}

extern "C" int kj_stats_zscore(int rounding_mode, uint64_t r1, uint64_t n1, uint64_t r2, uint64_t n2,
                               double *z, char *z_text, uint64_t z_text_cap) {
    if (rounding_mode < 0 || rounding_mode > 6) return KJ_E_INVALID;
    std::string t;
    if (!kj_exact_zscore(rounding_mode, r1, n1, r2, n2, z, &t)) return KJ_E_INVALID;
    if (z_text && z_text_cap) {
        size_t n = std::min<size_t>(t.size(), (size_t)z_text_cap - 1);
        memcpy(z_text, t.data(), n);
        z_text[n] = 0;
    }
    return KJ_OK;
}

extern "C" int kj_stats_fastp_text(const char *z_text, double *p) {
    if (!z_text || !p) return KJ_E_INVALID;
    return kj_exact_fastp_text(z_text, p) ? KJ_OK : KJ_E_INVALID;
}

extern "C" int kj_stats_row(int rounding_mode, uint64_t uscore, uint64_t tscore, uint64_t uscore0,
                            uint64_t tscore0, uint64_t lengths, uint64_t ulength, uint64_t hits,
                            uint64_t kmer_map_size, uint64_t summary_templates,
                            uint64_t summary_unique_lens, kj_row *out, int *accepted) {
    if (!out || !accepted || rounding_mode < 0 || rounding_mode > 6) return KJ_E_INVALID;
    memset(out, 0, sizeof(*out));
    return kj_exact_row(rounding_mode, uscore, tscore, uscore0, tscore0, lengths, ulength, hits,
                        kmer_map_size, summary_templates, summary_unique_lens, out, accepted)
               ? KJ_OK
               : KJ_E_INVALID;
}
