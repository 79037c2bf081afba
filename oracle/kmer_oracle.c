/*
 * CPU oracle for the kmerjs extraction+count path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the reference algorithm (josl/kmerjs), written from its behaviour:
 * byte-string keys in an insertion-ordered hash map, no 2-bit tricks, no SIMD.  It mirrors
 * the reference's algorithm, it is not an optimised CPU k-mer counter.  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may load the library built from
 * this file; nothing under kmerjs_b200/ does.
 *
 *   ko_complement      lib/kmers.js:12-17,31-38   (A<->T, G<->C upper-case only, reversed)
 *   kmers_in_line      lib/kmers.js:88-100        (substring clip => short keys when step>1)
 *   ko_count (lines)   lib/kmers.js:114-136       (split on '\n', non-empty tail flushed)
 *   ko_count (FSM)     lib/kmers.js:143-171       (i==1 && length>1 ; i==3 -> 0)
 *   ko_size            lib/kmers.js:172-178       (kmerMapSize)
 *
 * Parity pinning: tests/test_oracle_golden.py checks this library against the reference's
 * known answers and against oracle/kmer_oracle.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint64_t off;    /* key bytes offset in arena */
    uint32_t len;
    uint64_t count;
} ko_entry;

typedef struct {
    ko_entry *ent;   /* insertion order */
    uint64_t n_ent, cap_ent;
    uint8_t *arena;
    uint64_t n_arena, cap_arena;
    uint64_t *slots; /* index+1 into ent, 0 = empty */
    uint64_t n_slots; /* power of two */
    uint64_t lines;
} ko_map;

static uint64_t ko_hash(const uint8_t *p, uint32_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (uint32_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
    h ^= h >> 29; h *= 0xbf58476d1ce4e5b9ULL; h ^= h >> 32;
    return h;
}

static void ko_rehash(ko_map *m) {
    uint64_t ns = m->n_slots * 2;
    uint64_t *s = (uint64_t *)calloc(ns, sizeof(uint64_t));
    for (uint64_t i = 0; i < m->n_ent; i++) {
        uint64_t h = ko_hash(m->arena + m->ent[i].off, m->ent[i].len) & (ns - 1);
        while (s[h]) h = (h + 1) & (ns - 1);
        s[h] = i + 1;
    }
    free(m->slots);
    m->slots = s;
    m->n_slots = ns;
}

static void ko_bump(ko_map *m, const uint8_t *key, uint32_t len) {
    uint64_t h = ko_hash(key, len) & (m->n_slots - 1);
    while (m->slots[h]) {
        ko_entry *e = &m->ent[m->slots[h] - 1];
        if (e->len == len && memcmp(m->arena + e->off, key, len) == 0) { e->count++; return; }
        h = (h + 1) & (m->n_slots - 1);
    }
    if (m->n_ent == m->cap_ent) {
        m->cap_ent *= 2;
        m->ent = (ko_entry *)realloc(m->ent, m->cap_ent * sizeof(ko_entry));
    }
    if (m->n_arena + len > m->cap_arena) {
        while (m->n_arena + len > m->cap_arena) m->cap_arena *= 2;
        m->arena = (uint8_t *)realloc(m->arena, m->cap_arena);
    }
    memcpy(m->arena + m->n_arena, key, len);
    m->ent[m->n_ent].off = m->n_arena;
    m->ent[m->n_ent].len = len;
    m->ent[m->n_ent].count = 1;
    m->n_arena += len;
    m->slots[h] = ++m->n_ent;
    if (m->n_ent * 2 > m->n_slots) ko_rehash(m);
}

void ko_complement(const uint8_t *in, uint64_t n, uint8_t *out) {
    for (uint64_t i = 0; i < n; i++) {
        uint8_t c = in[n - 1 - i];
        switch (c) {
            case 'A': c = 'T'; break;
            case 'T': c = 'A'; break;
            case 'G': c = 'C'; break;
            case 'C': c = 'G'; break;
            default: break;
        }
        out[i] = c;
    }
}

static void kmers_in_line(ko_map *m, const uint8_t *line, int64_t L, int k, int step,
                          const uint8_t *prefix, int plen) {
    int64_t ini = 0;
    int64_t stop = L - k;
    for (int64_t index = 0; index <= stop; index++) {
        int64_t a = ini < L ? ini : L;
        int64_t b = ini + k < L ? ini + k : L;
        int64_t len = b - a;
        if (len >= plen && memcmp(line + a, prefix, (size_t)plen) == 0)
            ko_bump(m, line + a, (uint32_t)len);
        ini += step;
    }
}

ko_map *ko_count(const uint8_t *data, uint64_t n, const uint8_t *prefix, int plen, int k,
                 int step) {
    ko_map *m = (ko_map *)calloc(1, sizeof(ko_map));
    m->cap_ent = 1024;
    m->ent = (ko_entry *)malloc(m->cap_ent * sizeof(ko_entry));
    m->cap_arena = 1 << 16;
    m->arena = (uint8_t *)malloc(m->cap_arena);
    m->n_slots = 4096;
    m->slots = (uint64_t *)calloc(m->n_slots, sizeof(uint64_t));
    uint8_t *rc = NULL;
    uint64_t rc_cap = 0;
    int i = 0;
    uint64_t pos = 0;
    while (pos < n) {
        const uint8_t *nl = (const uint8_t *)memchr(data + pos, '\n', n - pos);
        uint64_t end = nl ? (uint64_t)(nl - data) : n;  /* unterminated tail is non-empty here */
        uint64_t L = end - pos;
        if (i == 1 && L > 1) {
            if (L > rc_cap) { rc_cap = L * 2; rc = (uint8_t *)realloc(rc, rc_cap); }
            kmers_in_line(m, data + pos, (int64_t)L, k, step, prefix, plen);
            ko_complement(data + pos, L, rc);
            kmers_in_line(m, rc, (int64_t)L, k, step, prefix, plen);
        } else if (i == 3) {
            i = -1;
        }
        i++;
        m->lines++;
        pos = end + 1;
    }
    free(rc);
    return m;
}

uint64_t ko_size(const ko_map *m) { return m->n_ent; }
uint64_t ko_lines(const ko_map *m) { return m->lines; }
uint64_t ko_key_bytes(const ko_map *m) { return m->n_arena; }

/* keys: concatenated key bytes (ko_key_bytes), key_len[i], counts[i]; insertion order */
void ko_export(const ko_map *m, uint8_t *keys, uint32_t *key_len, uint64_t *counts) {
    memcpy(keys, m->arena, m->n_arena);
    for (uint64_t i = 0; i < m->n_ent; i++) {
        key_len[i] = m->ent[i].len;
        counts[i] = m->ent[i].count;
    }
}

void ko_free(ko_map *m) {
    if (!m) return;
    free(m->ent); free(m->arena); free(m->slots); free(m);
}

/* ------------------------------------------------------------------------------------------------
 * Winner-takes-all template scoring on integer ids -- the integer part of the scoring path, so that
 * the parity tests can follow databases of 10^4 templates / 10^6..10^8 (k-mer, template) pairs in
 * seconds (the Python restatement in kmer_oracle.py is the readable one and is what this is checked
 * against in tests/test_oracle_golden.py).  It follows the reference literally, full recount per
 * round included:
 *
 *   first match        lib/kmerFinderServer.js:171-226  (query k-mers in Map order, lists in DB order)
 *   getMatches         lib/kmerFinderClient.js:232-271  (recount every template's `kmers` Set against
 *                                                        the shrinking query Map; dead templates are
 *                                                        deleted from firstMatches)
 *   sort + winner      lib/kmerFinderClient.js:100-109,179-186  (stable sort: first maximum in
 *                                                        firstMatches iteration order)
 *   removeWinnerKmers  lib/kmerFinderClient.js:220-230
 *   loop               lib/kmerFinderClient.js:273-289  (maxHits; nHits === 0 throws)
 *
 * The evalue gate (matchSummary, lib/kmerFinderClient.js:41-92) is exact decimal arithmetic and stays
 * in Python: `gate(round, template, u, tau, hits, u0, t0)` returns non-zero to accept the winner.
 *
 * Query entry q (0..Q-1, Map order) has count qcount[q] and the template list qt[qoff[q]..qoff[q+1])
 * in DB list order (duplicates inside one list count once per the `kmers` Set but every list entry
 * adds to uScore/tScore in first match, exactly as the reference does).
 * Returns the number of accepted rounds; *status: 0 loop ended normally, 1 'nHits === 0' thrown,
 * 2 'kmerResults.length === 0' thrown.  out[4*i..] = {template, u, tau, hits} of accepted round i;
 * order_out (T entries, may be NULL) receives the templates in first-encounter order, *n_order their
 * number; u0/t0 (T entries each, may be NULL) the first-match scores; *hits0 the first-match hits.
 */
typedef int (*ko_gate_fn)(uint32_t round, uint32_t tmpl, uint64_t u, uint64_t tau, uint64_t hits,
                          uint64_t u0, uint64_t t0);

uint32_t ko_wta(uint64_t Q, const uint64_t *qcount, const uint64_t *qoff, const uint32_t *qt, uint32_t T,
                uint32_t max_hits, ko_gate_fn gate, uint64_t *out, int *status, uint32_t *order_out,
                uint32_t *n_order, uint64_t *u0_out, uint64_t *t0_out, uint64_t *hits0) {
    uint64_t *u0 = (uint64_t *)calloc(T ? T : 1, 8), *t0 = (uint64_t *)calloc(T ? T : 1, 8);
    uint64_t *koff = (uint64_t *)calloc((uint64_t)T + 2, 8);
    uint32_t *order = (uint32_t *)malloc((T ? T : 1) * 4);
    uint8_t *seen = (uint8_t *)calloc(T ? T : 1, 1);
    uint8_t *alive_q = (uint8_t *)malloc(Q ? Q : 1);
    uint32_t n_ord = 0;
    uint64_t hits = 0;
    memset(alive_q, 1, Q ? Q : 1);
    /* first match: scores, first-encounter order, and the size of every template's `kmers` Set */
    uint32_t *stamp = (uint32_t *)malloc((T ? T : 1) * 4);
    memset(stamp, 0xFF, (T ? T : 1) * 4);
    for (uint64_t q = 0; q < Q; q++) {
        hits += qoff[q + 1] - qoff[q];
        for (uint64_t j = qoff[q]; j < qoff[q + 1]; j++) {
            uint32_t t = qt[j];
            if (!seen[t]) { seen[t] = 1; order[n_ord++] = t; }
            u0[t] += 1;
            t0[t] += qcount[q];
            if (stamp[t] != (uint32_t)q) { stamp[t] = (uint32_t)q; koff[t + 2]++; }   /* Set: once per k-mer */
        }
    }
    for (uint32_t t = 0; t < T; t++) koff[t + 2] += koff[t + 1];
    uint32_t *kq = (uint32_t *)malloc((koff[T + 1] ? koff[T + 1] : 1) * 4);   /* kmers Sets, insertion order */
    memset(stamp, 0xFF, (T ? T : 1) * 4);
    for (uint64_t q = 0; q < Q; q++)
        for (uint64_t j = qoff[q]; j < qoff[q + 1]; j++) {
            uint32_t t = qt[j];
            if (stamp[t] != (uint32_t)q) { stamp[t] = (uint32_t)q; kq[koff[t + 1]++] = (uint32_t)q; }
        }
    /* now koff[t]..koff[t+1] is the Set of template t */
    if (order_out) memcpy(order_out, order, n_ord * 4);
    if (n_order) *n_order = n_ord;
    if (u0_out) memcpy(u0_out, u0, (uint64_t)T * 8);
    if (t0_out) memcpy(t0_out, t0, (uint64_t)T * 8);
    if (hits0) *hits0 = hits;
    uint32_t hit_counter = 0;
    *status = 0;
    uint8_t *in_first = (uint8_t *)malloc(T ? T : 1);   /* still a key of firstMatches */
    memset(in_first, 1, T ? T : 1);
    uint64_t *u = (uint64_t *)malloc((T ? T : 1) * 8), *tau = (uint64_t *)malloc((T ? T : 1) * 8);
    /* `firstMatches = templates` in the first round (lib/kmerFinderClient.js:182-184): the totals of
     * matchSummary are the first round's recount */
    uint64_t *u1 = (uint64_t *)calloc(T ? T : 1, 8), *t1 = (uint64_t *)calloc(T ? T : 1, 8);
    int not_found = 1;
    if (hits == 0) { *status = 1; not_found = 0; }      /* 'No hits were found!' (first match) */
    while (not_found && hit_counter < max_hits) {
        /* getMatches: every live template recounted against the query Map */
        uint64_t n_hits = 0;
        uint32_t best = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < n_ord; i++) {
            uint32_t t = order[i];
            if (!in_first[t]) continue;
            uint64_t uu = 0, tt = 0;
            for (uint64_t j = koff[t]; j < koff[t + 1]; j++)
                if (alive_q[kq[j]]) { uu += 1; tt += qcount[kq[j]]; }
            u[t] = uu; tau[t] = tt;
            if (uu) {
                n_hits += uu;
                if (best == 0xFFFFFFFFu || uu > u[best]) best = t;   /* stable sort: first maximum */
            } else {
                in_first[t] = 0;                                     /* delete firstMatches[name] */
            }
        }
        if (n_hits == 0) { *status = 1; break; }
        if (hit_counter == 0) { memcpy(u1, u, (uint64_t)T * 8); memcpy(t1, tau, (uint64_t)T * 8); }
        if (gate(hit_counter, best, u[best], tau[best], n_hits, u1[best], t1[best])) {
            out[4 * hit_counter + 0] = best;
            out[4 * hit_counter + 1] = u[best];
            out[4 * hit_counter + 2] = tau[best];
            out[4 * hit_counter + 3] = n_hits;
            hit_counter++;
            for (uint64_t j = koff[best]; j < koff[best + 1]; j++) alive_q[kq[j]] = 0;
        } else {
            not_found = 0;
        }
    }
    if (hit_counter == 0 && *status == 0) *status = 2;
    free(u0); free(t0); free(koff); free(order); free(seen); free(alive_q); free(stamp); free(kq);
    free(in_first); free(u); free(tau); free(u1); free(t1);
    return hit_counter;
}
