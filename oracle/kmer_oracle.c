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
