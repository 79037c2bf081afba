// kj_device.cuh -- device data structures of the count path: the open-addressing k-mer table,
// the byte-string side table for irregular k-mers, overflow lists and the device counters.
#pragma once
#include <stdint.h>
#include "kj_bits.cuh"

#define KJ_EMPTY 0xFFFFFFFFFFFFFFFFull
#define KJ_MAX_PROBE 1024
#define KJ_MAX_PROBE_IRR 128      // side table: every probe compares 32-byte keys

// error bits raised by kernels (KjCounters::error_flags)
#define KJ_DEV_E_LINE_TOO_LONG 1u      // position does not fit KJ_POS_BITS
#define KJ_DEV_E_LINE_EXCEEDS_HALO 2u  // line-oriented kernel: line end not inside the readable bytes
#define KJ_DEV_E_READS_OVERFLOW 4u     // read index does not fit 36 bits
#define KJ_DEV_E_XCHG_OVERFLOW 8u      // fixed-capacity exchange: a segment held more records than its capacity
#define KJ_DEV_E_XCHG_INCOMPLETE 16u   // fixed-capacity exchange: a sender's count had emissions or entries left over

// regular table: ACGT-only, full-length k-mers as 2k-bit integers (first base most significant).
// SoA: three arrays of cap entries (8+8+8 bytes per slot; ords may be null with KJ_F_NO_ORDER).
struct KjTable {
    uint64_t *keys;
    uint64_t *counts;
    uint64_t *ords;
    uint64_t mask;   // cap - 1, cap is a power of two
};

// irregular table: exact byte-string keys (<= 32 bytes): k-mers containing a non-ACGT byte and
// the short windows of step > 1 (lib/kmers.js:89-99).  state: 0 empty, 1 busy, 2+len ready.
struct KjIrrTable {
    uint32_t *state;
    uint8_t *keys;      // cap * 32
    uint64_t *counts;
    uint64_t *ords;
    uint64_t mask;
};

struct KjOverflow {
    uint64_t *rec;      // regular: {key, ord} pairs; add is always 1
    uint64_t cap;       // entries
    uint64_t *irr_rec;  // irregular: {buffer offset, (len<<1)|strand, ord}
    uint64_t irr_cap;
};

struct KjCounters {
    unsigned long long n_unique;       // host view: sum of n_unique_part, filled in by pull_counters
    unsigned long long n_irr_unique;
    unsigned long long n_overflow;
    unsigned long long n_irr_overflow;
    unsigned long long n_cand;         // filter path: entry slots handed out in the current launch
    unsigned long long n_occ;          // emitted occurrences (regular + irregular)
    unsigned long long n_bases;        // sum of processed sequence-line lengths
    unsigned long long special_count;  // the one key equal to KJ_EMPTY (k = 32, all 'G')
    unsigned long long special_ord;
    unsigned long long carry_lines[2]; // stream state, double buffered by launch parity
    unsigned long long carry_last[2];  // virtual offset + 1 of the last '\n' so far (0 = none)
    unsigned int error_flags;
    unsigned int ticket;               // dynamic tile counter
    unsigned int pad_;
    // new keys are counted in 64 places (same-address atomics serialise in L2; a launch can add 10^5..10^9 keys)
    unsigned long long n_unique_part[64];
    unsigned long long n_compact;      // compaction cursors
    unsigned long long n_irr_compact;
    // totals of the whole job, summed from the segment headers of a fixed-capacity exchange (kj_counts_merge_segments)
    unsigned long long x_lines, x_bases, x_occ, x_bytes;
};

__device__ __forceinline__ uint64_t kj_ld_volatile(const uint64_t *p) {
    return *reinterpret_cast<const volatile uint64_t *>(p);
}
__device__ __forceinline__ uint32_t kj_ld_volatile(const uint32_t *p) {
    return *reinterpret_cast<const volatile uint32_t *>(p);
}

// count[key] += add ; ord[key] = min(ord[key], ord).  false = probe limit hit (caller spills).
// READ_ORD: look at the stored ordinal before the atomic min.  The stream is read in order, so after a key's first occurrences
// its stored ordinal is already the smaller one and the plain read filters most of the 64-bit atomics out: that pays where the
// atomics are the bound (dense emission, BASELINE config 5), and costs a round trip where they are not (filter path).
template <bool READ_ORD = true>
__device__ __forceinline__ bool kj_insert(const KjTable &t, KjCounters *ctr, uint64_t key,
                                          uint64_t ord, uint64_t add) {
    if (key == KJ_EMPTY) {
        atomicAdd(&ctr->special_count, (unsigned long long)add);
        atomicMin(&ctr->special_ord, (unsigned long long)ord);
        return true;
    }
    uint64_t slot = kj_mix64(key) & t.mask;
    for (int probe = 0; probe < KJ_MAX_PROBE; ++probe) {
        uint64_t cur = kj_ld_volatile(&t.keys[slot]);
        // the slot's ordinal is requested with its key, not after the comparison: one round trip instead of two.  It may be
        // stale by the time it is looked at, but ordinals only go down, so "stale <= ord" still proves "current <= ord".
        const uint64_t seen = (READ_ORD && t.ords) ? kj_ld_volatile(&t.ords[slot]) : ~0ull;
        if (cur == KJ_EMPTY) {
            cur = atomicCAS((unsigned long long *)&t.keys[slot], (unsigned long long)KJ_EMPTY,
                            (unsigned long long)key);
            if (cur == KJ_EMPTY) {
                atomicAdd(&ctr->n_unique_part[slot & 63], 1ull);
                cur = key;
            }
        }
        if (cur == key) {
            // results unused: both become fire-and-forget reductions (RED), nothing waits on them
            atomicAdd((unsigned long long *)&t.counts[slot], (unsigned long long)add);
            // the stream is read in order, so after a key's first occurrences its stored ordinal is already the
            // smaller one: a plain read filters most of the 64-bit atomics out
            if (t.ords && seen > ord)
                atomicMin((unsigned long long *)&t.ords[slot], (unsigned long long)ord);
            return true;
        }
        slot = (slot + 1) & t.mask;
    }
    return false;
}

__device__ __forceinline__ uint64_t kj_hash_bytes(const uint8_t *k, uint32_t len) {
    uint64_t h = 0xCBF29CE484222325ull ^ len;
    for (uint32_t i = 0; i < len; ++i) { h ^= k[i]; h *= 0x100000001B3ull; }
    return kj_mix64(h);
}

// key32: 32 bytes, zero padded beyond len.  Slot protocol: CAS state 0->1 claims, the claimer
// writes the key and publishes state = 2+len; readers of a busy slot wait for the publication.
static __device__ __noinline__ bool kj_insert_irr(const KjIrrTable &t, KjCounters *ctr,
                                           const uint8_t *key32, uint32_t len, uint64_t ord,
                                           uint64_t add) {
    uint64_t slot = kj_hash_bytes(key32, len) & t.mask;
    const uint64_t *kw = reinterpret_cast<const uint64_t *>(key32);
    for (int probe = 0; probe < KJ_MAX_PROBE_IRR; ++probe) {
        uint32_t st = kj_ld_volatile(&t.state[slot]);
        if (st == 0) {
            st = atomicCAS(&t.state[slot], 0u, 1u);
            if (st == 0) {
                uint64_t *dst = reinterpret_cast<uint64_t *>(t.keys + slot * 32);
                dst[0] = kw[0]; dst[1] = kw[1]; dst[2] = kw[2]; dst[3] = kw[3];
                __threadfence();
                atomicExch(&t.state[slot], 2u + len);
                atomicAdd(&ctr->n_irr_unique, 1ull);
                atomicAdd((unsigned long long *)&t.counts[slot], (unsigned long long)add);
                atomicMin((unsigned long long *)&t.ords[slot], (unsigned long long)ord);
                return true;
            }
        }
        while (st == 1) {          // another thread is writing this slot's key
            __nanosleep(64);
            st = kj_ld_volatile(&t.state[slot]);
        }
        if (st == 2u + len) {
            __threadfence();
            const uint64_t *src = reinterpret_cast<const uint64_t *>(t.keys + slot * 32);
            if (kj_ld_volatile(&src[0]) == kw[0] && kj_ld_volatile(&src[1]) == kw[1] &&
                kj_ld_volatile(&src[2]) == kw[2] && kj_ld_volatile(&src[3]) == kw[3]) {
                atomicAdd((unsigned long long *)&t.counts[slot], (unsigned long long)add);
                atomicMin((unsigned long long *)&t.ords[slot], (unsigned long long)ord);
                return true;
            }
        }
        slot = (slot + 1) & t.mask;
    }
    return false;
}

// owner of a regular key for the multi-GPU exchange; salted so it is independent of the slot hash
__host__ __device__ __forceinline__ uint32_t kj_owner_key(uint64_t key, uint32_t n_parts) {
    return (uint32_t)((kj_mix64(key ^ 0x9E3779B97F4A7C15ull) >> 32) % n_parts);
}
// owner of a byte-string (side-table) k-mer: hash of the zero-padded 32 key bytes and the length
__host__ __device__ __forceinline__ uint32_t kj_owner_bytes(const uint8_t *key32, uint32_t len, uint32_t n_parts) {
    uint64_t h = 0x243F6A8885A308D3ull ^ (uint64_t)len;
    for (int i = 0; i < 4; ++i) {
        uint64_t w = 0;
        for (int j = 0; j < 8; ++j) w |= (uint64_t)key32[8 * i + j] << (8 * j);
        h = kj_mix64(h ^ w);
    }
    return (uint32_t)((h >> 32) % n_parts);
}
