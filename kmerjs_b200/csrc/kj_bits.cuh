// kj_bits.cuh -- SIMD-in-register primitives shared by the extraction kernels.
// All functions are pure integer code, usable on host and device (the host build is what
// tests/test_bits_host.py exercises exhaustively before any GPU time is spent).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define KJ_HD __host__ __device__ __forceinline__
#else
#define KJ_HD static inline
#endif

// 2-bit code of a base byte: (b >> 1) & 3  ->  A=0 C=1 T=2 G=3 (upper and lower case alike).
// It is a function of the byte, so "codes equal" is a necessary condition for "bytes equal":
// matching in code space is an exact superset filter, verified on the bytes afterwards.
KJ_HD uint32_t kj_code(uint32_t b) { return (b >> 1) & 3u; }

// complement in code space: A<->T, C<->G  ==  code ^ 2
KJ_HD bool kj_is_acgt(uint32_t b) { return b == 'A' || b == 'C' || b == 'G' || b == 'T'; }

// reference complement (lib/kmers.js:12-17,31-38): only upper-case ACGT are mapped
KJ_HD uint8_t kj_comp_byte(uint8_t b) {
    switch (b) {
        case 'A': return 'T';
        case 'T': return 'A';
        case 'G': return 'C';
        case 'C': return 'G';
        default: return b;
    }
}

// 4 bytes (little-endian word) -> 8 bits: code of byte i at bits 2i..2i+1.
// t keeps bits 1..2 of every byte; one multiply gathers the four 2-bit fields into the top byte
// (the partial products never overlap, so there are no carries).
KJ_HD uint32_t kj_pack4(uint32_t w) {
    return ((w & 0x06060606u) * 0x00820820u) >> 24;
}

// 16 bytes -> 32 bits: code of byte p at bits 2p..2p+1
KJ_HD uint32_t kj_pack16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
#if defined(__CUDA_ARCH__)
    // the multiply leaves the four codes of a word in its top byte: three byte permutes collect the top
    // bytes of the four products (11 instructions instead of 17 with shifts and ORs)
    const uint32_t p0 = (w0 & 0x06060606u) * 0x00820820u, p1 = (w1 & 0x06060606u) * 0x00820820u;
    const uint32_t p2 = (w2 & 0x06060606u) * 0x00820820u, p3 = (w3 & 0x06060606u) * 0x00820820u;
    return __byte_perm(__byte_perm(p0, p1, 0x0073u), __byte_perm(p2, p3, 0x0073u), 0x5410u);
#else
    return kj_pack4(w0) | (kj_pack4(w1) << 8) | (kj_pack4(w2) << 16) | (kj_pack4(w3) << 24);
#endif
}

// 4 bytes -> 4 bits: bit i set iff byte i == '\n'
KJ_HD uint32_t kj_nl4(uint32_t w) {
    uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;   // msb set iff low 7 bits differ
    uint32_t z = ~(t | w) & 0x80808080u;                            // msb set iff byte == 0x0A
    return (z * 0x00204081u) >> 28;                                 // gather the 4 msbs
}

// 16 bytes -> 16 bits
// msb of byte i set iff byte i == '\n'
KJ_HD uint32_t kj_nl_flags4(uint32_t w) {
    uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}

KJ_HD uint32_t kj_nl16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
#if defined(__CUDA_ARCH__)
    // the 4-way byte dot product gathers four flag bytes (0x80 or 0) in one instruction: weights 1,2,4,8
    // for one word, 16..128 for the next, so two words leave 8 mask bits times 128 in the accumulator
    uint32_t lo = __dp4a(kj_nl_flags4(w0), 0x08040201u, 0u);
    lo = __dp4a(kj_nl_flags4(w1), 0x80402010u, lo);
    uint32_t hi = __dp4a(kj_nl_flags4(w2), 0x08040201u, 0u);
    hi = __dp4a(kj_nl_flags4(w3), 0x80402010u, hi);
    return (lo >> 7) | ((hi << 1) & 0xFF00u);
#else
    return kj_nl4(w0) | (kj_nl4(w1) << 4) | (kj_nl4(w2) << 8) | (kj_nl4(w3) << 12);
#endif
}

// 4 bytes -> nonzero iff some byte is not one of 'A','C','G','T' (upper case).  The 2-bit code of
// every byte selects the letter that has this code from the table word "ACTG" (one PRMT); a byte
// is a base iff it equals that letter.
KJ_HD uint32_t kj_not_acgt4(uint32_t w) {
    const uint32_t c = (w >> 1) & 0x03030303u;            // code of byte i in bits 8i..8i+1
    const uint32_t t = c | (c >> 4);                      // byte 0: c0 | c1 << 4, byte 2: c2 | c3 << 4
#if defined(__CUDA_ARCH__)
    const uint32_t sel = __byte_perm(t, 0u, 0x4420u);     // selector nibbles c0, c1, c2, c3
    const uint32_t expect = __byte_perm(0x47544341u, 0u, sel);   // table bytes: 'A','C','T','G' for codes 0..3
#else
    const uint32_t sel = (t & 0xFFu) | ((t >> 8) & 0xFF00u);
    const uint32_t tab = 0x47544341u;
    uint32_t expect = 0;
    for (int i = 0; i < 4; ++i) expect |= ((tab >> (8u * ((sel >> (4 * i)) & 3u))) & 0xFFu) << (8 * i);
#endif
    return w ^ expect;
}

// 4 bytes -> 4 bits: bit i set iff byte i != 0
KJ_HD uint32_t kj_nz4(uint32_t x) {
    const uint32_t t = (x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    const uint32_t z = (t | x) & 0x80808080u;
    return (z * 0x00204081u) >> 28;
}
// 16 bytes -> 16 bits: bit p set iff byte p is not one of 'A','C','G','T'
KJ_HD uint32_t kj_bad16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return kj_nz4(kj_not_acgt4(w0)) | (kj_nz4(kj_not_acgt4(w1)) << 4) | (kj_nz4(kj_not_acgt4(w2)) << 8) |
           (kj_nz4(kj_not_acgt4(w3)) << 12);
}

// 4 bytes -> msb of byte i set iff byte i == '\n'
KJ_HD uint32_t kj_nl_msb4(uint32_t w) {
    const uint32_t t = ((w ^ 0x0A0A0A0Au) & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | w) & 0x80808080u;
}

// reverse the order of the 32 two-bit fields of a 64-bit word
KJ_HD uint64_t kj_pairrev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    const uint32_t lo = __brev((uint32_t)(x >> 32)), hi = __brev((uint32_t)x);
#else
    uint32_t a = (uint32_t)(x >> 32), b = (uint32_t)x, lo = 0, hi = 0;
    for (int i = 0; i < 32; ++i) { if (a >> i & 1) lo |= 1u << (31 - i); if (b >> i & 1) hi |= 1u << (31 - i); }
#endif
    const uint64_t r = ((uint64_t)hi << 32) | lo;          // all 64 bits reversed: fields reversed, bits inside swapped
    return ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
}

// low 32 bits of (hi:lo) >> s, 0 <= s < 32
KJ_HD uint32_t kj_funnel_r(uint32_t lo, uint32_t hi, uint32_t s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, s);
#else
    return s ? (lo >> s) | (hi << (32 - s)) : lo;
#endif
}

// 16 code lanes starting `d` lanes (0..31) after the first lane of c0; c0,c1,c2 consecutive words
KJ_HD uint32_t kj_lanes(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t d) {
    uint32_t r = (d & 15u) * 2u;
    return (d & 16u) ? kj_funnel_r(c1, c2, r) : kj_funnel_r(c0, c1, r);
}

// lanes (bit 2p) whose 2-bit field in acc is zero
KJ_HD uint32_t kj_zero_lanes(uint32_t acc) {
    return ~(acc | (acc >> 1)) & 0x55555555u;
}

// 64-bit finalizer (splitmix64): slot hash and owner hash of 2-bit keys
KJ_HD uint64_t kj_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// ordinal = read index (36 bits) | strand (1 bit) | position (27 bits); see DESIGN.md "first-seen order"
#define KJ_POS_BITS 27
#define KJ_POS_MAX ((1ull << KJ_POS_BITS) - 1)
KJ_HD uint64_t kj_ordinal(uint64_t read_idx, uint32_t strand, uint64_t pos) {
    return (read_idx << (KJ_POS_BITS + 1)) | ((uint64_t)strand << KJ_POS_BITS) | pos;
}
