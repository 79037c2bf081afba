// kj_scan.cuh -- shared pieces of the extraction kernels (K1+K2+K3 of SURVEY.md appendix B) for sm_100a, and the two
// kernels that are not the hot one:
//
//   kj_scan_dense_kernel   empty prefix, step 1, k >= 2 (BASELINE config 5): every window is an emission
//   kj_scan_lines_kernel   everything else (step > 1, k = 1 with the empty prefix) and the independent second
//                          implementation the tests compare with (KJ_F_FORCE_GENERIC)
//
// Both read every FASTQ byte once, in tiles handed out by an atomic ticket: per 16-byte chunk a word of 2-bit base codes
// and a newline mask; a decoupled look-back over the tiles gives every tile the number of '\n' before it -- the
// reference's record FSM is "line index mod 4" (lib/kmers.js:151-163), which is global state.  The filter path
// (step 1, 1 <= |prefix| <= k) lives in kj_scan_warp.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "kj_device.cuh"

#define KJ_TILE_BYTES 28672                              // 56 rows of 512 bytes: 8 chunks for each of 224 stream threads,
#define KJ_TILE_CHUNKS (KJ_TILE_BYTES / 16)              // 7 for each of the 256 threads of the line kernel
#define KJ_ROWS (KJ_TILE_CHUNKS / 32)                    // a row = 32 chunks = one warp-wide 16-byte load
#define KJ_THREADS 256                                   // line kernel
#define KJ_QCAP 1536                                     // line queue entries (line kernel: <= 1024 per round)
#define KJ_MAX_MP 8                                      // filter symbols used in code space
#define KJ_NO_TILE 0xFFFFFFFFu
#define KJ_ST_AGG 1ull
#define KJ_ST_INC 2ull
#define KJ_ST_MASK 0x3FFFFFFFFFFFFFFFull

// where complement(prefix) sits inside the 48 code lanes a chunk looks at (c0,c1,c2)
#define KJ_RC_LOW 0      // every filter symbol within 16 lanes of the window start: (c0,c1)
#define KJ_RC_HIGH 1     // every one at 16 or more: (c1,c2)
#define KJ_RC_MIXED 2

struct KjScanArgs {
    const uint8_t *buf;   // 16-byte aligned
    uint64_t n;           // readable bytes
    uint64_t own_n;       // window starts / newlines in [0, own_n) belong to this launch
    uint64_t voff;        // virtual stream offset of buf[0] (base_col + bytes consumed so far)
    uint32_t n_tiles;
    uint32_t parity;      // which carry slot to read (the other is written)
    uint32_t final_;      // buf + n is end of stream
    uint32_t k, step, m;  // k-mer length, step, prefix length
    uint32_t order;       // track first-seen ordinals
    uint32_t n_strands;   // 2, or 1 with KJ_F_FORWARD_ONLY
    uint32_t line_gate;   // 1: lines of length <= 1 are not processed (lib/kmers.js:151)
    uint32_t count_bases; // KJ_F_COUNT_BASES
    uint32_t mp;          // number of filter symbols (<= min(m, KJ_MAX_MP))
    uint32_t rc_shift;    // k - m: lane offset of complement(prefix) inside a window
    uint32_t pat_f[KJ_MAX_MP];   // code of prefix[i] replicated to all 16 lanes
    uint32_t pat_r[KJ_MAX_MP];   // code of complement(prefix)[i], replicated
    uint8_t prefix[32];
    uint8_t rprefix[32];  // complement(prefix) (reverse complement as bytes, lib/kmers.js:31-38)
    // exact byte check of a window held in 8 words: (window ^ want[strand]) & mask[strand] must be 0
    // (strand 0: prefix at the start of the window, strand 1: complement(prefix) at its end)
    uint32_t want[2][8];
    uint32_t wmask[2][8];
    KjTable tab;
    KjIrrTable irr;
    KjOverflow ovf;
    uint64_t *status;     // per tile: flag << 62 | newline count; zeroed before every launch
    uint64_t *cand;       // filter path: candidate entries of this launch, 16 bytes each (kj_scan_warp.cuh)
    uint64_t cand_cap;    // in entries
    uint64_t *tile_cnt;   // filter path: '\n' per tile
    uint64_t *tile_excl;  // filter path: '\n' before the tile inside the launch (exclusive scan of tile_cnt)
    uint32_t n_fast;      // filter path: leading tiles that are whole, owned and readable through the tensor map
    uint32_t resolve_retry;   // kj_resolve_kernel: only the entries an earlier pass marked
    uint32_t c0a, c7f;    // 0x0A0A0A0A, 0x7F7F7F7F as kernel arguments: register operands of the newline test (kj_nl_flags_r)
    KjCounters *ctr;
};

// ----------------------------------------------------------------------------- helpers

__device__ __forceinline__ uint4 kj_ldg16(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
#else
    return *reinterpret_cast<const uint4 *>(p);
#endif
}

// 16 bytes at p, bytes at or beyond `limit` read as 0.  The byte-wise tail is kept out of line: it
// runs for the last chunk of a buffer only, and inlining it at every load site bloats the hot loops
// past the instruction cache.
static __device__ __noinline__ uint4 kj_load_chunk_tail(const uint8_t *buf, uint64_t off, uint64_t limit) {
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (uint32_t i = 0; i < 16; ++i)
        if (off + i < limit) w[i >> 2] |= (uint32_t)buf[off + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 kj_load_chunk(const uint8_t *buf, uint64_t off, uint64_t limit) {
    if (off + 16 <= limit) return kj_ldg16(buf + off);
    return kj_load_chunk_tail(buf, off, limit);
}

// per-tile state that outlives the code words: the filter kernel keeps two of these, so that the
// next tile can be converted (and its aggregate published) before the current one is finished
struct __align__(16) KjTileSmem {
    uint16_t nl[KJ_TILE_CHUNKS];          // '\n' mask of the chunk: together a bitmap of the tile, bit p = byte p
    uint32_t row_pre[KJ_ROWS];            // exclusive prefix of the rows' newline counts over the tile (kj_tile_rowscan_warp)
    uint32_t q_n;                         // entries in the kernel's candidate / line queue
    uint32_t tile_count;                  // '\n' in the tile
    unsigned long long excl_count;        // '\n' before the tile (whole stream)
};

// number of '\n' in the tile strictly before tile-relative position jt (jt <= KJ_TILE_BYTES).
// Only the rare consumers (candidates, line starts) need it, so the prefix inside the 32-chunk row
// is summed on demand instead of being scanned for every chunk in P1.
__device__ __forceinline__ uint32_t kj_count_before(const KjTileSmem &s, uint32_t jt) {
    const uint32_t c = jt >> 4;
    if (c >= KJ_TILE_CHUNKS) return s.tile_count;
    uint32_t cnt = s.row_pre[c >> 5] + __popc((uint32_t)s.nl[c] & ((1u << (jt & 15)) - 1u));
    for (uint32_t i = c & ~31u; i < c; ++i) cnt += __popc((uint32_t)s.nl[i]);
    return cnt;
}

// virtual offset of the first byte of the line that holds buffer position p (= offset + 1 of the
// last '\n' in [0, p)), searched backwards in global memory; falls back to the stream carry
static __device__ __noinline__ unsigned long long kj_line_start_global(const KjScanArgs &a, uint64_t p) {
    while (p > 0) {
        const uint64_t b = (p - 1) & ~15ull;
        const uint4 v = kj_load_chunk(a.buf, b, a.n);
        uint32_t m = kj_nl16(v.x, v.y, v.z, v.w);
        if (p - b < 16) m &= (1u << (uint32_t)(p - b)) - 1u;
        if (m) return a.voff + b + (31u - __clz(m)) + 1ull;
        p = b;
    }
    return a.ctr->carry_last[a.parity];
}

// same for tile-relative position jt, first in the tile's newline masks
__device__ __forceinline__ unsigned long long kj_line_start(const KjScanArgs &a, const KjTileSmem &s, uint32_t jt,
                                                            uint64_t tile_off, uint64_t tile_voff) {
    int c = (int)(jt >> 4);
    uint32_t m = (c < KJ_TILE_CHUNKS) ? ((uint32_t)s.nl[c] & ((1u << (jt & 15)) - 1u)) : 0u;
    if (c >= KJ_TILE_CHUNKS) c = KJ_TILE_CHUNKS;
    while (m == 0 && c > 0) { --c; m = s.nl[c]; }
    if (m) return tile_voff + (uint32_t)c * 16u + (31u - __clz(m)) + 1ull;
    return kj_line_start_global(a, tile_off);
}

// ----------------------------------------------------------------------------- P1 + look-back

// ---- asynchronous bulk copy global -> shared (TMA engine; SASS UBLKCP) completing on an mbarrier
#ifdef KJ_CPU_EMU   // tools/cuemu: mbarrier = {pending:16, count:16, phase:1}; the bulk copy happens at issue time
__device__ __forceinline__ void kj_bar_init(uint64_t *bar, uint32_t count) { *bar = (uint64_t)count | ((uint64_t)count << 16); }
__device__ __forceinline__ void kj_bar_arrive(uint64_t *bar) {
    uint64_t v = *bar;
    uint32_t pending = (uint32_t)(v & 0xFFFF) - 1, count = (uint32_t)((v >> 16) & 0xFFFF), phase = (uint32_t)(v >> 32);
    if (pending == 0) { pending = count; phase ^= 1u; }
    *bar = (uint64_t)pending | ((uint64_t)count << 16) | ((uint64_t)phase << 32);
}
__device__ __forceinline__ void kj_bar_wait(uint64_t *bar, uint32_t parity) {
    while ((uint32_t)(*reinterpret_cast<volatile uint64_t *>(bar) >> 32) == parity) emu_yield();
}
#else
__device__ __forceinline__ uint32_t kj_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void kj_bar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(kj_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void kj_bar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(kj_smem_u32(bar)) : "memory");
}
// returns once the phase with parity `parity` has completed.  try_wait suspends the warp in hardware
// for up to the hinted time, so waiting warps do not eat issue slots.  (A macro, so that profiles
// attribute the wait to the call site.)
#define kj_bar_wait(bar, parity)                                                                              \
    do {                                                                                                      \
        uint32_t ok__ = 0;                                                                                    \
        const uint32_t addr__ = kj_smem_u32(bar);                                                             \
        const uint32_t par__ = (parity);                                                                      \
        while (!ok__)                                                                                         \
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t" \
                         "selp.u32 %0, 1, 0, p;\n\t}"                                                         \
                         : "=r"(ok__) : "r"(addr__), "r"(par__), "r"(20000u) : "memory");                     \
    } while (0)
#endif

// one 16-byte chunk -> code word and newline mask (the rows are counted by kj_tile_rowscan_warp)
__device__ __forceinline__ void kj_p1_chunk(uint32_t *codes, KjTileSmem &s, uint32_t c, const uint4 v, uint32_t nl_keep) {
    codes[c] = kj_pack16(v.x, v.y, v.z, v.w);
    s.nl[c] = (uint16_t)(kj_nl16(v.x, v.y, v.z, v.w) & nl_keep);
}

// P1 straight from global memory (line kernel).
// Newlines are counted only inside the owned range.  The caller synchronises.
template <int NT>
static __device__ __noinline__ void kj_tile_p1_global(const KjScanArgs &a, uint32_t *codes, KjTileSmem &s, uint32_t tile,
                                                      uint32_t t) {
    constexpr int CPT = KJ_TILE_CHUNKS / NT;
    static_assert(CPT * NT == KJ_TILE_CHUNKS && NT % 32 == 0, "tile must split evenly into warp rows");
    const uint64_t tile_off = (uint64_t)tile * KJ_TILE_BYTES;
#pragma unroll 1
    for (int it = 0; it < CPT; ++it) {
        const uint32_t c = it * NT + t;
        const uint64_t off = tile_off + (uint64_t)c * 16u;
        const uint4 v = (off < a.n) ? kj_load_chunk(a.buf, off, a.n) : make_uint4(0, 0, 0, 0);
        uint32_t keep = 0xFFFFu;
        if (off >= a.own_n) keep = 0;
        else if (off + 16 > a.own_n) keep = (1u << (uint32_t)(a.own_n - off)) - 1u;
        kj_p1_chunk(codes, s, c, v, keep);
    }
    if (t < 2) {   // halo code words
        const uint64_t off = tile_off + (uint64_t)(KJ_TILE_CHUNKS + t) * 16u;
        const uint4 h = (off < a.n) ? kj_load_chunk(a.buf, off, a.n) : make_uint4(0, 0, 0, 0);
        codes[KJ_TILE_CHUNKS + t] = kj_pack16(h.x, h.y, h.z, h.w);
    }
}

// One warp turns the row counts into exclusive prefixes and publishes the tile's aggregate.
__device__ __forceinline__ void kj_tile_rowscan_warp(const KjScanArgs &a, KjTileSmem &s, uint32_t tile) {
    const uint32_t lane = threadIdx.x & 31;
    static_assert(KJ_ROWS <= 64, "two rows per lane");
    static_assert(KJ_ROWS % 2 == 0, "a lane counts a pair of rows");
    // newlines of rows 2 * lane and 2 * lane + 1, straight from the masks: a row is 32 masks = 16 words;
    // the lanes start at different words of their rows so that a load touches 16 banks, not one
    const uint32_t *w = reinterpret_cast<const uint32_t *>(s.nl) + 32u * lane;
    uint32_t c0 = 0, c1 = 0;
    if (2 * lane < KJ_ROWS) {
#pragma unroll
        for (uint32_t j = 0; j < 16; ++j) {
            const uint32_t i = (j + lane) & 15u;
            c0 += __popc(w[i]);
            c1 += __popc(w[16u + i]);
        }
    }
    const uint32_t sum = c0 + c1;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    const uint32_t excl = incl - sum;
    if (2 * lane < KJ_ROWS) s.row_pre[2 * lane] = excl;
    if (2 * lane + 1 < KJ_ROWS) s.row_pre[2 * lane + 1] = excl + c0;
    if (lane == 31) {
        s.tile_count = incl;
        *reinterpret_cast<volatile uint64_t *>(&a.status[tile]) = (KJ_ST_AGG << 62) | (uint64_t)incl;
    }
}

// One warp: decoupled look-back.  Writes excl_count, publishes the inclusive state and, for the last
// tile, the stream carry of the next launch.
__device__ __forceinline__ void kj_lookback(const KjScanArgs &a, KjTileSmem &s, uint32_t tile) {
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long ex_count = 0;
    int base = (int)tile - 1;
    bool done = false;
    while (!done) {
        const int idx = base - (int)lane;    // idx == -1 is the stream carry, below that: nothing
        uint64_t st;
        if (idx >= 0) {
            do {
                st = kj_ld_volatile(&a.status[idx]);
            } while ((st >> 62) == 0);
        } else if (idx == -1) {
            st = (KJ_ST_INC << 62) | (uint64_t)a.ctr->carry_lines[a.parity];
        } else {
            st = (KJ_ST_AGG << 62);          // neutral element
        }
        const uint32_t inc_mask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == KJ_ST_INC);
        const uint32_t upto = inc_mask ? (uint32_t)(__ffs(inc_mask) - 1) : 31u;   // lanes <= upto count
        unsigned long long cnt = (lane <= upto) ? (st & KJ_ST_MASK) : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        ex_count += cnt;
        if (inc_mask) done = true; else base -= 32;
    }
    if (lane == 0) {
        s.excl_count = ex_count;
        const unsigned long long inc_count = ex_count + s.tile_count;
        *reinterpret_cast<volatile uint64_t *>(&a.status[tile]) = (KJ_ST_INC << 62) | inc_count;
        if (tile == a.n_tiles - 1) {
            a.ctr->carry_lines[a.parity ^ 1] = inc_count;
            a.ctr->carry_last[a.parity ^ 1] = kj_line_start_global(a, a.own_n);
        }
    }
}

// Contribution of this tile's '\n' to the sum of sequence-line lengths (lines with index 1 mod 4):
// a '\n' that ends such a line adds its offset, one that ends the line before (index 0 mod 4)
// subtracts offset + 1.  The sums telescope over tiles and launches; the host adds the two ends
// of the stream (kj_counts_finish).  Optional (KJ_F_COUNT_BASES): it is a statistic the reference
// does not have, and it costs about as much as the prefix search itself.
__device__ __forceinline__ long long kj_tile_bases(const KjTileSmem &s, uint64_t tile_voff, uint32_t t, uint32_t nthreads) {
    long long sum = 0;
#pragma unroll 1
    for (uint32_t c = t; c < KJ_TILE_CHUNKS; c += nthreads) {
        uint32_t mk = s.nl[c];
        if (!mk) continue;
        uint32_t line = (uint32_t)s.excl_count + kj_count_before(s, c * 16u);    // mod 4 is all that matters
        const long long at = (long long)(tile_voff + (uint64_t)c * 16u);
        while (mk) {
            const uint32_t bit = __ffs(mk) - 1; mk &= mk - 1;
            const uint32_t ph = line & 3u;
            if (ph == 1u) sum += at + bit;
            else if (ph == 0u) sum -= at + bit + 1;
            ++line;
        }
    }
    return sum;
}

// ----------------------------------------------------------------------------- emit (verify + insert)

__device__ __forceinline__ void kj_spill(const KjScanArgs &a, uint64_t key, uint64_t ord) {
    unsigned long long i = atomicAdd(&a.ctr->n_overflow, 1ull);
    if (i < a.ovf.cap) { a.ovf.rec[2 * i] = key; a.ovf.rec[2 * i + 1] = ord; }
}
__device__ __forceinline__ void kj_spill_irr(const KjScanArgs &a, uint64_t off, uint32_t len,
                                             uint32_t strand, uint64_t ord) {
    unsigned long long i = atomicAdd(&a.ctr->n_irr_overflow, 1ull);
    if (i < a.ovf.irr_cap) {
        a.ovf.irr_rec[3 * i] = off;
        a.ovf.irr_rec[3 * i + 1] = ((uint64_t)len << 1) | strand;
        a.ovf.irr_rec[3 * i + 2] = ord;
    }
}

// key bytes of the window buf[off, off+len) on `strand` (1: complement(), i.e. reversed and
// A<->T, G<->C upper-case only) into key32 (zero padded)
__device__ __forceinline__ void kj_window_bytes(const uint8_t *buf, uint64_t off, uint32_t len,
                                                uint32_t strand, uint8_t *key32) {
    uint64_t *kw = reinterpret_cast<uint64_t *>(key32);
    kw[0] = kw[1] = kw[2] = kw[3] = 0;
    for (uint32_t t = 0; t < len; ++t)
        key32[t] = strand ? kj_comp_byte(buf[off + len - 1 - t]) : buf[off + t];
}

static __device__ __noinline__ void kj_emit_irregular(const KjScanArgs &a, uint64_t off, uint32_t len,
                                                      uint32_t strand, uint64_t ord) {
    __align__(8) uint8_t key32[32];
    kj_window_bytes(a.buf, off, len, strand, key32);
    if (!kj_insert_irr(a.irr, a.ctr, key32, len, ord, 1)) kj_spill_irr(a, off, len, strand, ord);
}

// the window's bytes: the (at most) three aligned 16-byte chunks that hold buf[j, j + k), requested together
__device__ __forceinline__ void kj_window_load(const KjScanArgs &a, uint64_t j, uint4 &v0, uint4 &v1, uint4 &v2) {
    const uint64_t base = j & ~15ull;
    v0 = kj_load_chunk(a.buf, base, a.n);
    v1 = make_uint4(0, 0, 0, 0);
    v2 = make_uint4(0, 0, 0, 0);
    if (base + 16u < j + a.k) v1 = kj_load_chunk(a.buf, base + 16u, a.n);
    if (base + 32u < j + a.k) v2 = kj_load_chunk(a.buf, base + 32u, a.n);
}

// ----------------------------------------------------------------------------- code-space filter

// 16 code lanes holding symbol i of complement(prefix) for the 16 windows of a chunk
template <int RC>
__device__ __forceinline__ uint32_t kj_rc_lanes(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t d) {
    if (RC == KJ_RC_LOW) return kj_funnel_r(c0, c1, 2u * d);
    if (RC == KJ_RC_HIGH) return kj_funnel_r(c1, c2, 2u * (d - 16u));
    return kj_lanes(c0, c1, c2, d);
}

// ----------------------------------------------------------------------------- dense kernel

// Empty prefix, step == 1, k >= 2: every window of a sequence line is an emission on both strands
// (BASELINE config 5), so there is nothing to filter and the hash table is the bound.  Same tiles,
// tickets and look-back as the other kernels; per chunk a thread walks its 16 window starts with
// the 2-bit codes, a newline bitmap and a "not A/C/G/T" bitmap of the chunk and its 32-byte halo in
// registers: a window is regular iff no bit of either bitmap falls inside it, its key is a shift of
// the code words away.  Windows with another byte (N, lower case, ...) take the byte-string side path.
struct KjDenseSmem {
    uint16_t nlt[KJ_TILE_CHUNKS + 2];     // '\n' bitmap incl. halo, NOT clipped to the owned range (window validity)
    uint16_t bad[KJ_TILE_CHUNKS + 2];     // bytes that are not A/C/G/T, incl. halo ('\n' is one of them)
    uint16_t pre[KJ_TILE_CHUNKS];         // counted '\n' before the chunk inside its row
};

__global__ void __launch_bounds__(KJ_THREADS)
kj_scan_dense_kernel(const __grid_constant__ KjScanArgs a) {
    __shared__ KjTileSmem s;
    __shared__ KjDenseSmem ds;
    __shared__ uint32_t codes[KJ_TILE_CHUNKS + 2];
    __shared__ uint32_t cur_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t k = a.k;
    const uint64_t kbits = k == 32 ? 0xFFFFFFFFull : ((1ull << k) - 1ull);           // k window bytes
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);             // 2k key bits
    uint32_t n_emit = 0;
    long long n_bases = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) cur_tile = atomicAdd(&a.ctr->ticket, 1u);
        __syncthreads();
        const uint32_t tile = cur_tile;
        if (tile >= a.n_tiles) break;
        const uint64_t tile_off = (uint64_t)tile * KJ_TILE_BYTES;
        const uint64_t tile_voff = a.voff + tile_off;
        const uint32_t own_in_tile =
            (a.own_n - tile_off < KJ_TILE_BYTES) ? (uint32_t)(a.own_n - tile_off) : KJ_TILE_BYTES;

        // P1: codes, both bitmaps, counted newlines with their prefix inside the row
#pragma unroll 1
        for (int it = 0; it < KJ_TILE_CHUNKS / KJ_THREADS; ++it) {
            const uint32_t c = it * KJ_THREADS + tid;
            const uint64_t off = tile_off + (uint64_t)c * 16u;
            const uint4 v = (off < a.n) ? kj_load_chunk(a.buf, off, a.n) : make_uint4(0, 0, 0, 0);
            uint32_t keep = 0xFFFFu;
            if (off >= a.own_n) keep = 0;
            else if (off + 16 > a.own_n) keep = (1u << (uint32_t)(a.own_n - off)) - 1u;
            codes[c] = kj_pack16(v.x, v.y, v.z, v.w);
            const uint32_t nlt = kj_nl16(v.x, v.y, v.z, v.w);
            ds.nlt[c] = (uint16_t)nlt;
            ds.bad[c] = (uint16_t)kj_bad16(v.x, v.y, v.z, v.w);
            const uint32_t nl = nlt & keep;
            s.nl[c] = (uint16_t)nl;
            const uint32_t cnt = __popc(nl);
            uint32_t incl = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((int)lane >= d) incl += o;
            }
            ds.pre[c] = (uint16_t)(incl - cnt);
        }
        if (tid < 2) {   // halo
            const uint64_t off = tile_off + (uint64_t)(KJ_TILE_CHUNKS + tid) * 16u;
            const uint4 h = (off < a.n) ? kj_load_chunk(a.buf, off, a.n) : make_uint4(0, 0, 0, 0);
            codes[KJ_TILE_CHUNKS + tid] = kj_pack16(h.x, h.y, h.z, h.w);
            ds.nlt[KJ_TILE_CHUNKS + tid] = (uint16_t)kj_nl16(h.x, h.y, h.z, h.w);
            ds.bad[KJ_TILE_CHUNKS + tid] = (uint16_t)kj_bad16(h.x, h.y, h.z, h.w);
        }
        __syncthreads();
        if (warp == 0) { kj_tile_rowscan_warp(a, s, tile); kj_lookback(a, s, tile); }
        __syncthreads();
        if (a.count_bases) n_bases += kj_tile_bases(s, tile_voff, tid, KJ_THREADS);

        // P2: every owned window start
#pragma unroll 1
        for (int it = 0; it < KJ_TILE_CHUNKS / KJ_THREADS; ++it) {
            const uint32_t c = it * KJ_THREADS + tid;
            const uint32_t pos0 = c * 16u;
            if (pos0 >= own_in_tile) continue;
            const uint32_t n_own = own_in_tile - pos0 < 16u ? own_in_tile - pos0 : 16u;
            const uint64_t NLT = (uint64_t)ds.nlt[c] | ((uint64_t)ds.nlt[c + 1] << 16) | ((uint64_t)ds.nlt[c + 2] << 32);
            const uint64_t BAD = (uint64_t)ds.bad[c] | ((uint64_t)ds.bad[c + 1] << 16) | ((uint64_t)ds.bad[c + 2] << 32);
            const uint64_t CLO = (uint64_t)codes[c] | ((uint64_t)codes[c + 1] << 32);
            const uint64_t CHI = codes[c + 2];
            const uint32_t nlm = s.nl[c];
            const uint64_t line0 = s.excl_count + s.row_pre[c >> 5] + ds.pre[c];
            unsigned long long start0 = 0;
            bool have_start0 = false;
#pragma unroll 1
            for (uint32_t p = 0; p < n_own; ++p) {
                const uint64_t j = tile_off + pos0 + p;
                if (j + k > a.n) break;                                // window must lie inside the stream
                const uint32_t below = nlm & ((1u << p) - 1u);
                const uint64_t line = line0 + __popc(below);
                if ((line & 3ull) != 1ull) continue;                   // lib/kmers.js:151  i === 1
                if ((NLT >> p) & kbits) continue;                      // crosses the end of the line
                uint64_t ord_f = 0, ord_r = 0;
                if (a.order) {
                    unsigned long long start;
                    if (below) start = tile_voff + pos0 + (31u - __clz(below)) + 1ull;
                    else {
                        if (!have_start0) { start0 = kj_line_start(a, s, pos0, tile_off, tile_voff); have_start0 = true; }
                        start = start0;
                    }
                    const uint64_t col = tile_voff + pos0 + p - start;
                    const uint64_t read_idx = line >> 2;
                    if (col > KJ_POS_MAX) { atomicOr(&a.ctr->error_flags, KJ_DEV_E_LINE_TOO_LONG); continue; }
                    if (read_idx >> 36) { atomicOr(&a.ctr->error_flags, KJ_DEV_E_READS_OVERFLOW); continue; }
                    ord_f = kj_ordinal(read_idx, 0, col);
                    ord_r = kj_ordinal(read_idx, 1, KJ_POS_MAX - col);
                }
                if ((BAD >> p) & kbits) {                              // some other byte inside: byte-string keys
                    kj_emit_irregular(a, j, k, 0, ord_f);
                    ++n_emit;
                    if (a.n_strands > 1) { kj_emit_irregular(a, j, k, 1, ord_r); ++n_emit; }
                    continue;
                }
                const uint64_t P = ((CLO >> (2u * p)) | (p ? CHI << (64u - 2u * p) : 0ull)) & kmask;   // code of window byte i at bits 2i
                const uint64_t fk = kj_pairrev64(P) >> (64u - 2u * k);
                if (!kj_insert(a.tab, a.ctr, fk, ord_f, 1)) kj_spill(a, fk, ord_f);
                ++n_emit;
                if (a.n_strands > 1) {
                    const uint64_t rk = (P ^ 0xAAAAAAAAAAAAAAAAull) & kmask;
                    if (!kj_insert(a.tab, a.ctr, rk, ord_r, 1)) kj_spill(a, rk, ord_r);
                    ++n_emit;
                }
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {       // one atomic per warp
        n_emit += __shfl_xor_sync(0xFFFFFFFFu, n_emit, d);
        n_bases += __shfl_xor_sync(0xFFFFFFFFu, n_bases, d);
    }
    if ((tid & 31) == 0) {
        if (n_emit) atomicAdd(&a.ctr->n_occ, (unsigned long long)n_emit);
        if (n_bases) atomicAdd(&a.ctr->n_bases, (unsigned long long)n_bases);
    }
}

// ----------------------------------------------------------------------------- line-oriented kernel

// One whole sequence line starting at buffer offset ls with line index `line`; `width` threads
// (a warp, lane = 0..31) stride over the window index exactly as lib/kmers.js:88-100 does.
__device__ __forceinline__ void kj_process_line(const KjScanArgs &a, uint64_t ls, uint64_t line,
                                                uint32_t lane, uint32_t &n_emit) {
    // line length: distance to the next '\n' or to the end of the stream
    uint64_t L = 0;
    bool found = false;
    for (uint64_t base = 0; !found; base += 32) {
        uint64_t p = ls + base + lane;
        bool hit = (p >= a.n) || a.buf[p] == '\n';
        uint32_t b = __ballot_sync(0xFFFFFFFFu, hit);
        if (b) { L = base + (uint32_t)(__ffs(b) - 1); found = true; }
    }
    if (ls + L >= a.n && !a.final_) {     // ran into the end of the readable bytes, not a '\n'
        if (lane == 0) atomicOr(&a.ctr->error_flags, KJ_DEV_E_LINE_EXCEEDS_HALO);
        return;
    }
    if (lane == 0 && L) atomicAdd(&a.ctr->n_bases, (unsigned long long)L);   // every sequence line counts
    if (L <= 1 && a.line_gate) return;                    // lib/kmers.js:151
    if (L == 0) return;
    const uint64_t k = a.k, step = a.step, m = a.m;
    if (L < k) return;                                    // stop < 0: no iterations
    const uint64_t stop = L - k;
    const uint64_t read_idx = line >> 2;
    if ((read_idx >> 36) || stop > KJ_POS_MAX) {
        if (lane == 0) atomicOr(&a.ctr->error_flags,
                                (read_idx >> 36) ? KJ_DEV_E_READS_OVERFLOW : KJ_DEV_E_LINE_TOO_LONG);
        return;
    }
    for (uint32_t strand = 0; strand < a.n_strands; ++strand) {
        for (uint64_t i = lane; i <= stop; i += 32) {
            const uint64_t ini = i * step;
            const uint64_t x0 = ini < L ? ini : L;
            const uint64_t x1 = ini + k < L ? ini + k : L;
            const uint32_t len = (uint32_t)(x1 - x0);     // substring clips (lib/kmers.js:92)
            if (len < m) continue;
            // the window covers positions [x0, x1) of line (strand 0) or of complement(line)
            // (strand 1); position t of complement(line) is comp(line[L-1-t])
            uint64_t key = 0;
            bool regular = (len == k), pass = true;
            for (uint32_t t = 0; t < len; ++t) {
                uint32_t c = strand ? kj_comp_byte(a.buf[ls + (L - 1 - (x0 + t))])
                                    : a.buf[ls + x0 + t];
                if (t < m && c != a.prefix[t]) { pass = false; break; }
                regular = regular && kj_is_acgt(c);
                key = (key << 2) | ((c >> 1) & 3u);
            }
            if (!pass) continue;
            const uint64_t ord = a.order ? kj_ordinal(read_idx, strand, i) : 0;
            ++n_emit;
            if (regular) {
                if (!kj_insert(a.tab, a.ctr, key, ord, 1)) kj_spill(a, key, ord);
            } else {
                // forward-strand byte range that holds this window
                uint64_t off = strand ? ls + (L - x1) : ls + x0;
                kj_emit_irregular(a, off, len, strand, ord);
            }
        }
    }
}

__global__ void __launch_bounds__(KJ_THREADS)
kj_scan_lines_kernel(const __grid_constant__ KjScanArgs a) {
    __shared__ KjTileSmem s;
    __shared__ uint32_t codes[KJ_TILE_CHUNKS + 2];
    __shared__ uint32_t queue[KJ_QCAP];
    __shared__ uint32_t cur_tile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t n_emit = 0;
    for (;;) {
        __syncthreads();
        if (tid == 0) { cur_tile = atomicAdd(&a.ctr->ticket, 1u); s.q_n = 0; }
        __syncthreads();
        const uint32_t tile = cur_tile;
        if (tile >= a.n_tiles) break;
        const uint64_t tile_off = (uint64_t)tile * KJ_TILE_BYTES;

        kj_tile_p1_global<KJ_THREADS>(a, codes, s, tile, tid);
        __syncthreads();
        if (warp == 0) { kj_tile_rowscan_warp(a, s, tile); kj_lookback(a, s, tile); }
        __syncthreads();

        // the line that starts exactly at the first byte of the stream piece: owned iff the
        // previous stream byte was a '\n' (or there is no previous byte at all)
        if (tile == 0 && tid == 0 && a.ctr->carry_last[a.parity] == a.voff && (s.excl_count & 3ull) == 1ull &&
            a.own_n > 0)
            queue[s.q_n++] = 0;
        // P2': every '\n' at e starts a line at e+1 (owned by the tile that holds the '\n')
        for (int it = 0; it < KJ_TILE_CHUNKS / KJ_THREADS; ++it) {
            __syncthreads();
            const uint32_t c = it * KJ_THREADS + tid;
            uint32_t mk = s.nl[c];
            if (mk) {
                uint32_t before = kj_count_before(s, c * 16u);
                while (mk) {
                    uint32_t bit = __ffs(mk) - 1; mk &= mk - 1;
                    ++before;                                  // index of the line after this '\n'
                    unsigned long long line = s.excl_count + before;
                    uint64_t start = tile_off + c * 16u + bit + 1u;
                    if ((line & 3ull) == 1ull && start < a.n) {
                        uint32_t q = atomicAdd(&s.q_n, 1u);
                        queue[q] = c * 16u + bit + 1u;       // at most 4096 / 4 entries per round
                    }
                }
            }
            __syncthreads();
            // P3': one warp per line
            const uint32_t qn = s.q_n;
            for (uint32_t q = warp; q < qn; q += KJ_THREADS / 32) {
                uint32_t st = queue[q];
                unsigned long long line = s.excl_count + kj_count_before(s, st);
                kj_process_line(a, tile_off + st, line, lane, n_emit);
            }
            __syncthreads();
            if (tid == 0) s.q_n = 0;
        }
    }
    if (n_emit) atomicAdd(&a.ctr->n_occ, (unsigned long long)n_emit);
}
