// kj_count.cu -- host side of the extraction + count path (kj_counts_* of include/kmerjs_b200.h)
// and the small maintenance kernels around the table (rehash, replay, compaction, exchange).
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <future>
#include <thread>
#include <zlib.h>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include "kj_internal.hpp"
#include "kj_scan_warp.cuh"
#ifndef KJ_CPU_EMU
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#endif

// the filter-path piece in flight
struct KjPiece {
    KjScanArgs args;
    KjTensorMap tmap;
    bool timed = false;
};

// ------------------------------------------------------------------------------------ kernels

__global__ void kj_rehash_kernel(KjTable oldt, uint64_t old_cap, KjTable newt, KjCounters *ctr) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_cap;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = oldt.keys[i];
        if (key == KJ_EMPTY) continue;
        uint64_t ord = oldt.ords ? oldt.ords[i] : ~0ull;
        if (!kj_insert(newt, ctr, key, ord, oldt.counts[i])) atomicOr(&ctr->error_flags, 0x80000000u);
    }
}

__global__ void kj_rehash_irr_kernel(KjIrrTable oldt, uint64_t old_cap, KjIrrTable newt,
                                     KjCounters *ctr) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < old_cap;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t st = oldt.state[i];
        if (st < 2) continue;
        __align__(8) uint8_t key32[32];
        const uint64_t *src = reinterpret_cast<const uint64_t *>(oldt.keys + i * 32);
        uint64_t *dst = reinterpret_cast<uint64_t *>(key32);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        if (!kj_insert_irr(newt, ctr, key32, st - 2, oldt.ords[i], oldt.counts[i]))
            atomicOr(&ctr->error_flags, 0x80000000u);
    }
}

// re-insert spilled regular emissions {key, ord}
__global__ void kj_replay_kernel(KjTable t, KjCounters *ctr, const uint64_t *rec, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        if (!kj_insert(t, ctr, rec[2 * i], rec[2 * i + 1], 1)) atomicOr(&ctr->error_flags, 0x80000000u);
}

// re-insert spilled irregular emissions {buffer offset, len<<1|strand, ord}
__global__ void kj_replay_irr_kernel(KjIrrTable t, KjCounters *ctr, const uint8_t *buf,
                                     const uint64_t *rec, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        __align__(8) uint8_t key32[32];
        uint32_t len = (uint32_t)(rec[3 * i + 1] >> 1), strand = (uint32_t)(rec[3 * i + 1] & 1);
        kj_window_bytes(buf, rec[3 * i], len, strand, key32);
        if (!kj_insert_irr(t, ctr, key32, len, rec[3 * i + 2], 1))
            atomicOr(&ctr->error_flags, 0x80000000u);
    }
}

// one atomic per warp: ballot the occupied slots, lane 0 claims a run of output positions
__global__ void kj_compact_kernel(KjTable t, uint64_t cap, KjCounters *ctr, uint64_t *keys,
                                  uint64_t *counts, uint64_t *ords, uint64_t out_cap) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t rounds = (cap + stride - 1) / stride;       // same trip count for every thread
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t r = 0; r < rounds; ++r, i += stride) {
        uint64_t key = (i < cap) ? t.keys[i] : KJ_EMPTY;
        bool have = key != KJ_EMPTY;
        uint32_t b = __ballot_sync(0xFFFFFFFFu, have);
        if (!b) continue;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctr->n_compact, (unsigned long long)__popc(b));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (have) {
            uint64_t o = base + __popc(b & ((1u << lane) - 1u));
            if (o >= out_cap) continue;                  // sized by a hint that was too small: the host sees n_compact and repeats
            keys[o] = key;
            counts[o] = t.counts[i];
            ords[o] = t.ords ? t.ords[i] : ~0ull;
        }
    }
}

__global__ void kj_compact_irr_kernel(KjIrrTable t, uint64_t cap, KjCounters *ctr,
                                      KjIrrRecord *out, uint64_t out_cap) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t st = t.state[i];
        if (st < 2) continue;
        unsigned long long o = atomicAdd(&ctr->n_irr_compact, 1ull);
        if (o >= out_cap) continue;
        const uint64_t *src = reinterpret_cast<const uint64_t *>(t.keys + i * 32);
        uint64_t *dst = reinterpret_cast<uint64_t *>(out[o].key);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        out[o].len = st - 2;
        out[o].count = t.counts[i];
        out[o].ord = t.ords[i];
    }
}

// newline statistics of a byte range (sharded ingest: a rank needs the number of '\n' before its
// range, lib/kmers.js:151-163 is "line index mod 4")
__global__ void kj_newline_kernel(const uint8_t *buf, uint64_t n, unsigned long long *out /* count, last+1 */) {
    unsigned long long cnt = 0, last = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * 16;
    for (uint64_t off = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 16; off < n; off += stride) {
        uint4 v = kj_load_chunk(buf, off, n);
        uint32_t m = kj_nl16(v.x, v.y, v.z, v.w);
        if (off + 16 > n) m &= (1u << (uint32_t)(n - off)) - 1u;
        if (m) { cnt += __popc(m); last = off + (31 - __clz(m)) + 1; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, d);
        unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, last, d);
        last = o > last ? o : last;
    }
    if ((threadIdx.x & 31) == 0) {
        if (cnt) atomicAdd(&out[0], cnt);
        if (last) atomicMax(&out[1], last);
    }
}

__global__ void kj_merge_records_kernel(KjTable t, KjCounters *ctr, const KjRecord *rec, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        if (!kj_insert(t, ctr, rec[i].key, rec[i].ord, rec[i].count))
            atomicOr(&ctr->error_flags, 0x80000000u);
}

__global__ void kj_merge_irr_kernel(KjIrrTable t, KjCounters *ctr, const KjIrrRecord *rec, uint64_t n, uint32_t part,
                                    uint32_t n_parts) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        if (n_parts > 1 && kj_owner_bytes(rec[i].key, (uint32_t)rec[i].len, n_parts) != part) continue;   // not ours
        __align__(8) uint8_t key32[32];
        const uint64_t *src = reinterpret_cast<const uint64_t *>(rec[i].key);
        uint64_t *dst = reinterpret_cast<uint64_t *>(key32);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        if (!kj_insert_irr(t, ctr, key32, (uint32_t)rec[i].len, rec[i].ord, rec[i].count))
            atomicOr(&ctr->error_flags, 0x80000000u);
    }
}

// export, on the device: entries in first-insertion order (sorted by first-seen ordinal) -> 32-byte key
// strings (zero padded), key lengths and counts; irregular entries report their position instead (their
// byte-string keys live on the host)
__global__ void kj_export_iota_kernel(uint32_t *idx, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        idx[i] = (uint32_t)i;
}
__global__ void kj_export_gather_kernel(const uint32_t *perm, const uint64_t *keys, const uint64_t *counts, uint64_t q,
                                        uint64_t n_reg, uint32_t k, uint4 *out_keys, uint32_t *out_len,
                                        uint64_t *out_counts, uint64_t *irr_pos) {
    for (uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; o < q; o += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t i = perm[o];
        uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (i < n_reg) {
            const uint64_t key = keys[i];
            for (uint32_t j = 0; j < k; ++j) {
                const uint32_t code = (uint32_t)(key >> (2 * (k - 1 - j))) & 3u;       // A C T G
                w[j >> 2] |= ((0x47544341u >> (8 * code)) & 0xFFu) << (8 * (j & 3));
            }
        } else {
            irr_pos[i - n_reg] = o;
        }
        out_keys[2 * o] = make_uint4(w[0], w[1], w[2], w[3]);
        out_keys[2 * o + 1] = make_uint4(w[4], w[5], w[6], w[7]);
        out_len[o] = k;
        out_counts[o] = counts[i];
    }
}

// exchange: histogram of owners, then scatter into owner-grouped records
__global__ void kj_part_hist_kernel(const uint64_t *keys, uint64_t n, uint32_t n_parts,
                                    unsigned long long *hist) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&hist[kj_owner_key(keys[i], n_parts)], 1ull);
}
__global__ void kj_part_scatter_kernel(const uint64_t *keys, const uint64_t *counts,
                                       const uint64_t *ords, uint64_t n, uint32_t n_parts,
                                       unsigned long long *cursor, KjRecord *out) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long o = atomicAdd(&cursor[kj_owner_key(keys[i], n_parts)], 1ull);
        out[o].key = keys[i];
        out[o].count = counts[i];
        out[o].ord = ords[i];
    }
}

// ---- fixed-capacity exchange (no size round trip, no host synchronisation) ---------------------------------------
// A segment = {KjSegHeader, KjRecord[cap_reg], KjIrrRecord[cap_irr]}, one per destination rank.
struct KjSegHeader {
    unsigned long long n_reg, n_irr;               // records the sender had for this owner (may exceed the capacity: overflow)
    unsigned long long lines, bases, occ, bytes;   // the sender's totals (the same in all of its segments)
    unsigned long long flags, pad_;
};
static_assert(sizeof(KjSegHeader) == 64, "segment header");
struct KjSegArgs {
    uint8_t *seg;            // n_parts segments
    uint64_t seg_bytes;
    uint32_t n_parts, cap_reg, cap_irr;
    uint32_t parity;
    uint64_t voff, consumed;           // stream position of the sender (host knowledge)
    long long bases_fix;               // closes the telescoping sum of KJ_F_COUNT_BASES at the start of the range
    uint32_t count_bases, bases_tail;  // bases_tail: the range may end inside a sequence line (adds voff)
    uint64_t cand_cap;
};
__device__ __forceinline__ KjSegHeader *kj_seg_hdr(const KjSegArgs &a, uint32_t p) {
    return reinterpret_cast<KjSegHeader *>(a.seg + (uint64_t)p * a.seg_bytes);
}

// the sender's totals into every header (the counters are final: this runs after the count kernels in stream order)
__global__ void kj_seg_totals_kernel(const KjSegArgs a, const KjCounters *ctr) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= a.n_parts) return;
    KjSegHeader *h = kj_seg_hdr(a, p);
    unsigned long long lines = ctr->carry_lines[a.parity];
    long long bases = (long long)ctr->n_bases;
    if (a.count_bases == 1 && a.consumed) {       // filter / dense kernels: close the telescoping sum at both ends
        bases += a.bases_fix;
        if (a.bases_tail && (lines & 3ull) == 1ull) bases += (long long)a.voff;
    }
    // a non-empty unterminated tail is one more line (lib/kmers.js:130-136)
    if (a.consumed && ctr->carry_last[a.parity] != a.voff) lines += 1;
    h->lines = lines;
    h->bases = a.count_bases ? (unsigned long long)bases : 0ull;
    h->occ = ctr->n_occ;
    h->bytes = a.consumed;
    h->flags = (ctr->n_overflow || ctr->n_irr_overflow || ctr->n_cand > a.cand_cap || ctr->error_flags) ? 1ull : 0ull;
    h->pad_ = 0;
}

__global__ void kj_seg_scatter_kernel(const KjSegArgs a, KjTable t, uint64_t cap, KjIrrTable it, uint64_t irr_cap,
                                      const KjCounters *ctr) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i = i0; i < cap; i += stride) {
        const uint64_t key = t.keys[i];
        if (key == KJ_EMPTY) continue;
        const uint32_t p = kj_owner_key(key, a.n_parts);
        KjSegHeader *h = kj_seg_hdr(a, p);
        const unsigned long long o = atomicAdd(&h->n_reg, 1ull);
        if (o < a.cap_reg) {
            KjRecord *r = reinterpret_cast<KjRecord *>(reinterpret_cast<uint8_t *>(h) + sizeof(KjSegHeader)) + o;
            r->key = key; r->count = t.counts[i]; r->ord = t.ords ? t.ords[i] : ~0ull;
        }
    }
    if (i0 == 0 && ctr->special_count) {          // the one key equal to KJ_EMPTY (k = 32, all 'G')
        const uint32_t p = kj_owner_key(KJ_EMPTY, a.n_parts);
        KjSegHeader *h = kj_seg_hdr(a, p);
        const unsigned long long o = atomicAdd(&h->n_reg, 1ull);
        if (o < a.cap_reg) {
            KjRecord *r = reinterpret_cast<KjRecord *>(reinterpret_cast<uint8_t *>(h) + sizeof(KjSegHeader)) + o;
            r->key = KJ_EMPTY; r->count = ctr->special_count; r->ord = ctr->special_ord;
        }
    }
    for (uint64_t i = i0; i < irr_cap; i += stride) {
        const uint32_t st = it.state[i];
        if (st < 2) continue;
        const uint32_t p = kj_owner_bytes(it.keys + i * 32, st - 2, a.n_parts);
        KjSegHeader *h = kj_seg_hdr(a, p);
        const unsigned long long o = atomicAdd(&h->n_irr, 1ull);
        if (o < a.cap_irr) {
            KjIrrRecord *r = reinterpret_cast<KjIrrRecord *>(reinterpret_cast<uint8_t *>(h) + sizeof(KjSegHeader) +
                                                             (uint64_t)a.cap_reg * sizeof(KjRecord)) + o;
            const uint64_t *src = reinterpret_cast<const uint64_t *>(it.keys + i * 32);
            uint64_t *dst = reinterpret_cast<uint64_t *>(r->key);
            dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
            r->len = st - 2; r->count = it.counts[i]; r->ord = it.ords[i];
        }
    }
}

// the receiver: every record of every segment into the owner's tables, the totals into the counters
__global__ void kj_seg_merge_kernel(const KjSegArgs a, KjTable t, KjIrrTable it, KjCounters *ctr) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (uint64_t i = i0; i < (uint64_t)a.n_parts * a.cap_reg; i += stride) {
        const uint32_t p = (uint32_t)(i / a.cap_reg), j = (uint32_t)(i % a.cap_reg);
        const KjSegHeader *h = kj_seg_hdr(a, p);
        if (j >= h->n_reg) continue;
        const KjRecord *r = reinterpret_cast<const KjRecord *>(reinterpret_cast<const uint8_t *>(h) + sizeof(KjSegHeader)) + j;
        if (!kj_insert(t, ctr, r->key, r->ord, r->count)) atomicOr(&ctr->error_flags, 0x80000000u);
    }
    for (uint64_t i = i0; i < (uint64_t)a.n_parts * a.cap_irr; i += stride) {
        const uint32_t p = (uint32_t)(i / a.cap_irr), j = (uint32_t)(i % a.cap_irr);
        const KjSegHeader *h = kj_seg_hdr(a, p);
        if (j >= h->n_irr) continue;
        const KjIrrRecord *r = reinterpret_cast<const KjIrrRecord *>(reinterpret_cast<const uint8_t *>(h) + sizeof(KjSegHeader) +
                                                                     (uint64_t)a.cap_reg * sizeof(KjRecord)) + j;
        __align__(8) uint8_t key32[32];
        const uint64_t *src = reinterpret_cast<const uint64_t *>(r->key);
        uint64_t *dst = reinterpret_cast<uint64_t *>(key32);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = src[3];
        if (!kj_insert_irr(it, ctr, key32, (uint32_t)r->len, r->ord, r->count)) atomicOr(&ctr->error_flags, 0x80000000u);
    }
    if (i0 < a.n_parts) {
        const KjSegHeader *h = kj_seg_hdr(a, (uint32_t)i0);
        atomicMax(&ctr->x_lines, h->lines);       // a rank's line count includes the lines before its range: the last rank's is the file's
        atomicAdd(&ctr->x_bases, h->bases);
        atomicAdd(&ctr->x_occ, h->occ);
        atomicAdd(&ctr->x_bytes, h->bytes);
        if (h->n_reg > a.cap_reg || h->n_irr > a.cap_irr) atomicOr(&ctr->error_flags, KJ_DEV_E_XCHG_OVERFLOW);
        if (h->flags) atomicOr(&ctr->error_flags, KJ_DEV_E_XCHG_INCOMPLETE);
    }
}

// ------------------------------------------------------------------------------------ helpers

static uint64_t next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

int kj_grid_for(const kj_ctx *ctx, uint64_t n, int threads) {
    uint64_t g = (n + threads - 1) / threads;
    uint64_t mx = (uint64_t)ctx->sm_count * 8;
    return (int)std::max<uint64_t>(1, std::min(g, mx));
}
static int grid_for(const kj_ctx *ctx, uint64_t n, int threads = 256) { return kj_grid_for(ctx, n, threads); }

static void free_table(kj_ctx *ctx, KjTable &t) {
    kj_dfree(ctx, t.keys); kj_dfree(ctx, t.counts); kj_dfree(ctx, t.ords);
    t = KjTable{};
}
static void free_irr(kj_ctx *ctx, KjIrrTable &t) {
    kj_dfree(ctx, t.state); kj_dfree(ctx, t.keys); kj_dfree(ctx, t.counts); kj_dfree(ctx, t.ords);
    t = KjIrrTable{};
}

static int alloc_table(kj_counts *c, uint64_t cap, KjTable *out) {
    kj_ctx *ctx = c->ctx;
    KjTable t{};
    cudaError_t e = kj_dmalloc(ctx, &t.keys, cap * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &t.counts, cap * 8);
    if (e == cudaSuccess && c->order) e = kj_dmalloc(ctx, &t.ords, cap * 8);
    if (e != cudaSuccess) {
        free_table(ctx, t);
        cudaGetLastError();
        return kj_fail(ctx, KJ_E_TABLE_FULL, "k-mer table of " + std::to_string(cap) +
                                                 " slots does not fit in device memory");
    }
    t.mask = cap - 1;
    KJ_CUDA(ctx, cudaMemsetAsync(t.keys, 0xFF, cap * 8, ctx->stream));
    KJ_CUDA(ctx, cudaMemsetAsync(t.counts, 0, cap * 8, ctx->stream));
    if (t.ords) KJ_CUDA(ctx, cudaMemsetAsync(t.ords, 0xFF, cap * 8, ctx->stream));
    *out = t;
    return KJ_OK;
}
static int alloc_irr(kj_counts *c, uint64_t cap, KjIrrTable *out) {
    kj_ctx *ctx = c->ctx;
    KjIrrTable t{};
    KJ_CUDA(ctx, kj_dmalloc(ctx, &t.state, cap * 4));
    KJ_CUDA(ctx, kj_dmalloc(ctx, &t.keys, cap * 32));
    KJ_CUDA(ctx, kj_dmalloc(ctx, &t.counts, cap * 8));
    KJ_CUDA(ctx, kj_dmalloc(ctx, &t.ords, cap * 8));
    t.mask = cap - 1;
    KJ_CUDA(ctx, cudaMemsetAsync(t.state, 0, cap * 4, ctx->stream));
    KJ_CUDA(ctx, cudaMemsetAsync(t.counts, 0, cap * 8, ctx->stream));
    KJ_CUDA(ctx, cudaMemsetAsync(t.ords, 0xFF, cap * 8, ctx->stream));
    *out = t;
    return KJ_OK;
}

static int pull_counters(kj_counts *c) {
    kj_ctx *ctx = c->ctx;
    KJ_CUDA(ctx, cudaMemcpyAsync(c->h_ctr, c->ctr, sizeof(KjCounters), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long nu = 0;
    for (int i = 0; i < 64; ++i) nu += c->h_ctr->n_unique_part[i];
    c->h_ctr->n_unique = nu;
    c->irr_bound = c->h_ctr->n_irr_unique;
    return KJ_OK;
}

// make the regular table hold at least `want` slots (power of two); rehash when it exists
static int grow_table(kj_counts *c, uint64_t want) {
    kj_ctx *ctx = c->ctx;
    want = next_pow2(std::max<uint64_t>(want, 1ull << 12));
    if (want <= c->cap) return KJ_OK;
    KjTable nt{};
    int rc = alloc_table(c, want, &nt);
    if (rc) return rc;
    if (c->cap) {
        KJ_CUDA(ctx, cudaMemsetAsync(c->ctr->n_unique_part, 0, sizeof(c->ctr->n_unique_part), ctx->stream));
        KJ_LAUNCH(kj_rehash_kernel, grid_for(ctx, c->cap), 256, 0, ctx->stream, c->tab, c->cap, nt, c->ctr);
        ctx->launches++;
        free_table(ctx, c->tab);
    }
    c->tab = nt;
    c->cap = want;
    return KJ_OK;
}

static int grow_irr(kj_counts *c, uint64_t want) {
    kj_ctx *ctx = c->ctx;
    want = next_pow2(std::max<uint64_t>(want, 1ull << 10));
    if (want <= c->irr_cap) return KJ_OK;
    KjIrrTable nt{};
    int rc = alloc_irr(c, want, &nt);
    if (rc) return rc;
    if (c->irr_cap) {
        KJ_CUDA(ctx, cudaMemsetAsync(&c->ctr->n_irr_unique, 0, sizeof(unsigned long long), ctx->stream));
        KJ_LAUNCH(kj_rehash_irr_kernel, grid_for(ctx, c->irr_cap), 256, 0, ctx->stream, c->irr, c->irr_cap,
                  nt, c->ctr);
        ctx->launches++;
        free_irr(ctx, c->irr);
    }
    c->irr = nt;
    c->irr_cap = want;
    return KJ_OK;
}

static int ensure_overflow(kj_counts *c, uint64_t reg_cap, uint64_t irr_cap) {
    kj_ctx *ctx = c->ctx;
    if (reg_cap > c->ovf.cap) {
        kj_dfree(ctx, c->ovf.rec);
        c->ovf.rec = nullptr; c->ovf.cap = 0;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->ovf.rec, reg_cap * 16));
        c->ovf.cap = reg_cap;
    }
    if (irr_cap > c->ovf.irr_cap) {
        kj_dfree(ctx, c->ovf.irr_rec);
        c->ovf.irr_rec = nullptr; c->ovf.irr_cap = 0;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->ovf.irr_rec, irr_cap * 24));
        c->ovf.irr_cap = irr_cap;
    }
    return KJ_OK;
}

static int check_device_errors(kj_counts *c) {
    uint32_t f = c->h_ctr->error_flags;
    if (!f) return KJ_OK;
    if (f & 0x80000000u) return kj_fail(c->ctx, KJ_E_TABLE_FULL, "k-mer table probe limit reached while rehashing/merging");
    if (f & KJ_DEV_E_LINE_EXCEEDS_HALO)
        return kj_fail(c->ctx, KJ_E_RANGE, "a sequence line extends beyond the halo of its buffer (the line-oriented kernel needs whole lines; give a larger halo)");
    if (f & KJ_DEV_E_LINE_TOO_LONG)
        return kj_fail(c->ctx, KJ_E_RANGE, "a sequence line is longer than 2^27 bytes (first-seen position field); use KJ_F_NO_ORDER");
    if (f & (KJ_DEV_E_XCHG_OVERFLOW | KJ_DEV_E_XCHG_INCOMPLETE))
        return kj_fail(c->ctx, KJ_E_RANGE, (f & KJ_DEV_E_XCHG_OVERFLOW)
                           ? "fixed-capacity exchange: a segment overflowed (use larger capacities or the two-phase exchange)"
                           : "fixed-capacity exchange: a sender's count was incomplete (use the two-phase exchange)");
    if (f & KJ_DEV_E_READS_OVERFLOW)
        return kj_fail(c->ctx, KJ_E_RANGE, "more than 2^36 reads (first-seen read field); use KJ_F_NO_ORDER");
    return kj_fail(c->ctx, KJ_E_CUDA, "unknown device error flag");
}

// expected number of emissions of `bytes` input bytes (i.i.d. bases): both strands, every
// position, times 4^-m
static double expected_emissions(const kj_counts *c, uint64_t bytes) {
    double e = (double)bytes;   // ~ half the bytes are bases, two strands
    const uint32_t m = (uint32_t)c->prefix.size();
    for (uint32_t i = 0; i < std::min<uint32_t>(m, 16); ++i) e *= 0.25;
    return e;
}

// kernel arguments of one piece (everything but the table pointers, which are filled in at launch time)
static KjScanArgs make_args(kj_counts *c, const uint8_t *dbuf, uint64_t n, uint64_t own_n, int final_) {
    const uint32_t m = (uint32_t)c->prefix.size();
    KjScanArgs a{};
    a.buf = dbuf; a.n = n; a.own_n = own_n; a.voff = c->voff;
    a.parity = c->parity; a.final_ = final_ ? 1u : 0u;
    a.k = c->k; a.step = c->step; a.m = m; a.order = c->order ? 1u : 0u;
    a.n_strands = (c->flags & KJ_F_FORWARD_ONLY) ? 1u : 2u;
    a.line_gate = (c->flags & KJ_F_NO_LINE_GATE) ? 0u : 1u;
    a.count_bases = (c->flags & KJ_F_COUNT_BASES) ? 1u : 0u;
    a.mp = std::min<uint32_t>(m, KJ_MAX_MP);
    a.rc_shift = (m <= c->k) ? c->k - m : 0;
    for (uint32_t i = 0; i < a.mp; ++i) {
        a.pat_f[i] = kj_code(c->prefix[i]) * 0x55555555u;
        a.pat_r[i] = kj_code(c->rprefix[i]) * 0x55555555u;
    }
    memset(a.prefix, 0, 32); memset(a.rprefix, 0, 32);
    if (m) {
        memcpy(a.prefix, c->prefix.data(), std::min<size_t>(32, m));
        memcpy(a.rprefix, c->rprefix.data(), std::min<size_t>(32, m));
    }
    {   // byte-exact window check of the filter path: wanted bytes and their mask, per strand
        uint8_t want[2][32], mask[2][32];
        memset(want, 0, sizeof(want)); memset(mask, 0, sizeof(mask));
        if (m <= c->k && c->k <= 32) {
            for (uint32_t i = 0; i < m; ++i) {
                want[0][i] = c->prefix[i]; mask[0][i] = 0xFF;
                want[1][c->k - m + i] = c->rprefix[i]; mask[1][c->k - m + i] = 0xFF;
            }
        }
        memcpy(a.want, want, sizeof(want));
        memcpy(a.wmask, mask, sizeof(mask));
    }
    a.ctr = c->ctr;
    a.c0a = 0x0A0A0A0Au; a.c7f = 0x7F7F7F7Fu;
    return a;
}

static int account_scan_time(kj_counts *c, uint64_t own_n, bool has_verify) {
    kj_ctx *ctx = c->ctx;
    float ms = 0.f;
    KJ_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    ctx->scan_ms += ms;
    if (has_verify) {
        float vms = 0.f;
        KJ_CUDA(ctx, cudaEventElapsedTime(&vms, ctx->ev2, ctx->ev1));
        ctx->verify_ms += vms;
    }
    ctx->scan_launches++;
    ctx->scan_bytes += own_n;
    return KJ_OK;
}

// keep the load factor at or below one half between launches
static int keep_load_factor(kj_counts *c) {
    int rc = KJ_OK;
    if (c->h_ctr->n_unique * 2 > c->cap) rc = grow_table(c, c->cap * 4);
    if (rc == KJ_OK && c->h_ctr->n_irr_unique * 2 > c->irr_cap) rc = grow_irr(c, c->irr_cap * 4);
    return rc;
}

// ------------------------------------------------------------------------------------ filter path

#ifndef KJ_CPU_EMU
typedef CUresult (*KjEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#endif

// the input as a 2-D tensor of rows of 128 bytes (whole rows only), boxes of 32 rows, 128-byte swizzle
static int make_tensor_map(kj_ctx *ctx, const uint8_t *dbuf, uint64_t n, KjTensorMap *tm) {
    const uint64_t rows = n / 128;
#ifdef KJ_CPU_EMU
    tm->base = dbuf; tm->rows = rows;
#else
    static KjEncodeTiled encode = nullptr;
    if (!encode) {
        cudaDriverEntryPointQueryResult q;
        void *fn = nullptr;
        KJ_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
        if (!fn) return kj_fail(ctx, KJ_E_CUDA, "the driver has no cuTensorMapEncodeTiled");
        encode = (KjEncodeTiled)fn;
    }
    memset(tm, 0, sizeof(*tm));
    if (rows == 0) return KJ_OK;                       // no whole row: every tile takes the bounds-checked path
    const cuuint64_t gdim[2] = {128, rows};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {128, 32};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<uint8_t *>(dbuf), gdim, gstr, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return kj_fail(ctx, KJ_E_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
#endif
    return KJ_OK;
}

typedef void (*KjFilterFn)(const KjTensorMap, const KjScanArgs);
static KjFilterFn pick_filter_kernel(const KjScanArgs &a) {
    switch (a.mp) {
        case 1: return kj_warp_filter_kernel<1>;
        case 2: return kj_warp_filter_kernel<2>;
        case 3: return kj_warp_filter_kernel<3>;
        case 4: return kj_warp_filter_kernel<4>;
        case 5: return kj_warp_filter_kernel<5>;
        case 6: return kj_warp_filter_kernel<6>;
        case 7: return kj_warp_filter_kernel<7>;
        default: return kj_warp_filter_kernel<8>;
    }
}
typedef void (*KjResolveFn)(const KjScanArgs);
// the byte check looks at 4 words of the window when k <= 16 (the KmerFinder default), at 8 otherwise
static KjResolveFn pick_resolve_kernel(const KjScanArgs &a) { return a.k <= 16 ? kj_resolve_kernel<4> : kj_resolve_kernel<8>; }

// scan -> exclusive scan of the tile counts -> resolve, all stream-ordered;
// retry_only: the resolve pass over the marked entries
static int launch_filter(kj_counts *c, KjPiece &pc, bool retry_only) {
    kj_ctx *ctx = c->ctx;
    KjScanArgs &a = pc.args;
    a.tab = c->tab; a.irr = c->irr; a.ovf = c->ovf;
    a.cand = c->cand; a.cand_cap = c->cand_cap;
    a.tile_cnt = c->tile_cnt; a.tile_excl = c->tile_mem;
    a.resolve_retry = retry_only ? 1u : 0u;
    // the failure counters of this pass (n_overflow, n_irr_overflow) and, for a full pass, the entry counter
    KJ_CUDA(ctx, cudaMemsetAsync(&c->ctr->n_overflow, 0, (retry_only ? 2 : 3) * sizeof(unsigned long long), ctx->stream));
    const bool timed = ctx->timers_on && !retry_only;
    if (timed) KJ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if (!retry_only) {
        KjFilterFn fn = pick_filter_kernel(a);
        // attribute and occupancy of a kernel are asked once per context (the host work in front of the launch is GPU idle time)
        int &occ = ctx->occ_cache[(const void *)fn];
        if (occ == 0) {
            KJ_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KJ_WT_SMEM_BYTES));
            KJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, KJ_WT_THREADS, KJ_WT_SMEM_BYTES));
            occ = std::max(occ, 1);
        }
        const uint64_t want_ctas = ((uint64_t)a.n_tiles + KJ_WT_WARPS - 1) / KJ_WT_WARPS;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want_ctas, (uint64_t)ctx->sm_count * std::max(occ, 1)));
        KJ_LAUNCH(fn, grid, KJ_WT_THREADS, KJ_WT_SMEM_BYTES, ctx->stream, pc.tmap, a);
        ctx->launches++;
        if (timed) KJ_CUDA(ctx, cudaEventRecord(ctx->ev2, ctx->stream));
#ifdef KJ_CPU_EMU
        {
            uint64_t run = 0;
            for (uint32_t t = 0; t < a.n_tiles; ++t) { a.tile_excl[t] = run; run += a.tile_cnt[t]; }
        }
#else
        {
            size_t tmp = c->scan_tmp_bytes;
            KJ_CUDA(ctx, cub::DeviceScan::ExclusiveSum(c->scan_tmp, tmp, a.tile_cnt, a.tile_excl, (int)a.n_tiles, ctx->stream));
            ctx->launches += 2;
        }
#endif
    }
    {
        // exactly the blocks that are resident together: every thread then runs its software pipeline over many rounds
        KjResolveFn rf = pick_resolve_kernel(a);
        int &occ = ctx->occ_cache[(const void *)rf];
        if (occ == 0) {
            KJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rf, 256, 0));
            occ = std::max(occ, 1);
        }
        KJ_LAUNCH(rf, ctx->sm_count * occ, 256, 0, ctx->stream, a);
    }
    ctx->launches++;
    if (a.count_bases && !retry_only) {
        KJ_LAUNCH(kj_bases_kernel, ctx->sm_count * 8, 256, 0, ctx->stream, a);
        ctx->launches++;
    }
    if (timed) KJ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KJ_CUDA(ctx, cudaGetLastError());
    pc.timed = timed;
    return KJ_OK;
}

// entry buffer and tile arrays for a piece of n_tiles tiles expecting `want_ent` entries
static int ensure_filter_buffers(kj_counts *c, uint32_t n_tiles, uint64_t want_ent) {
    kj_ctx *ctx = c->ctx;
    if (n_tiles > c->tile_cap) {
        kj_dfree(ctx, c->tile_mem); kj_dfree(ctx, c->tile_cnt); kj_dfree(ctx, c->scan_tmp);
        c->tile_mem = nullptr; c->tile_cnt = nullptr; c->scan_tmp = nullptr; c->tile_cap = 0;
        const uint64_t tc = std::max<uint64_t>(n_tiles, 4096);
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->tile_mem, tc * 8));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->tile_cnt, tc * 8));
        size_t tmp = 0;
#ifndef KJ_CPU_EMU
        KJ_CUDA(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, c->tile_cnt, c->tile_mem, (int)tc, ctx->stream));
#endif
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->scan_tmp, std::max<size_t>(tmp, 16)));
        c->scan_tmp_bytes = tmp;
        c->tile_cap = tc;
    }
    if (want_ent > c->cand_cap) {
        kj_dfree(ctx, c->cand);
        c->cand = nullptr; c->cand_cap = 0;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->cand, want_ent * 16));
        c->cand_cap = want_ent;
    }
    return KJ_OK;
}

// Wait for the piece in flight and deal with what the device reports: an entry buffer that was too small (the
// launch touched nothing: repeat it with a larger one), emissions that found no table slot (grow, run the
// retry pass over the marked entries), error flags.  Nothing is ever dropped.
static int settle_filter(kj_counts *c) {
    kj_ctx *ctx = c->ctx;
    if (!c->pending) return KJ_OK;
    KjPiece &pc = *c->piece;
    int rc = pull_counters(c);
    if (rc) return rc;
    if (pc.timed) { rc = account_scan_time(c, pc.args.own_n, true); if (rc) return rc; pc.timed = false; }
    for (int round = 0; round < 4 && c->h_ctr->n_cand > c->cand_cap; ++round) {
        rc = ensure_filter_buffers(c, pc.args.n_tiles, c->h_ctr->n_cand + (c->h_ctr->n_cand >> 3) + 1024);
        if (rc) return rc;
        rc = launch_filter(c, pc, false);
        if (rc) return rc;
        rc = pull_counters(c);
        if (rc) return rc;
    }
    if (c->h_ctr->n_cand > c->cand_cap) return kj_fail(ctx, KJ_E_CUDA, "internal: candidate entries exceed their own count");
    rc = check_device_errors(c);
    if (rc) return rc;
    for (int round = 0; c->h_ctr->n_overflow; ++round) {
        if (round >= 40) return kj_fail(ctx, KJ_E_TABLE_FULL, "k-mer table cannot take the emissions of this piece (device memory?)");
        // entries with lanes left over: the tables are too small or too crowded for them.  Their number bounds the
        // emissions still to come from below; every round at least doubles both tables.
        const uint64_t left = c->h_ctr->n_overflow;
        rc = grow_table(c, std::max<uint64_t>(c->cap * 4, 4 * (c->h_ctr->n_unique + left)));
        if (rc) return rc;
        rc = grow_irr(c, std::max<uint64_t>(c->irr_cap * 4, 4 * (c->h_ctr->n_irr_unique + left)));
        if (rc) return rc;
        rc = launch_filter(c, pc, true);
        if (rc) return rc;
        rc = pull_counters(c);
        if (rc) return rc;
        rc = check_device_errors(c);
        if (rc) return rc;
    }
    c->pending = false;
    return keep_load_factor(c);
}

// One piece through the filter path.  With a capacity hint on a device-resident buffer the host does not wait:
// the piece is settled by the next call that needs its results (kj_counts_finish, or another add).
static int scan_piece_filter(kj_counts *c, const uint8_t *dbuf, uint64_t n, uint64_t own_n, int final_, bool may_defer) {
    kj_ctx *ctx = c->ctx;
    int rc = settle_filter(c);                      // one piece in flight per handle
    if (rc) return rc;
    const uint32_t m = (uint32_t)c->prefix.size();
    // complement(prefix) is searched where it starts: rc_shift = k - m bytes behind the start of its window, so the tiles
    // reach that far past the owned range
    const uint64_t reach = (c->flags & KJ_F_FORWARD_ONLY) ? 0 : (uint64_t)(c->k - m);
    const uint64_t n_tiles64 = (own_n + reach + KJ_WT_BYTES - 1) / KJ_WT_BYTES;
    if (n_tiles64 > 0x7FFFFFF0ull / KJ_WT_OWN_ROWS) return kj_fail(ctx, KJ_E_RANGE, "piece too large");
    // table capacity: the caller's hint, else what this piece is expected to add (it grows by rehash between
    // pieces; emissions that find no slot wait for the retry pass)
    const uint64_t hard_bound = 2 * own_n + 16;
    uint64_t expect = std::min<uint64_t>((uint64_t)(4.0 * expected_emissions(c, own_n)) + 65536, hard_bound);
    const uint64_t known = c->h_ctr->n_unique;
    const uint64_t want = c->capacity_hint ? std::max<uint64_t>(2 * c->capacity_hint, 2 * known)
                                           : std::min<uint64_t>(2 * (known + expect), std::max<uint64_t>(1ull << 24, 4 * known));
    rc = grow_table(c, want);
    if (rc) return rc;
    rc = grow_irr(c, std::max<uint64_t>(std::max<uint64_t>(1ull << 12, 4 * c->h_ctr->n_irr_unique), expect / 32));
    if (rc) return rc;
    // entries: one per 16-byte chunk that holds a code-space match, 2 * 16 * 4^-mp of them for random bytes
    const uint64_t chunks = own_n / 16 + 1;
    double share = 48.0;
    for (uint32_t i = 0; i < std::min<uint32_t>(m, KJ_MAX_MP); ++i) share *= 0.25;
    // (with slack: 48 instead of 32); slots are reserved in blocks of KJ_WT_BLOCK per warp, so the buffer also holds a few
    // blocks for every warp
    const uint64_t want_ent = std::min<uint64_t>(hard_bound, (uint64_t)(std::min(1.0, share) * (double)chunks)) + (1ull << 20);
    rc = ensure_filter_buffers(c, (uint32_t)n_tiles64, want_ent);
    if (rc) return rc;

    KjPiece &pc = *c->piece;
    pc.args = make_args(c, dbuf, n, own_n, final_);
    pc.args.n_tiles = (uint32_t)n_tiles64;
    const uint64_t rows = n / 128;
    const uint64_t fast_rows = rows >= 32 ? (rows - 32) / KJ_WT_OWN_ROWS + 1 : 0;       // rows [31 t, 31 t + 32) exist
    pc.args.n_fast = (uint32_t)std::min<uint64_t>(std::min<uint64_t>(fast_rows, own_n / KJ_WT_BYTES), n_tiles64);
    rc = make_tensor_map(ctx, dbuf, n, &pc.tmap);
    if (rc) return rc;
    rc = launch_filter(c, pc, false);
    if (rc) return rc;
    c->pending = true;
    c->voff += own_n;
    c->consumed += own_n;
    c->parity ^= 1u;
    if (may_defer && c->capacity_hint) return KJ_OK;
    return settle_filter(c);
}

// ------------------------------------------------------------------------------------ dense and line kernels

// One kernel launch over a device-resident piece.  Synchronises, handles spills and growth.
static int scan_piece(kj_counts *c, const uint8_t *dbuf, uint64_t n, uint64_t own_n, int final_, bool may_defer) {
    kj_ctx *ctx = c->ctx;
    if (own_n == 0) return KJ_OK;
    if (c->use_filter) return scan_piece_filter(c, dbuf, n, own_n, final_, may_defer);
    const uint64_t n_tiles64 = (own_n + KJ_TILE_BYTES - 1) / KJ_TILE_BYTES;
    if (n_tiles64 > 0xFFFFFFF0ull) return kj_fail(ctx, KJ_E_RANGE, "piece too large");
    const uint32_t n_tiles = (uint32_t)n_tiles64;

    // every window can be an emission: the spill lists take the worst case of a piece, so nothing is ever dropped
    const uint64_t hard_bound = 2 * own_n + 16;
    uint64_t known = c->h_ctr->n_unique;
    uint64_t want = c->capacity_hint ? std::max<uint64_t>(2 * c->capacity_hint, 2 * known)
                                     : std::min<uint64_t>(2 * (known + hard_bound), std::max<uint64_t>(1ull << 24, 4 * known));
    int rc = grow_table(c, want);
    if (rc) return rc;
    rc = grow_irr(c, std::max<uint64_t>(std::max<uint64_t>(1ull << 12, 4 * c->h_ctr->n_irr_unique), std::min<uint64_t>(hard_bound, 1ull << 22)));
    if (rc) return rc;
    rc = ensure_overflow(c, hard_bound, hard_bound);
    if (rc) return rc;

    // tile state
    if (n_tiles > c->tile_cap) {
        kj_dfree(ctx, c->tile_mem);
        c->tile_mem = nullptr; c->tile_cap = 0;
        uint64_t tc = std::max<uint64_t>(n_tiles, 4096);
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->tile_mem, tc * 8));
        c->tile_cap = tc;
    }
    KJ_CUDA(ctx, cudaMemsetAsync(c->tile_mem, 0, (uint64_t)n_tiles * 8, ctx->stream));  // status only
    KJ_CUDA(ctx, cudaMemsetAsync(&c->ctr->ticket, 0, sizeof(unsigned int), ctx->stream));

    KjScanArgs a = make_args(c, dbuf, n, own_n, final_);
    a.n_tiles = n_tiles;
    a.tab = c->tab; a.irr = c->irr; a.ovf = c->ovf;
    a.status = c->tile_mem;

    void (*fn)(const KjScanArgs) = c->use_dense ? kj_scan_dense_kernel : kj_scan_lines_kernel;
    int occ = 0;
    KJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, KJ_THREADS, 0));
    // persistent CTAs: a whole number of CTAs per SM, all resident (tiles are handed out by ticket and
    // a tile waits for the aggregates of the tiles before it)
    const int grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * std::max(occ, 1));
    if (ctx->timers_on) KJ_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    KJ_LAUNCH(fn, grid, KJ_THREADS, 0, ctx->stream, a);
    ctx->launches++;
    if (ctx->timers_on) KJ_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KJ_CUDA(ctx, cudaGetLastError());
    rc = pull_counters(c);
    if (rc) return rc;
    if (ctx->timers_on) { rc = account_scan_time(c, own_n, false); if (rc) return rc; }
    rc = check_device_errors(c);
    if (rc) return rc;

    // spills: grow and replay until clean
    for (int round = 0; round < 8 && (c->h_ctr->n_overflow || c->h_ctr->n_irr_overflow); ++round) {
        uint64_t nov = c->h_ctr->n_overflow, niov = c->h_ctr->n_irr_overflow;
        if (nov > c->ovf.cap || niov > c->ovf.irr_cap)
            return kj_fail(ctx, KJ_E_TABLE_FULL, "internal: more spills than a piece can emit");
        if (nov) { rc = grow_table(c, std::max<uint64_t>(c->cap * 4, 4 * (c->h_ctr->n_unique + nov))); if (rc) return rc; }
        if (niov) { rc = grow_irr(c, std::max<uint64_t>(c->irr_cap * 4, 4 * (c->h_ctr->n_irr_unique + niov))); if (rc) return rc; }
        KJ_CUDA(ctx, cudaMemsetAsync(&c->ctr->n_overflow, 0, 2 * sizeof(unsigned long long), ctx->stream));
        // after growing 4x a spill during the replay means a broken hash: the replay kernels raise
        // the error flag instead of appending to the list they are reading
        if (nov) {
            KJ_LAUNCH(kj_replay_kernel, grid_for(ctx, nov), 256, 0, ctx->stream, c->tab, c->ctr, c->ovf.rec, nov);
            ctx->launches++;
        }
        if (niov) {
            KJ_LAUNCH(kj_replay_irr_kernel, grid_for(ctx, niov), 256, 0, ctx->stream, c->irr, c->ctr, dbuf,
                      c->ovf.irr_rec, niov);
            ctx->launches++;
        }
        rc = pull_counters(c);
        if (rc) return rc;
        rc = check_device_errors(c);
        if (rc) return rc;
    }
    rc = keep_load_factor(c);
    if (rc) return rc;

    c->voff += own_n;
    c->consumed += own_n;
    c->parity ^= 1u;
    return KJ_OK;
}

// device-resident buffer.  The filter kernel takes it in one launch when the caller sized the
// table (capacity_hint), else in 512 MiB pieces so the table can grow in between; the line
// kernel (whose spill lists are sized for the worst case) in 4 MiB pieces.
static int scan_device(kj_counts *c, const uint8_t *dbuf, uint64_t n, uint64_t own_n, int final_, bool may_defer = false) {
    // pieces are multiples of the filter path's tile, so that every piece but the last ends on a tile boundary
    const uint64_t piece = c->use_filter ? (c->capacity_hint ? (1ull << 40) : (uint64_t)KJ_WT_BYTES * 16 * 8192)
                                         : (c->use_dense ? (8ull << 20) : (4ull << 20));
    uint64_t lo = 0;
    while (lo < own_n) {
        uint64_t len = std::min(piece, own_n - lo);
        int rc = scan_piece(c, dbuf + lo, n - lo, len, (final_ && lo + len == own_n) ? 1 : 0, may_defer && len == own_n);
        if (rc) return rc;
        lo += len;
    }
    return KJ_OK;
}

// ------------------------------------------------------------------------------------ API

extern "C" int kj_counts_create(kj_ctx *ctx, const kj_count_params *p, kj_counts **out) {
    if (!ctx || !p || !out) return kj_fail(ctx, KJ_E_INVALID, "kj_counts_create: null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (p->k < 1 || p->k > 32) return kj_fail(ctx, KJ_E_RANGE, "k must be in 1..32");
    if (p->step < 1) return kj_fail(ctx, KJ_E_INVALID, "step must be >= 1");
    if (p->prefix_len && !p->prefix) return kj_fail(ctx, KJ_E_INVALID, "prefix is null");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    kj_counts *c = new kj_counts();
    c->ctx = ctx;
    c->piece = new KjPiece();
    c->prefix.assign(p->prefix, p->prefix + p->prefix_len);
    c->rprefix.resize(p->prefix_len);
    for (uint32_t i = 0; i < p->prefix_len; ++i)
        c->rprefix[i] = kj_comp_byte(p->prefix[p->prefix_len - 1 - i]);
    c->k = p->k; c->step = p->step; c->flags = p->flags;
    c->order = !(p->flags & KJ_F_NO_ORDER);
    c->use_filter = p->step == 1 && p->prefix_len >= 1 && p->prefix_len <= p->k &&
                    !(p->flags & KJ_F_FORCE_GENERIC);
    c->use_dense = p->step == 1 && p->prefix_len == 0 && p->k >= 2 && !(p->flags & KJ_F_FORCE_GENERIC);
    c->capacity_hint = p->capacity_hint;
    c->voff = p->base_col;
    c->base_line = p->base_line;
    c->base_col = p->base_col;
    cudaError_t e = kj_dmalloc(ctx, &c->ctr, sizeof(KjCounters));
    static_assert(sizeof(KjCounters) <= 1024, "pinned block size");
    if (e == cudaSuccess) c->h_ctr = (KjCounters *)kj_pinned_get(ctx);
    if (e != cudaSuccess || !c->h_ctr) {
        if (e == cudaSuccess) e = cudaErrorMemoryAllocation;
        kj_counts_free(c);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_create: ") + cudaGetErrorString(e));
    }
    memset(c->h_ctr, 0, sizeof(KjCounters));
    c->h_ctr->special_ord = ~0ull;
    c->h_ctr->carry_lines[0] = p->base_line;
    c->h_ctr->carry_last[0] = 0;
    e = cudaMemcpyAsync(c->ctr, c->h_ctr, sizeof(KjCounters), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        kj_counts_free(c);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_create: ") + cudaGetErrorString(e));
    }
    *out = c;
    return KJ_OK;
}

// host -> device staging chunk of the context: kj_set_stage_chunk, else KJ_STAGE_CHUNK_MB, else 64 MiB
static uint64_t kj_stage_chunk(kj_ctx *ctx) {
    if (!ctx->stage_chunk) {
        const char *e = getenv("KJ_STAGE_CHUNK_MB");
        long mb = e ? atol(e) : 0;
        ctx->stage_chunk = (mb >= 1 && mb <= 4096) ? ((uint64_t)mb << 20) : (64ull << 20);
    }
    return ctx->stage_chunk;
}
#define KJ_STAGE_CHUNK kj_stage_chunk(ctx)
static const uint64_t KJ_STAGE_HALO_LINES = 1ull << 20;   // line kernel: longest line it can finish

extern "C" int kj_set_stage_chunk(kj_ctx *ctx, uint64_t bytes) {
    if (!ctx) return KJ_E_INVALID;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (bytes < 4096 || bytes > (4096ull << 20)) return kj_fail(ctx, KJ_E_INVALID, "staging chunk must be 4 KiB .. 4 GiB");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
    for (int i = 0; i < 2; ++i) {              // reallocated at the next use
        cudaFree(ctx->d_stage[i]); ctx->d_stage[i] = nullptr;
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        ctx->h_stage[i] = nullptr;
    }
    ctx->stage_chunk = (bytes + 15) / 16 * 16;
    ctx->stage_cap = 0;
    return KJ_OK;
}

static int ensure_staging(kj_ctx *ctx, bool need_host) {
    const uint64_t cap = KJ_STAGE_CHUNK + KJ_STAGE_HALO_LINES;
    if (!ctx->d_stage[0]) {
        for (int i = 0; i < 2; ++i) {
            KJ_CUDA(ctx, cudaMalloc(&ctx->d_stage[i], cap + 64));
            if (!ctx->ev_copy[i]) KJ_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
        }
        ctx->stage_cap = cap;
    }
    if (need_host && !ctx->h_stage[0])
        for (int i = 0; i < 2; ++i) KJ_CUDA(ctx, cudaMallocHost(&ctx->h_stage[i], cap));
    return KJ_OK;
}

static int add_host(kj_counts *c, const uint8_t *buf, uint64_t n, uint64_t own_n, int final_) {
    kj_ctx *ctx = c->ctx;
    const uint64_t halo = (c->use_filter || c->use_dense) ? 64 : KJ_STAGE_HALO_LINES;
    const uint64_t chunk = KJ_STAGE_CHUNK;
    cudaPointerAttributes attr{};
    bool pinned = cudaPointerGetAttributes(&attr, buf) == cudaSuccess &&
                  (attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged);
    cudaGetLastError();
    int rc = ensure_staging(ctx, !pinned);
    if (rc) return rc;

    const uint64_t n_chunks = (own_n + chunk - 1) / chunk;
    auto issue_copy = [&](uint64_t ci) -> int {
        uint64_t lo = ci * chunk;
        uint64_t hi = std::min(own_n, lo + chunk);
        uint64_t rd = std::min(n, hi + halo) - lo;          // bytes readable by this piece
        int s = (int)(ci & 1);
        const uint8_t *src = buf + lo;
        if (!pinned) { memcpy(ctx->h_stage[s], src, rd); src = ctx->h_stage[s]; }
        KJ_CUDA(ctx, cudaMemcpyAsync(ctx->d_stage[s], src, rd, cudaMemcpyHostToDevice, ctx->copy_stream));
        KJ_CUDA(ctx, cudaEventRecord(ctx->ev_copy[s], ctx->copy_stream));
        return KJ_OK;
    };
    // the launch stream is idle here (every scan_piece ends synchronised), so both slots are free
    rc = n_chunks ? issue_copy(0) : KJ_OK;
    if (rc) return rc;
    for (uint64_t ci = 0; ci < n_chunks; ++ci) {
        uint64_t lo = ci * chunk;
        uint64_t hi = std::min(own_n, lo + chunk);
        uint64_t rd = std::min(n, hi + halo) - lo;
        int s = (int)(ci & 1);
        KJ_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[s], 0));
        // the other slot's kernel was synchronised by the previous iteration: start the next copy
        if (ci + 1 < n_chunks) { rc = issue_copy(ci + 1); if (rc) return rc; }
        rc = scan_device(c, ctx->d_stage[s], rd, hi - lo, (final_ && lo + rd == n) ? 1 : 0);
        if (rc) return rc;
    }
    return KJ_OK;
}

extern "C" int kj_counts_add_buffer(kj_counts *c, const uint8_t *buf, uint64_t n, uint64_t own_n,
                                    int mem_kind, int final_) {
    if (!c) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (c->finished || c->saw_final) return kj_fail(ctx, KJ_E_STATE, "kj_counts_add_buffer after the final buffer / finish");
    if (own_n > n) return kj_fail(ctx, KJ_E_INVALID, "own_n > n");
    if (n && !buf) return kj_fail(ctx, KJ_E_INVALID, "buf is null");
    if (final_ && own_n != n) return kj_fail(ctx, KJ_E_INVALID, "final buffer must own all its bytes");
    if (!final_ && n - own_n < 32) return kj_fail(ctx, KJ_E_INVALID, "non-final buffer needs a halo of at least 32 bytes");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    if (mem_kind == KJ_MEM_DEVICE) {
        if ((uintptr_t)buf & 15) return kj_fail(ctx, KJ_E_INVALID, "device buffers must be 16-byte aligned");
        rc = scan_device(c, buf, n, own_n, final_, true);
    } else if (mem_kind == KJ_MEM_HOST) {
        rc = add_host(c, buf, n, own_n, final_);
    } else {
        return kj_fail(ctx, KJ_E_INVALID, "mem_kind");
    }
    if (rc) return rc;
    if (final_) c->saw_final = true;
    return KJ_OK;
}

// pread of [off, off + len) by three threads per four cores, twelve at most (page-cache / tmpfs copies scale with threads
// up to there on a 16-core host: 15.4 / 20.3 / 23.4 / 21.8 GB/s with 4 / 8 / 12 / 16, profiles/r02_file_leg_threads.txt: the
// read is what bounds the file leg); environment KJ_READ_THREADS overrides
static bool read_range(int fd, uint8_t *dst, uint64_t off, uint64_t len) {
    const unsigned hw = std::max(2u, std::thread::hardware_concurrency());
    const char *env = getenv("KJ_READ_THREADS");
    const unsigned want = env && *env ? (unsigned)std::max(1, atoi(env)) : std::max(1u, std::min(12u, hw * 3u / 4u));
    const unsigned nt = len >= (8u << 20) ? std::min(want, 64u) : 1u;
    std::atomic<bool> ok(true);
    auto part = [&](unsigned i) {
        uint64_t lo = len * i / nt, hi = len * (i + 1) / nt;
        while (lo < hi) {
            ssize_t r = pread(fd, dst + lo, hi - lo, (off_t)(off + lo));
            if (r <= 0) { ok = false; return; }
            lo += (uint64_t)r;
        }
    };
    std::vector<std::thread> th;
    for (unsigned i = 1; i < nt; ++i) th.emplace_back(part, i);
    part(0);
    for (auto &t : th) t.join();
    return ok;
}

// gzip'd FASTQ (an ingest format the reference does not have: SURVEY.md 8f rank 4).  zlib inflates on the reader thread
// straight into the pinned staging buffers; pieces, halos and the copy / count overlap are those of the plain file, except
// that the length of the stream is only known at its end: a piece is final when the stream ended inside its halo.
struct KjGzFill { uint64_t n; bool eof; bool ok; };
static int add_file_gz(kj_counts *c, int fd, const char *path) {
    kj_ctx *ctx = c->ctx;
    const uint64_t halo = (c->use_filter || c->use_dense) ? 64 : KJ_STAGE_HALO_LINES;
    const uint64_t chunk = KJ_STAGE_CHUNK;
    int rc = KJ_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) rc = kj_fail(ctx, KJ_E_CUDA, "cudaSetDevice");
    if (rc == KJ_OK) rc = ensure_staging(ctx, true);
    gzFile gz = rc == KJ_OK ? gzdopen(fd, "rb") : nullptr;          // owns fd from here on
    if (!gz) { close(fd); return rc ? rc : kj_fail(ctx, KJ_E_IO, std::string("cannot read ") + path); }
    gzbuffer(gz, 1u << 20);
    auto fill = [&, gz](int s, const uint8_t *carry, uint64_t carry_n) -> KjGzFill {
        uint8_t *dst = ctx->h_stage[s];
        if (carry_n) memcpy(dst, carry, carry_n);
        uint64_t n = carry_n;
        const uint64_t want = chunk + halo;
        while (n < want) {
            const int r = gzread(gz, dst + n, (unsigned)std::min<uint64_t>(want - n, 1u << 30));
            if (r < 0) return {n, false, false};
            if (r == 0) {
                // end of the stream -- or of a truncated file: zlib reports that one only through gzerror
                int zerr = Z_OK;
                gzerror(gz, &zerr);
                return {n, true, zerr == Z_OK || zerr == Z_STREAM_END};
            }
            n += (uint64_t)r;
        }
        return {n, false, true};
    };
    uint64_t total = 0;
    std::future<KjGzFill> fut = std::async(std::launch::async, fill, 0, (const uint8_t *)nullptr, (uint64_t)0);
    for (size_t i = 0; rc == KJ_OK; ++i) {
        const int s = (int)(i & 1);
        const KjGzFill f = fut.get();
        if (!f.ok) { rc = kj_fail(ctx, KJ_E_IO, std::string("gzip stream error in ") + path); break; }
        // the stream ended inside this piece's halo (or before): it owns everything it holds and is the last one
        const bool final_ = f.eof && (f.n <= chunk || f.n - chunk < 32);
        const uint64_t own = final_ ? f.n : chunk;
        if (!final_) {
            // the other pinned buffer was the source of the copy of piece i - 1
            if (i >= 1 && cudaEventSynchronize(ctx->ev_copy[s ^ 1]) != cudaSuccess) { rc = kj_fail(ctx, KJ_E_CUDA, "cudaEventSynchronize"); break; }
            fut = std::async(std::launch::async, fill, s ^ 1, (const uint8_t *)(ctx->h_stage[s] + own), f.n - own);
        }
        total += own;
        if (f.n) {
            cudaError_t e = cudaMemcpyAsync(ctx->d_stage[s], ctx->h_stage[s], f.n, cudaMemcpyHostToDevice, ctx->copy_stream);
            if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy[s], ctx->copy_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[s], 0);
            if (e != cudaSuccess) { rc = kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_add_file: ") + cudaGetErrorString(e)); break; }
            rc = scan_device(c, ctx->d_stage[s], f.n, own, final_ ? 1 : 0);
        }
        if (final_) break;
    }
    if (fut.valid()) fut.wait();
    gzclose(gz);
    if (rc == KJ_OK) { c->bytes_read = total; c->saw_final = true; }
    return rc;
}

// KmerJS#readFile (lib/kmers.js:106-185): the file in staging chunks through two pinned buffers of the context; a
// reader thread fills one while the other is copied to the device and counted.
extern "C" int kj_counts_add_file(kj_counts *c, const char *path) {
    if (!c || !path) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (c->finished || c->saw_final) return kj_fail(ctx, KJ_E_STATE, "kj_counts_add_file after the final buffer / finish");
    int fd = open(path, O_RDONLY);
    if (fd < 0) return kj_fail(ctx, KJ_E_IO, std::string("cannot open ") + path);
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return kj_fail(ctx, KJ_E_IO, std::string("cannot stat ") + path); }
    {
        uint8_t magic[2] = {0, 0};
        if (pread(fd, magic, 2, 0) == 2 && magic[0] == 0x1F && magic[1] == 0x8B) return add_file_gz(c, fd, path);
    }
    const uint64_t size = (uint64_t)st.st_size;
    const uint64_t halo = (c->use_filter || c->use_dense) ? 64 : KJ_STAGE_HALO_LINES;
    const uint64_t chunk = KJ_STAGE_CHUNK;
    int rc = KJ_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) rc = kj_fail(ctx, KJ_E_CUDA, "cudaSetDevice");
    if (rc == KJ_OK) rc = ensure_staging(ctx, true);
    // piece boundaries: a tail shorter than the 32-byte halo a non-final piece must bring belongs to the piece before it
    std::vector<uint64_t> cuts;
    for (uint64_t lo = 0; lo < size;) {
        uint64_t hi = std::min(size, lo + chunk);
        if (size - hi < 32) hi = size;
        cuts.push_back(hi);
        lo = hi;
    }
    const size_t np = cuts.size();
    auto lo_of = [&](size_t i) { return i ? cuts[i - 1] : 0; };
    auto rd_of = [&](size_t i) { return std::min(size, cuts[i] + halo) - lo_of(i); };      // bytes the piece reads (with halo)
    std::future<bool> fut;
    if (rc == KJ_OK && np) fut = std::async(std::launch::async, [&, fd]() { return read_range(fd, ctx->h_stage[0], 0, rd_of(0)); });
    for (size_t i = 0; i < np && rc == KJ_OK; ++i) {
        const int s = (int)(i & 1);
        if (!fut.get()) { rc = kj_fail(ctx, KJ_E_IO, std::string("read error on ") + path); break; }
        if (i + 1 < np) {
            // the other pinned buffer was the source of the copy of piece i - 1
            if (i >= 1 && cudaEventSynchronize(ctx->ev_copy[s ^ 1]) != cudaSuccess) { rc = kj_fail(ctx, KJ_E_CUDA, "cudaEventSynchronize"); break; }
            const size_t nx = i + 1;
            fut = std::async(std::launch::async, [&, fd, nx, s]() { return read_range(fd, ctx->h_stage[s ^ 1], lo_of(nx), rd_of(nx)); });
        }
        const uint64_t lo = lo_of(i), hi = cuts[i], rd = rd_of(i);
        // device slot s was read by the scan of piece i - 2, which has been settled (host-staged pieces are not deferred)
        cudaError_t e = cudaMemcpyAsync(ctx->d_stage[s], ctx->h_stage[s], rd, cudaMemcpyHostToDevice, ctx->copy_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy[s], ctx->copy_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[s], 0);
        if (e != cudaSuccess) { rc = kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_add_file: ") + cudaGetErrorString(e)); break; }
        rc = scan_device(c, ctx->d_stage[s], rd, hi - lo, hi == size ? 1 : 0);
    }
    if (fut.valid()) fut.wait();
    close(fd);
    if (rc == KJ_OK) { c->bytes_read = size; c->saw_final = true; }
    return rc;
}

static void drop_compact(kj_counts *c) {
    kj_ctx *ctx = c->ctx;
    kj_dfree(ctx, c->reg.keys); kj_dfree(ctx, c->reg.counts); kj_dfree(ctx, c->reg.ords); kj_dfree(ctx, c->reg.alive);
    c->reg = KjCompact{};
    c->irr_host.clear();
    c->export_perm.clear();
    c->export_rank.clear();
}

// Large irregular sets (dense emission over reads with N: BASELINE config 5 has some 10^8 of them) are put into
// first-seen order on the device: sort keys out, cub radix sort of (ordinal, index), one gather that also writes the
// count / ordinal columns of the compact arrays.  The host gets the records already sorted.
__global__ void kj_irr_sortkeys_kernel(const KjIrrRecord *rec, uint64_t n, uint64_t *ord, uint32_t *idx) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        ord[i] = rec[i].ord;
        idx[i] = (uint32_t)i;
    }
}
__global__ void kj_irr_gather_kernel(const KjIrrRecord *rec, const uint32_t *perm, uint64_t n, KjIrrRecord *sorted,
                                     uint64_t *counts, uint64_t *ords) {
    // a record is 56 bytes = 7 words of 8: seven threads per record keep the copies coalesced
    const uint64_t total = n * 7;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = w / 7, part = w % 7;
        const uint64_t src = perm ? perm[i] : i;
        const uint64_t v = reinterpret_cast<const uint64_t *>(rec + src)[part];
        reinterpret_cast<uint64_t *>(sorted + i)[part] = v;
        if (part == 5) counts[i] = v;          // {key[32], len, count, ord}
        if (part == 6) ords[i] = v;
    }
}
// below this many records the host sorts them while it waits anyway (test hook: environment KJ_IRR_DEVICE_SORT_MIN)
static uint64_t irr_device_sort_min() {
    const char *e = getenv("KJ_IRR_DEVICE_SORT_MIN");
    return e && *e ? strtoull(e, nullptr, 10) : 65536ull;
}

// The single-wait finish does the same without knowing any count on the host: d_irr has `cap` slots (unused ones all
// ones, so that they sort last), the numbers come from the device counters, and nothing waits.
__global__ void kj_irr_gather_spec_kernel(const KjIrrRecord *rec, const uint32_t *perm, uint64_t cap, const KjCounters *ctr,
                                          uint64_t cap_tab, KjIrrRecord *sorted, uint64_t *keys, uint64_t *counts, uint64_t *ords) {
    unsigned long long n_tab = 0;
    for (int i = 0; i < 64; ++i) n_tab += ctr->n_unique_part[i];
    if (n_tab > cap_tab) return;                           // the compaction did not fit: the host takes the long way
    const uint64_t n = ctr->n_irr_compact < cap ? ctr->n_irr_compact : cap;
    const uint64_t base = n_tab + (ctr->special_count ? 1ull : 0ull);
    const uint64_t total = n * 7;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total; w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = w / 7, part = w % 7;
        const uint64_t v = reinterpret_cast<const uint64_t *>(rec + (perm ? perm[i] : i))[part];
        reinterpret_cast<uint64_t *>(sorted + i)[part] = v;
        if (part == 0) keys[base + i] = ~0ull;
        if (part == 5) counts[base + i] = v;
        if (part == 6) ords[base + i] = v;
    }
}
struct KjIrrSpec { KjIrrRecord *sorted = nullptr; void *tmp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; };
static void irregular_spec_free(kj_ctx *ctx, KjIrrSpec &q) {
    kj_dfree(ctx, q.sorted);
    for (void *p : q.tmp) kj_dfree(ctx, p);
    q = KjIrrSpec{};
}
// queue: (sort: sort keys out of the `cap` slots of d_irr, radix sort,) gather into q.sorted + the columns behind the table
// entries.  Without the sort the records keep the order the compaction gave them: the order of the query entries inside a
// handle is not observable (exports are sorted by first-seen ordinal on the device, matches go by key), and a radix sort of
// a few thousand records is eight passes of launch latency (0.1 ms) between the kernels of a step.
static int irregular_spec_queue(kj_counts *c, const KjIrrRecord *d_irr, uint64_t cap, uint64_t cap_tab, KjIrrSpec &q, bool sort) {
    kj_ctx *ctx = c->ctx;
    uint64_t *d_ord = nullptr, *d_ord2 = nullptr;
    uint32_t *d_idx = nullptr, *d_perm = nullptr;
    void *d_tmp = nullptr;
    cudaError_t e = kj_dmalloc(ctx, &q.sorted, cap * sizeof(KjIrrRecord));
    if (!sort) {
        if (e == cudaSuccess) {
            KJ_LAUNCH(kj_irr_gather_spec_kernel, grid_for(ctx, cap * 7), 256, 0, ctx->stream, d_irr, (const uint32_t *)nullptr, cap, c->ctr,
                      cap_tab, q.sorted, c->reg.keys, c->reg.counts, c->reg.ords);
            ctx->launches++;
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { irregular_spec_free(ctx, q); return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_finish (irregular k-mers): ") + cudaGetErrorString(e)); }
        return KJ_OK;
    }
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_ord, cap * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_ord2, cap * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_idx, cap * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_perm, cap * 4);
    q.tmp[0] = d_ord; q.tmp[1] = d_ord2; q.tmp[2] = d_idx; q.tmp[3] = d_perm;
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_irr_sortkeys_kernel, grid_for(ctx, cap), 256, 0, ctx->stream, d_irr, cap, d_ord, d_idx);
        ctx->launches++;
#ifndef KJ_CPU_EMU
        size_t tmp_bytes = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_ord, d_ord2, d_idx, d_perm, (int)cap, 0, 64, ctx->stream);
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_tmp, std::max<size_t>(tmp_bytes, 16));
        q.tmp[4] = d_tmp;
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_ord, d_ord2, d_idx, d_perm, (int)cap, 0, 64, ctx->stream);
        ctx->launches += 4;
#else
        {   // tools/cuemu: "device" memory is host memory
            std::vector<uint32_t> p(cap);
            for (uint64_t i = 0; i < cap; ++i) p[i] = (uint32_t)i;
            std::stable_sort(p.begin(), p.end(), [&](uint32_t x, uint32_t y) { return d_irr[x].ord < d_irr[y].ord; });
            memcpy(d_perm, p.data(), cap * 4);
        }
#endif
    }
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_irr_gather_spec_kernel, grid_for(ctx, cap * 7), 256, 0, ctx->stream, d_irr, d_perm, cap, c->ctr, cap_tab,
                  q.sorted, c->reg.keys, c->reg.counts, c->reg.ords);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { irregular_spec_free(ctx, q); return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_finish (irregular k-mers): ") + cudaGetErrorString(e)); }
    return KJ_OK;
}

// d_irr: the n_irr compacted records in slot order.  Leaves them sorted in c->irr_host and writes the key (unused: all ones),
// count and ordinal columns of the compact arrays from entry n_reg on.
static int irregular_sorted_on_device(kj_counts *c, const KjIrrRecord *d_irr, uint64_t n_irr, uint64_t n_reg) {
    kj_ctx *ctx = c->ctx;
    c->irr_host.resize(n_irr * sizeof(KjIrrRecord));
    if (!n_irr) return KJ_OK;
    KjIrrRecord *d_sorted = nullptr;
    uint64_t *d_ord = nullptr, *d_ord2 = nullptr;
    uint32_t *d_idx = nullptr, *d_perm = nullptr;
    void *d_tmp = nullptr;
    const bool sort = c->order && n_irr > 1;
    cudaError_t e = kj_dmalloc(ctx, &d_sorted, n_irr * sizeof(KjIrrRecord));
    if (sort) {
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_ord, n_irr * 8);
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_ord2, n_irr * 8);
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_idx, n_irr * 4);
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_perm, n_irr * 4);
        if (e == cudaSuccess) {
            KJ_LAUNCH(kj_irr_sortkeys_kernel, grid_for(ctx, n_irr), 256, 0, ctx->stream, d_irr, n_irr, d_ord, d_idx);
            ctx->launches++;
#ifndef KJ_CPU_EMU
            size_t tmp_bytes = 0;
            e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_ord, d_ord2, d_idx, d_perm, (int)n_irr, 0, 64, ctx->stream);
            if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_tmp, std::max<size_t>(tmp_bytes, 16));
            if (e == cudaSuccess)
                e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_ord, d_ord2, d_idx, d_perm, (int)n_irr, 0, 64, ctx->stream);
            ctx->launches += 4;
#else
            {   // tools/cuemu: "device" memory is host memory
                std::vector<uint32_t> p(n_irr);
                for (uint64_t i = 0; i < n_irr; ++i) p[i] = (uint32_t)i;
                std::stable_sort(p.begin(), p.end(), [&](uint32_t x, uint32_t y) { return d_irr[x].ord < d_irr[y].ord; });
                memcpy(d_perm, p.data(), n_irr * 4);
            }
#endif
        }
    }
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_irr_gather_kernel, grid_for(ctx, n_irr * 7), 256, 0, ctx->stream, d_irr, sort ? d_perm : nullptr, n_irr,
                  d_sorted, c->reg.counts + n_reg, c->reg.ords + n_reg);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(c->reg.keys + n_reg, 0xFF, n_irr * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->irr_host.data(), d_sorted, n_irr * sizeof(KjIrrRecord), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    kj_dfree(ctx, d_sorted); kj_dfree(ctx, d_ord); kj_dfree(ctx, d_ord2); kj_dfree(ctx, d_idx); kj_dfree(ctx, d_perm); kj_dfree(ctx, d_tmp);
    if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_finish (irregular k-mers): ") + cudaGetErrorString(e));
    return KJ_OK;
}

extern "C" int kj_counts_finish(kj_counts *c) {
    if (!c) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    // Everything the device has to do is queued first -- compaction of the table, compaction and copy back of the irregular
    // records, the counters -- and the host waits once.  With a capacity hint that holds even for a piece that is still in
    // flight: the compaction is sized by the hint and queued behind the count kernels; should the device report a table
    // or buffer that was too small, the piece is settled the long way and the compaction repeated with the exact sizes.
    bool irr_on_device = false;      // the irregular records came back sorted, their columns are written
    // (the owner side of a fixed-capacity exchange -- a handle that kj_counts_merge_segments filled -- is the same case: its
    // capacity hint bounds what the segments can hold)
    bool spec = (c->pending || c->exchange_totals) && c->capacity_hint != 0;
    int rc = KJ_OK;
    for (;;) {
        uint64_t cap_tab, cap_irr, irr_copied = 0;
        if (spec) {
            cap_tab = c->capacity_hint;
            // room for 65536 irregular k-mers in the compaction; the first 8192 come back with the counters, a larger set
            // takes one more copy (10 M reads of the bench workload leave some 10^4 of them: reads with N)
            cap_irr = std::min<uint64_t>(c->irr_cap, 65536);
        } else {
            rc = settle_filter(c);
            if (rc) return rc;
            rc = pull_counters(c);
            if (rc) return rc;
            rc = check_device_errors(c);
            if (rc) return rc;
            cap_tab = c->h_ctr->n_unique;
            cap_irr = c->h_ctr->n_irr_unique;
        }
        drop_compact(c);   // finish may be called again after merges
        const uint64_t q_cap = cap_tab + 1 + cap_irr;
        if (q_cap > 0xFFFFFFF0ull) return kj_fail(ctx, KJ_E_RANGE, "more than 2^32 distinct k-mers on one GPU");
        KJ_CUDA(ctx, cudaMemsetAsync(&c->ctr->n_compact, 0, 2 * sizeof(unsigned long long), ctx->stream));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->reg.keys, q_cap * 8));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->reg.counts, q_cap * 8));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->reg.ords, q_cap * 8));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &c->reg.alive, q_cap));
        KJ_CUDA(ctx, cudaMemsetAsync(c->reg.alive, 1, q_cap, ctx->stream));
        if (cap_tab && c->cap) {
            KJ_LAUNCH(kj_compact_kernel, grid_for(ctx, c->cap), 256, 0, ctx->stream, c->tab, c->cap, c->ctr,
                      c->reg.keys, c->reg.counts, c->reg.ords, cap_tab);
            ctx->launches++;
        }
        KjIrrRecord *d_irr = nullptr;
        KjIrrSpec irr_spec;
        // single-wait path with first-seen order: the records are sorted and their columns written on the device, queued
        // with everything else (the host sort of some 10^4 records was 0.15 ms between two kernels of the bench step)
        const bool spec_sort = spec && cap_irr > 1;
        if (cap_irr && c->irr_cap) {
            KJ_CUDA(ctx, kj_dmalloc(ctx, &d_irr, cap_irr * sizeof(KjIrrRecord)));
            const bool sort_on_device = getenv("KJ_SPEC_SORT_IRREGULAR") != nullptr;        // developer switch: keep first-seen order inside the handle
            if (spec_sort && sort_on_device) KJ_CUDA(ctx, cudaMemsetAsync(d_irr, 0xFF, cap_irr * sizeof(KjIrrRecord), ctx->stream));
            KJ_LAUNCH(kj_compact_irr_kernel, grid_for(ctx, c->irr_cap), 256, 0, ctx->stream, c->irr, c->irr_cap,
                      c->ctr, d_irr, cap_irr);
            ctx->launches++;
            if (spec_sort) {
                rc = irregular_spec_queue(c, d_irr, cap_irr, cap_tab, irr_spec, sort_on_device);
                if (rc) { kj_dfree(ctx, d_irr); return rc; }
            }
            if (cap_irr <= irr_device_sort_min() || cap_irr >= 0x7FFFFFFFull || spec) {
                irr_copied = spec ? std::min<uint64_t>(cap_irr, 8192) : cap_irr;
                c->irr_host.resize(cap_irr * sizeof(KjIrrRecord));
                cudaError_t e = cudaMemcpyAsync(c->irr_host.data(), spec_sort ? irr_spec.sorted : d_irr, irr_copied * sizeof(KjIrrRecord),
                                                cudaMemcpyDeviceToHost, ctx->stream);
                if (e != cudaSuccess) { kj_dfree(ctx, d_irr); irregular_spec_free(ctx, irr_spec); return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e)); }
            }
        }
        rc = pull_counters(c);           // the one wait: n_compact and (a small set of) irregular records are back with it
        if (rc == KJ_OK && spec && d_irr && c->h_ctr->n_irr_unique > irr_copied && c->h_ctr->n_irr_unique <= cap_irr) {
            // more irregular k-mers than came back with the counters: the rest of the compacted records
            cudaError_t e = cudaMemcpyAsync(c->irr_host.data() + irr_copied * sizeof(KjIrrRecord),
                                            (spec_sort ? irr_spec.sorted : d_irr) + irr_copied,
                                            (c->h_ctr->n_irr_unique - irr_copied) * sizeof(KjIrrRecord), cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) { kj_dfree(ctx, d_irr); irregular_spec_free(ctx, irr_spec); return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e)); }
        }
        irregular_spec_free(ctx, irr_spec);
        if (rc == KJ_OK && cap_irr > irr_device_sort_min() && cap_irr < 0x7FFFFFFFull && !spec) {
            rc = irregular_sorted_on_device(c, d_irr, c->h_ctr->n_irr_unique, c->h_ctr->n_unique + (c->h_ctr->special_count ? 1 : 0));
            irr_on_device = (rc == KJ_OK);
        }
        kj_dfree(ctx, d_irr);
        if (rc) return rc;
        if (!spec) break;
        const KjCounters &h = *c->h_ctr;
        const bool fits = !h.n_overflow && !h.n_irr_overflow && !h.error_flags && h.n_cand <= c->cand_cap &&
                          h.n_unique <= cap_tab && h.n_irr_unique <= cap_irr && h.n_unique * 2 <= c->cap;
        if (fits) {
            if (c->piece->timed) { rc = account_scan_time(c, c->piece->args.own_n, true); if (rc) return rc; c->piece->timed = false; }
            c->pending = false;
            irr_on_device = spec_sort && cap_irr && c->irr_cap;
            break;
        }
        spec = false;                    // the long way: settle (retries, growth), then compact with the exact sizes
    }
    const uint64_t n_tab = c->h_ctr->n_unique;
    const uint64_t n_reg = n_tab + (c->h_ctr->special_count ? 1 : 0);
    const uint64_t n_irr = c->h_ctr->n_irr_unique;
    const uint64_t q = n_reg + n_irr;
    c->irr_host.resize(n_irr * sizeof(KjIrrRecord));
    std::vector<uint64_t> &tail_counts = c->tail_counts, &tail_ords = c->tail_ords;   // special + irregular entries
    tail_counts.clear(); tail_ords.clear();
    if (c->h_ctr->special_count) {
        tail_counts.push_back(c->h_ctr->special_count);
        tail_ords.push_back(c->h_ctr->special_ord);
    }
    if (n_irr && irr_on_device) {
        // sorted on the device, count / ordinal columns written there: only the special key is left for the host
    } else if (n_irr) {
        // order of the irregular entries: by first-seen ordinal (distinct per entry), so that it does not
        // depend on the table size.  LSD radix sort of (ordinal, index), then one permutation pass: the
        // stress configs have ~10^6 of these.  Without ordinals (KJ_F_NO_ORDER) they stay in slot order.
        KjIrrRecord *ir = reinterpret_cast<KjIrrRecord *>(c->irr_host.data());
        if (c->order && n_irr > 1) {
            std::vector<uint64_t> k0(n_irr), k1(n_irr);
            std::vector<uint32_t> p0(n_irr), p1(n_irr);
            for (uint64_t i = 0; i < n_irr; ++i) { k0[i] = ir[i].ord; p0[i] = (uint32_t)i; }
            std::vector<uint64_t> *ks = &k0, *kd = &k1;
            std::vector<uint32_t> *ps = &p0, *pd = &p1;
            for (int pass = 0; pass < 8; ++pass) {
                size_t hist[257] = {0};
                const int sh = 8 * pass;
                for (uint64_t i = 0; i < n_irr; ++i) hist[(((*ks)[i]) >> sh & 0xFF) + 1]++;
                bool trivial = false;
                for (int b = 1; b <= 256; ++b) if (hist[b] == n_irr) trivial = true;
                if (trivial) continue;
                for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
                for (uint64_t i = 0; i < n_irr; ++i) {
                    const size_t at = hist[((*ks)[i]) >> sh & 0xFF]++;
                    (*kd)[at] = (*ks)[i];
                    (*pd)[at] = (*ps)[i];
                }
                std::swap(ks, kd);
                std::swap(ps, pd);
            }
            std::vector<uint8_t> sorted(c->irr_host.size());
            KjIrrRecord *dst = reinterpret_cast<KjIrrRecord *>(sorted.data());
            for (uint64_t i = 0; i < n_irr; ++i) dst[i] = ir[(*ps)[i]];
            c->irr_host.swap(sorted);
            ir = reinterpret_cast<KjIrrRecord *>(c->irr_host.data());
        }
        for (uint64_t i = 0; i < n_irr; ++i) { tail_counts.push_back(ir[i].count); tail_ords.push_back(ir[i].ord); }
    }
    if (!tail_counts.empty()) {
        // keys of the tail entries: KJ_EMPTY for the special key, unused (KJ_EMPTY) for irregular ones.  The
        // copies are stream-ordered; the vectors belong to the handle, so nothing has to wait for them here
        KJ_CUDA(ctx, cudaMemsetAsync(c->reg.keys + n_tab, 0xFF, tail_counts.size() * 8, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(c->reg.counts + n_tab, tail_counts.data(), tail_counts.size() * 8,
                                     cudaMemcpyHostToDevice, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(c->reg.ords + n_tab, tail_ords.data(), tail_ords.size() * 8,
                                     cudaMemcpyHostToDevice, ctx->stream));
    }
    c->reg.n = q;
    c->reg.n_reg = n_reg;
    c->reg.n_tab = n_tab;
    if (c->h_ctr->n_compact != n_tab) return kj_fail(ctx, KJ_E_CUDA, "internal: compaction count mismatch");
    c->lines = c->h_ctr->carry_lines[c->parity];
    // sum of the sequence-line lengths (lines with index 1 mod 4).  The filter kernel accumulates
    // signed newline offsets that telescope; the two ends of this handle's byte range close the sum:
    // a range that starts inside a sequence line does not own the bytes before it, a range that ends
    // inside one (unterminated last line, or a shard cut) owns the bytes up to its end.
    long long bases = (long long)c->h_ctr->n_bases;
    if ((c->use_filter || c->use_dense) && c->consumed && (c->flags & KJ_F_COUNT_BASES)) {
        if ((c->base_line & 3ull) == 1ull) bases -= (long long)c->base_col;
        if ((c->lines & 3ull) == 1ull) bases += (long long)c->voff;
    }
    c->bases = (c->flags & KJ_F_COUNT_BASES) ? (uint64_t)bases : 0;
    // a non-empty unterminated tail is one more line (lib/kmers.js:130-136)
    if (c->consumed && c->h_ctr->carry_last[c->parity] != c->voff) c->lines += 1;
    c->occurrences = c->h_ctr->n_occ;
    if (c->exchange_totals) {     // this handle holds the k-mers one rank owns: the totals are the whole job's
        c->lines = c->h_ctr->x_lines; c->bases = c->h_ctr->x_bases;
        c->occurrences = c->h_ctr->x_occ; c->bytes_read = c->h_ctr->x_bytes;
    }
    c->finished = true;
    return KJ_OK;
}

int kj_counts_check_finished(const kj_counts *c) {
    if (!c) return KJ_E_INVALID;
    if (!c->finished) return kj_fail(c->ctx, KJ_E_STATE, "kj_counts_finish has not been called");
    return KJ_OK;
}

extern "C" uint64_t kj_counts_size(const kj_counts *c) { return c ? c->reg.n : 0; }
extern "C" uint64_t kj_counts_lines(const kj_counts *c) { return c ? c->lines : 0; }
extern "C" uint64_t kj_counts_bases(const kj_counts *c) { return c ? c->bases : 0; }
extern "C" uint64_t kj_counts_bytes_read(const kj_counts *c) { return c ? (c->bytes_read ? c->bytes_read : c->consumed) : 0; }
extern "C" uint64_t kj_counts_occurrences(const kj_counts *c) { return c ? c->occurrences : 0; }
extern "C" uint64_t kj_counts_irregular_size(const kj_counts *c) {
    return c ? c->irr_host.size() / sizeof(KjIrrRecord) : 0;
}

void kj_decode_key(uint64_t key, uint32_t k, uint8_t *out) {
    static const char L[4] = {'A', 'C', 'T', 'G'};   // code = (byte >> 1) & 3
    for (uint32_t i = 0; i < k; ++i) out[i] = (uint8_t)L[(key >> (2 * (k - 1 - i))) & 3];
}

// export_perm[o] = query index (position in the compact arrays) of the o-th key in
// first-insertion order = ascending first-seen ordinal (lib/kmers.js:46-54,95)
static int build_export_perm(kj_counts *c, std::vector<uint64_t> &hk) {
    kj_ctx *ctx = c->ctx;
    const uint64_t q = c->reg.n, n_reg = c->reg.n_reg;
    std::vector<uint64_t> ho(q);
    hk.resize(q);
    if (q) {
        KJ_CUDA(ctx, cudaMemcpyAsync(hk.data(), c->reg.keys, q * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(ho.data(), c->reg.ords, q * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const KjIrrRecord *ir = reinterpret_cast<const KjIrrRecord *>(c->irr_host.data());
    c->export_perm.resize(q);
    for (uint64_t i = 0; i < q; ++i) c->export_perm[i] = i;
    if (c->order) {
        // first-seen ordinals are distinct: LSD radix sort of (ordinal, index), skipping the byte
        // positions on which all ordinals agree (the export sits on the end-to-end path)
        std::vector<uint64_t> k0(ho), k1(q), p1(q);
        std::vector<uint64_t> *ks = &k0, *kd = &k1, *ps = &c->export_perm, *pd = &p1;
        for (int pass = 0; pass < 8; ++pass) {
            size_t hist[257] = {0};
            const int sh = 8 * pass;
            for (uint64_t i = 0; i < q; ++i) hist[(((*ks)[i]) >> sh & 0xFF) + 1]++;
            bool trivial = false;
            for (int b = 1; b <= 256; ++b) if (hist[b] == q) trivial = true;
            if (trivial) continue;
            for (int b = 0; b < 256; ++b) hist[b + 1] += hist[b];
            for (uint64_t i = 0; i < q; ++i) {
                const size_t at = hist[((*ks)[i]) >> sh & 0xFF]++;
                (*kd)[at] = (*ks)[i];
                (*pd)[at] = (*ps)[i];
            }
            std::swap(ks, kd);
            std::swap(ps, pd);
        }
        if (ps != &c->export_perm) c->export_perm = *ps;
    } else {
        // KJ_F_NO_ORDER: every ordinal is ~0; order by key so the export is deterministic
        std::sort(c->export_perm.begin(), c->export_perm.end(), [&](uint64_t x, uint64_t y) {
            bool xr = x < n_reg, yr = y < n_reg;
            if (xr && yr) return hk[x] < hk[y];
            if (xr != yr) return xr;
            return memcmp(ir[x - n_reg].key, ir[y - n_reg].key, 32) < 0;
        });
    }
    return KJ_OK;
}

// kj_counts_export with ordinals: radix sort of (ordinal, index) and the gather on the device, one copy
// back per output array (the export sits on the end-to-end path: 174 k keys took 9.6 ms on the host)
static int export_on_device(kj_counts *c, uint8_t *keys, uint32_t *key_len, uint64_t *counts) {
    kj_ctx *ctx = c->ctx;
    const uint64_t q = c->reg.n, n_reg = c->reg.n_reg, n_irr = q - n_reg;
    if (!q) return KJ_OK;
    uint32_t *d_idx = nullptr, *d_perm = nullptr, *d_len = nullptr;
    uint64_t *d_ord = nullptr, *d_cnt = nullptr, *d_irr_pos = nullptr;
    uint4 *d_keys = nullptr;
    void *d_tmp = nullptr;
    cudaError_t e = kj_dmalloc(ctx, &d_idx, q * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_perm, q * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_ord, q * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_keys, q * 32);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_len, q * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_cnt, q * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_irr_pos, std::max<uint64_t>(n_irr, 1) * 8);
    std::vector<uint64_t> irr_pos(n_irr);
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_export_iota_kernel, grid_for(ctx, q), 256, 0, ctx->stream, d_idx, q);
        ctx->launches++;
#ifndef KJ_CPU_EMU
        size_t tmp_bytes = 0;
        e = cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, c->reg.ords, d_ord, d_idx, d_perm, (int)q, 0, 64, ctx->stream);
        if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_tmp, std::max<size_t>(tmp_bytes, 16));
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, c->reg.ords, d_ord, d_idx, d_perm, (int)q, 0, 64, ctx->stream);
        ctx->launches += 4;      // cub: histogram + one pass per non-trivial digit (counted as a lower bound)
#else
        {   // tools/cuemu: "device" memory is host memory
            std::vector<uint32_t> p(q);
            for (uint64_t i = 0; i < q; ++i) p[i] = (uint32_t)i;
            std::stable_sort(p.begin(), p.end(), [&](uint32_t x, uint32_t y) { return c->reg.ords[x] < c->reg.ords[y]; });
            memcpy(d_perm, p.data(), q * 4);
        }
#endif
    }
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_export_gather_kernel, grid_for(ctx, q), 256, 0, ctx->stream, d_perm, c->reg.keys, c->reg.counts, q, n_reg,
                  c->k, d_keys, d_len, d_cnt, d_irr_pos);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(keys, d_keys, q * 32, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(key_len, d_len, q * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(counts, d_cnt, q * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && n_irr) e = cudaMemcpyAsync(irr_pos.data(), d_irr_pos, n_irr * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    kj_dfree(ctx, d_idx); kj_dfree(ctx, d_perm); kj_dfree(ctx, d_ord); kj_dfree(ctx, d_keys); kj_dfree(ctx, d_len);
    kj_dfree(ctx, d_cnt); kj_dfree(ctx, d_irr_pos); kj_dfree(ctx, d_tmp);
    if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, std::string("kj_counts_export: ") + cudaGetErrorString(e));
    const KjIrrRecord *ir = reinterpret_cast<const KjIrrRecord *>(c->irr_host.data());
    for (uint64_t j = 0; j < n_irr; ++j) {
        const uint64_t o = irr_pos[j];
        memcpy(keys + 32 * o, ir[j].key, 32);
        key_len[o] = (uint32_t)ir[j].len;
    }
    return KJ_OK;
}

extern "C" int kj_counts_export(kj_counts *c, uint8_t *keys, uint32_t *key_len, uint64_t *counts) {
    int rc = kj_counts_check_finished(c);
    if (rc) return rc;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t q = c->reg.n, n_reg = c->reg.n_reg;
    if (c->order && q < 0x7FFFFFFFull) return export_on_device(c, keys, key_len, counts);
    std::vector<uint64_t> hk, hc(q);     // KJ_F_NO_ORDER: ordered by key on the host (deterministic, off the hot path)
    rc = build_export_perm(c, hk);
    if (rc) return rc;
    if (q) {
        KJ_CUDA(ctx, cudaMemcpyAsync(hc.data(), c->reg.counts, q * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    const KjIrrRecord *ir = reinterpret_cast<const KjIrrRecord *>(c->irr_host.data());
    for (uint64_t o = 0; o < q; ++o) {
        uint64_t i = c->export_perm[o];
        uint8_t *dst = keys + 32 * o;
        memset(dst, 0, 32);
        if (i < n_reg) {
            kj_decode_key(hk[i], c->k, dst);
            key_len[o] = c->k;
        } else {
            const KjIrrRecord &r = ir[i - n_reg];
            memcpy(dst, r.key, 32);
            key_len[o] = (uint32_t)r.len;
        }
        counts[o] = hc[i];
    }
    return KJ_OK;
}

int kj_counts_export_rank(kj_counts *c, const std::vector<uint64_t> **rank) {
    int rc = kj_counts_check_finished(c);
    if (rc) return rc;
    const uint64_t q = c->reg.n;
    if (c->export_perm.size() != q) {
        std::vector<uint64_t> hk;
        rc = build_export_perm(c, hk);
        if (rc) return rc;
    }
    if (c->export_rank.size() != q) {
        c->export_rank.resize(q);
        for (uint64_t o = 0; o < q; ++o) c->export_rank[c->export_perm[o]] = o;
    }
    *rank = &c->export_rank;
    return KJ_OK;
}

// alive flags in export order (1 = still in the map; kj_wta_next clears the winner's k-mers,
// mirroring kmerMap.delete of lib/kmerFinderClient.js:220-230)
extern "C" int kj_counts_alive(kj_counts *c, uint8_t *alive) {
    int rc = kj_counts_check_finished(c);
    if (rc) return rc;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t q = c->reg.n;
    if (c->export_perm.size() != q) {
        std::vector<uint64_t> hk;
        rc = build_export_perm(c, hk);
        if (rc) return rc;
    }
    std::vector<uint8_t> ha(q);
    if (q) {
        KJ_CUDA(ctx, cudaMemcpyAsync(ha.data(), c->reg.alive, q, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    for (uint64_t o = 0; o < q; ++o) alive[o] = ha[c->export_perm[o]];
    return KJ_OK;
}

extern "C" void kj_counts_free(kj_counts *c) {
    if (!c) return;
    kj_ctx *ctx = c->ctx;
    if (ctx) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        cudaSetDevice(ctx->device);
        if (c->pending) {
            cudaStreamSynchronize(ctx->stream);                 // a kernel may still be reading the caller's buffer
            // a piece nobody settled (the fixed-capacity exchange partitions straight from the table): its kernel times
            // are still in the context's events
            if (c->piece && c->piece->timed && ctx->timers_on) { account_scan_time(c, c->piece->args.own_n, true); c->piece->timed = false; }
        }
        free_table(ctx, c->tab);
        free_irr(ctx, c->irr);
        kj_dfree(ctx, c->ovf.rec); kj_dfree(ctx, c->ovf.irr_rec);
        kj_dfree(ctx, c->ctr);
        kj_pinned_put(ctx, c->h_ctr);
        kj_dfree(ctx, c->tile_mem); kj_dfree(ctx, c->tile_cnt); kj_dfree(ctx, c->scan_tmp);
        kj_dfree(ctx, c->cand);
        drop_compact(c);
        kj_dfree(ctx, c->part_rec);
    }
    delete c->piece;
    delete c;
}

// ------------------------------------------------------------------------------------ exchange

extern "C" int kj_counts_partition(kj_counts *c, uint32_t n_parts, const void **dev_records,
                                   uint64_t *part_sizes) {
    int rc = kj_counts_check_finished(c);
    if (rc) return rc;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!n_parts || n_parts > 1024 || !dev_records || !part_sizes)
        return kj_fail(ctx, KJ_E_INVALID, "kj_counts_partition arguments");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t n = c->reg.n_reg;   // regular entries only; irregular ones travel as host records
    unsigned long long *d_hist = nullptr;
    KJ_CUDA(ctx, kj_dmalloc(ctx, &d_hist, n_parts * 8));
    KJ_CUDA(ctx, cudaMemsetAsync(d_hist, 0, n_parts * 8, ctx->stream));
    std::vector<unsigned long long> hist(n_parts, 0);
    if (n) {
        KJ_LAUNCH(kj_part_hist_kernel, grid_for(ctx, n), 256, 0, ctx->stream, c->reg.keys, n, n_parts, d_hist);
        ctx->launches++;
    }
    cudaError_t e = cudaMemcpyAsync(hist.data(), d_hist, n_parts * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { kj_dfree(ctx, d_hist); return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e)); }
    std::vector<unsigned long long> cursor(n_parts, 0);
    unsigned long long run = 0;
    for (uint32_t p = 0; p < n_parts; ++p) { cursor[p] = run; run += hist[p]; part_sizes[p] = hist[p]; }
    if (n > c->part_cap || !c->part_rec) {
        kj_dfree(ctx, c->part_rec);
        c->part_rec = nullptr; c->part_cap = 0;
        e = kj_dmalloc(ctx, &c->part_rec, std::max<uint64_t>(n, 1) * sizeof(KjRecord));
        if (e != cudaSuccess) { kj_dfree(ctx, d_hist); return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e)); }
        c->part_cap = std::max<uint64_t>(n, 1);
    }
    if (n) {
        e = cudaMemcpyAsync(d_hist, cursor.data(), n_parts * 8, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) {
            KJ_LAUNCH(kj_part_scatter_kernel, grid_for(ctx, n), 256, 0, ctx->stream, c->reg.keys, c->reg.counts,
                      c->reg.ords, n, n_parts, d_hist, reinterpret_cast<KjRecord *>(c->part_rec));
            ctx->launches++;
            e = cudaStreamSynchronize(ctx->stream);
        }
    }
    kj_dfree(ctx, d_hist);
    if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
    *dev_records = c->part_rec;
    return KJ_OK;
}

extern "C" uint64_t kj_segment_bytes(uint32_t cap_reg, uint32_t cap_irr) {
    const uint64_t b = sizeof(KjSegHeader) + (uint64_t)cap_reg * sizeof(KjRecord) + (uint64_t)cap_irr * sizeof(KjIrrRecord);
    return (b + 255) / 256 * 256;
}

// Fixed-capacity exchange, sender side: the table as it stands after the adds (no kj_counts_finish, no host wait) is
// scattered by owner into n_parts segments of kj_segment_bytes(cap_reg, cap_irr) bytes each.  Whatever does not fit --
// or a count that is not complete -- shows in the headers and makes the receivers' kj_counts_finish fail with
// KJ_E_RANGE, so that the caller can fall back to the two-phase exchange (kj_counts_partition).
extern "C" int kj_counts_partition_segments(kj_counts *c, uint32_t n_parts, void *dev_segments, uint32_t cap_reg,
                                            uint32_t cap_irr) {
    if (!c || !dev_segments || !n_parts || n_parts > 1024 || !cap_reg) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    KjSegArgs a{};
    a.seg = reinterpret_cast<uint8_t *>(dev_segments);
    a.seg_bytes = kj_segment_bytes(cap_reg, cap_irr);
    a.n_parts = n_parts; a.cap_reg = cap_reg; a.cap_irr = cap_irr;
    a.parity = c->parity; a.voff = c->voff; a.consumed = c->consumed;
    a.count_bases = ((c->flags & KJ_F_COUNT_BASES) && (c->use_filter || c->use_dense)) ? 1u : 0u;
    a.bases_fix = ((c->base_line & 3ull) == 1ull) ? -(long long)c->base_col : 0;
    a.bases_tail = 1;
    if ((c->flags & KJ_F_COUNT_BASES) && !a.count_bases) a.count_bases = 2;     // line kernel: n_bases is already the sum
    a.cand_cap = c->use_filter ? c->cand_cap : ~0ull;
    for (uint32_t p = 0; p < n_parts; ++p)
        KJ_CUDA(ctx, cudaMemsetAsync(a.seg + (uint64_t)p * a.seg_bytes, 0, sizeof(KjSegHeader), ctx->stream));
    KJ_LAUNCH(kj_seg_totals_kernel, (n_parts + 127) / 128, 128, 0, ctx->stream, a, c->ctr);
    if (c->cap || c->irr_cap)
        KJ_LAUNCH(kj_seg_scatter_kernel, grid_for(ctx, std::max(c->cap, c->irr_cap)), 256, 0, ctx->stream, a, c->tab, c->cap,
                  c->irr, c->irr_cap, c->ctr);
    ctx->launches += 2;
    KJ_CUDA(ctx, cudaGetLastError());
    return KJ_OK;
}

// receiver side: merge n_parts segments into this (fresh) handle; stream-ordered, the totals and the overflow flags
// surface in kj_counts_finish
extern "C" int kj_counts_merge_segments(kj_counts *c, const void *dev_segments, uint32_t n_parts, uint32_t cap_reg,
                                        uint32_t cap_irr) {
    if (!c || !dev_segments || !n_parts || n_parts > 1024 || !cap_reg) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = settle_filter(c);
    if (rc) return rc;
    const uint64_t bound = (uint64_t)n_parts * cap_reg;
    rc = grow_table(c, 2 * (c->h_ctr->n_unique + (c->capacity_hint ? std::min<uint64_t>(c->capacity_hint, bound) : bound)));
    if (rc) return rc;
    c->irr_bound += (uint64_t)n_parts * cap_irr;
    rc = grow_irr(c, 2 * c->irr_bound);
    if (rc) return rc;
    KjSegArgs a{};
    a.seg = reinterpret_cast<uint8_t *>(const_cast<void *>(dev_segments));
    a.seg_bytes = kj_segment_bytes(cap_reg, cap_irr);
    a.n_parts = n_parts; a.cap_reg = cap_reg; a.cap_irr = cap_irr;
    KJ_LAUNCH(kj_seg_merge_kernel, grid_for(ctx, std::max<uint64_t>(bound, n_parts)), 256, 0, ctx->stream, a, c->tab, c->irr, c->ctr);
    ctx->launches++;
    KJ_CUDA(ctx, cudaGetLastError());
    c->finished = false;
    c->exchange_totals = true;
    return KJ_OK;
}

extern "C" int kj_counts_merge_records(kj_counts *c, const void *dev_records, uint64_t n) {
    if (!c) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!n) return KJ_OK;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = settle_filter(c);
    if (rc) return rc;
    rc = pull_counters(c);
    if (rc) return rc;
    rc = grow_table(c, std::max<uint64_t>(2 * (c->h_ctr->n_unique + n), c->capacity_hint * 2));
    if (rc) return rc;
    KJ_LAUNCH(kj_merge_records_kernel, grid_for(ctx, n), 256, 0, ctx->stream, c->tab, c->ctr,
              reinterpret_cast<const KjRecord *>(dev_records), n);
    ctx->launches++;
    rc = pull_counters(c);
    if (rc) return rc;
    c->finished = false;
    return check_device_errors(c);
}

// same as kj_counts_merge_records for records in HOST memory (a k-mer map that arrives as JSON)
extern "C" int kj_counts_merge_host_records(kj_counts *c, const void *host_records, uint64_t n) {
    if (!c || (n && !host_records)) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!n) return KJ_OK;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    KjRecord *d = nullptr;
    KJ_CUDA(ctx, kj_dmalloc(ctx, &d, n * sizeof(KjRecord)));
    cudaError_t e = cudaMemcpyAsync(d, host_records, n * sizeof(KjRecord), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    int rc = e == cudaSuccess ? kj_counts_merge_records(c, d, n) : kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
    kj_dfree(ctx, d);
    return rc;
}

extern "C" int kj_counts_irregular_export(kj_counts *c, void *host_records) {
    int rc = kj_counts_check_finished(c);
    if (rc) return rc;
    if (!c->irr_host.empty()) memcpy(host_records, c->irr_host.data(), c->irr_host.size());
    return KJ_OK;
}

extern "C" int kj_counts_irregular_merge(kj_counts *c, const void *host_records, uint64_t n) {
    return kj_counts_irregular_merge_part(c, host_records, n, 0, 1);
}

extern "C" int kj_counts_irregular_merge_part(kj_counts *c, const void *host_records, uint64_t n, uint32_t part,
                                              uint32_t n_parts) {
    if (!c) return KJ_E_INVALID;
    if (n_parts == 0 || part >= n_parts) return kj_fail(c->ctx, KJ_E_INVALID, "part / n_parts");
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    // the records this part owns are picked on the host: only they cross PCIe and size the table
    std::vector<KjIrrRecord> mine;
    if (n_parts > 1) {
        const KjIrrRecord *all = reinterpret_cast<const KjIrrRecord *>(host_records);
        for (uint64_t i = 0; i < n; ++i)
            if (kj_owner_bytes(all[i].key, (uint32_t)all[i].len, n_parts) == part) mine.push_back(all[i]);
        host_records = mine.data();
        n = mine.size();
        n_parts = 1;
    }
    if (!n) return KJ_OK;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc0 = settle_filter(c);
    if (rc0) return rc0;
    // no round trip for the current size: every settled scan ends with the counters pulled, merges add at most n
    c->irr_bound += n;
    int rc = grow_irr(c, 2 * c->irr_bound);
    if (rc) return rc;
    KjIrrRecord *d = nullptr;
    KJ_CUDA(ctx, kj_dmalloc(ctx, &d, n * sizeof(KjIrrRecord)));
    cudaError_t e = cudaMemcpyAsync(d, host_records, n * sizeof(KjIrrRecord), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_merge_irr_kernel, grid_for(ctx, n), 256, 0, ctx->stream, c->irr, c->ctr, d, n, part, n_parts);
        ctx->launches++;
        e = cudaGetLastError();
    }
    // no synchronisation here: the copy from pageable host memory has been staged when cudaMemcpyAsync
    // returns, the free is stream-ordered, and kj_counts_finish pulls the counters and the error flags
    kj_dfree(ctx, d);
    if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
    c->finished = false;
    return KJ_OK;
}

// set totals that the exchange cannot reconstruct (lines / bases / occurrences / bytes of the
// whole job when this handle holds only the k-mers one rank owns)
extern "C" int kj_counts_set_totals(kj_counts *c, uint64_t lines, uint64_t bases, uint64_t occurrences,
                                    uint64_t bytes_read) {
    if (!c) return KJ_E_INVALID;
    c->lines = lines; c->bases = bases; c->occurrences = occurrences; c->bytes_read = bytes_read;
    return KJ_OK;
}

extern "C" int kj_count_newlines(kj_ctx *ctx, const uint8_t *buf, uint64_t n, int mem_kind,
                                 uint64_t *n_newlines, uint64_t *last_newline_plus1) {
    if (!ctx || (n && !buf)) return kj_fail(ctx, KJ_E_INVALID, "kj_count_newlines arguments");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    unsigned long long *d_out = nullptr;
    KJ_CUDA(ctx, kj_dmalloc(ctx, &d_out, 16));
    KJ_CUDA(ctx, cudaMemsetAsync(d_out, 0, 16, ctx->stream));
    unsigned long long total[2] = {0, 0};
    int rc = KJ_OK;
    if (mem_kind == KJ_MEM_DEVICE) {
        if ((uintptr_t)buf & 15) rc = kj_fail(ctx, KJ_E_INVALID, "device buffers must be 16-byte aligned");
        else if (n) {
            KJ_LAUNCH(kj_newline_kernel, grid_for(ctx, (n + 15) / 16), 256, 0, ctx->stream, buf, n, d_out);
            ctx->launches++;
        }
        if (rc == KJ_OK) {
            cudaError_t e = cudaMemcpyAsync(total, d_out, 16, cudaMemcpyDeviceToHost, ctx->stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
            if (e != cudaSuccess) rc = kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
        }
    } else {
        rc = kj_fail(ctx, KJ_E_INVALID, "kj_count_newlines takes device buffers (stage the range first)");
    }
    kj_dfree(ctx, d_out);
    if (rc) return rc;
    if (n_newlines) *n_newlines = total[0];
    if (last_newline_plus1) *last_newline_plus1 = total[1];
    return KJ_OK;
}

static bool ascii_regular(const uint8_t *kmer, uint32_t len) {
    for (uint32_t i = 0; i < len; ++i) if (!kj_is_acgt(kmer[i])) return false;
    return len >= 1 && len <= 32;
}

extern "C" uint32_t kj_owner(const uint8_t *kmer, uint32_t len, uint32_t n_parts) {
    if (!kmer || !n_parts) return 0;
    if (!ascii_regular(kmer, len)) {               // byte-string (side-table) k-mers: hash of the padded bytes
        if (len > 32) return 0;
        uint8_t pad[32] = {0};
        memcpy(pad, kmer, len);
        return kj_owner_bytes(pad, len, n_parts);
    }
    uint64_t key = 0;
    for (uint32_t i = 0; i < len; ++i) key = (key << 2) | kj_code(kmer[i]);
    return kj_owner_key(key, n_parts);
}
