// kj_scan_warp.cuh -- the filter path (step = 1, 1 <= |prefix| <= k): the hot kernels of the library.
//
//   kj_warp_filter_kernel   every warp is its own pipeline, no warp ever waits for another one.  A warp
//                           streams tiles of 31 rows x 128 bytes; the TMA engine brings 32 rows at a time
//                           (one 2-D tensor-map copy with the 128-byte swizzle, SASS UTMALDG.2D) into the
//                           warp's private ring of stages, so that every lane reads ITS OWN contiguous 128
//                           bytes (8 chunks of 16) without bank conflicts: the code words of a lane's
//                           neighbours in the stream are its own registers, and only the word behind its
//                           last chunk comes from the next lane (one shuffle per 8 chunks; lane 31 converts
//                           the row behind the tile for lane 30 and owns nothing).
//                           Per chunk: 2-bit code word, newline mask, and a bit-parallel search in code space,
//                           16 positions per 32-bit operation, for the places where the prefix starts (forward
//                           windows start there) and where complement(prefix) starts (reverse-strand windows
//                           start k - |prefix| bytes before): lib/kmers.js:88-100,151-155 as an exact superset
//                           filter.  Both strands share the shifted words.  Chunks with a candidate go to a queue
//                           per lane (a predicated store, no atomic, no branch) and are drained by the whole
//                           warp every few tiles into 16-byte entries {chunk, candidate lanes, '\n' before it in the
//                           tile, its own '\n' mask, distance back to the line start}.  The tile's newline count goes to
//                           tile_cnt[]; nothing here needs the number of lines before the tile.
//   (exclusive scan of tile_cnt -> tile_excl, cub::DeviceScan; the record FSM of lib/kmers.js:151-163 is
//    "line index mod 4" over the whole stream)
//   kj_resolve_kernel       one thread per entry, the whole GPU: line index of the candidate = lines before the launch +
//                           tile_excl + in-tile count -> keep iff 1 mod 4 (the reference's i === 1); the window's bytes
//                           (prefix, no '\n' inside, alphabet) -> 2k-bit key; first-seen ordinal; hash-table update.  An
//                           emission that finds no slot marks its entry for a retry pass after the host has grown the
//                           table: nothing is ever dropped, whatever the input looks like.
// Where the byte check lives was measured three ways on 10 M reads (3.46 GB): in the resolve kernel (this file: the windows
// come back from DRAM, 1-2 sectors each, but 64 warps per SM hide it); split into a compacting filter kernel and a dense
// emit kernel (0.22 + 0.19 ms: the append counter and the second pass cost more than the divergence they removed); inside the
// drain of the scan kernel, where the bytes are still in L2 (scan 0.77 -> 1.20 ms: 16 warps per SM cannot hide the L2
// latency, and the key extraction competes with the search for the ALU pipe the scan is bound by).
#pragma once
#include "kj_scan.cuh"

#define KJ_WT_OWN_ROWS 31u
#define KJ_WT_BYTES (KJ_WT_OWN_ROWS * 128u)      // 3968 bytes = 248 chunks owned by a tile
#define KJ_WT_CHUNKS (KJ_WT_BYTES / 16u)
#define KJ_WT_STAGE_BYTES 4096u                  // what one TMA copy brings: the tile and the row behind it
#define KJ_WT_STAGES 2
#define KJ_WT_RING 4u                            // tiles whose newline bitmaps stay in shared memory
#define KJ_WT_PQCAP 12u                          // queue entries per LANE; a tile adds at most 8 to a lane
#define KJ_WT_BLOCK 128u                         // entry slots a warp reserves at a time (one global atomic, taken ahead)
#define KJ_WT_WARPS 8
#define KJ_WT_THREADS (KJ_WT_WARPS * 32)
#define KJ_ENT_NODIST 0xFFFFu
// entry (16 bytes, one per chunk with candidates) = {chunk index in the launch (40 bits) | '\n' of the tile before the chunk
// (13) << 40 | retry << 63, candidate lanes (bit 2p: forward window at byte p of the chunk, 2p + 1: reverse),
// the chunk's own '\n' mask (16) | distance from the chunk back to the byte after the last '\n' of the tile << 16
// (KJ_ENT_NODIST: the line starts before the tile)}.  A blank entry (unused slot of a reserved block) has no candidate lanes.
#define KJ_ENT_RETRY32 0x80000000u               // in the entry's second 32-bit word

struct __align__(16) KjWarpSmem {
    uint4 bitmap[KJ_WT_RING][32];                // lane l: the '\n' masks of its 8 chunks, 16 bits each, in stream order
    uint32_t lanepre[KJ_WT_RING][32];            // '\n' of the tile before lane l's first byte
    uint2 pq[KJ_WT_PQCAP][32];                   // a queue per lane, entry j of lane l at pq[j][l]: {candidate lanes of a chunk
                                                 // (bit 2p: forward window at byte p, 2p + 1: reverse), ring slot << 3 | chunk of the lane}
};
#define KJ_WT_SMEM_BYTES (KJ_WT_WARPS * KJ_WT_STAGES * KJ_WT_STAGE_BYTES + KJ_WT_WARPS * KJ_WT_STAGES * 8 + \
                          KJ_WT_WARPS * sizeof(KjWarpSmem) + 1024)

// ----------------------------------------------------------------------------- TMA tensor copy
#ifdef KJ_CPU_EMU
struct KjTensorMap { const uint8_t *base; uint64_t rows; };      // tools/cuemu: rows of 128 readable bytes
__device__ __forceinline__ void kj_tma_tile(void *dst, const KjTensorMap *tm, uint32_t row, uint64_t *bar) {
    uint8_t *d = reinterpret_cast<uint8_t *>(dst);
    for (uint32_t r = 0; r < 32; ++r)
        for (uint32_t c = 0; c < 8; ++c) {
            uint8_t *to = d + r * 128u + ((c ^ (r & 7u)) << 4);            // the 128-byte swizzle
            if ((uint64_t)row + r < tm->rows) memcpy(to, tm->base + ((uint64_t)row + r) * 128u + c * 16u, 16);
            else memset(to, 0, 16);                                          // out of bounds reads as zero
        }
    kj_bar_arrive(bar);
}
__device__ __forceinline__ uint4 kj_lds128(const void *p) { return *reinterpret_cast<const uint4 *>(p); }
typedef const uint8_t *kj_saddr;
__device__ __forceinline__ kj_saddr kj_saddr_of(const void *p) { return reinterpret_cast<const uint8_t *>(p); }
#else
#include <cuda.h>
typedef CUtensorMap KjTensorMap;
// one thread: announce the bytes on the barrier, start the copy of rows [row, row + 32) of the tensor
__device__ __forceinline__ void kj_tma_tile(void *dst, const KjTensorMap *tm, uint32_t row, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kj_smem_u32(bar)), "r"(KJ_WT_STAGE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(kj_smem_u32(dst)), "l"(tm), "r"(0), "r"(row), "r"(kj_smem_u32(bar)) : "memory");
}
typedef uint32_t kj_saddr;
__device__ __forceinline__ kj_saddr kj_saddr_of(const void *p) { return kj_smem_u32(p); }
__device__ __forceinline__ uint4 kj_lds128(kj_saddr a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
#endif

// newline mask of 16 bytes with the two constants of the zero-byte test held in registers: (w ^ a) & b is
// then ONE LOP3 (with immediates the compiler needs two, an immediate operand being 32 bits per instruction)
__device__ __forceinline__ uint32_t kj_nl_flags_r(uint32_t w, uint32_t c0a, uint32_t c7f) {
    const uint32_t t = ((w ^ c0a) & c7f) + c7f;
    return ~(t | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t kj_nl16_r(const uint4 v, uint32_t c0a, uint32_t c7f) {
#if defined(__CUDA_ARCH__)
    uint32_t lo = __dp4a(kj_nl_flags_r(v.x, c0a, c7f), 0x08040201u, 0u);
    lo = __dp4a(kj_nl_flags_r(v.y, c0a, c7f), 0x80402010u, lo);
    uint32_t hi = __dp4a(kj_nl_flags_r(v.z, c0a, c7f), 0x08040201u, 0u);
    hi = __dp4a(kj_nl_flags_r(v.w, c0a, c7f), 0x80402010u, hi);
    return (lo >> 7) | ((hi << 1) & 0xFF00u);
#else
    (void)c0a; (void)c7f;
    return kj_nl16(v.x, v.y, v.z, v.w);
#endif
}

// candidate lanes of one chunk from its code word and the two behind it: bit 2p = the forward window at byte p passes the
// code-space filter, bit 2p + 1 = the reverse-strand window.  live = 0x55555555 with KJ_F_FORWARD_ONLY, else all ones.
// Candidate lanes of one chunk from its code word and the one behind it.  Bit 2p: the prefix starts at byte p of the
// chunk in code space (a forward window starts there); bit 2p + 1: complement(prefix) starts at byte p (a reverse-strand
// window starts rc_shift = k - |prefix| bytes before p: complement() reverses, so the prefix of the reverse strand is the
// END of the window).  Both strands look at the same shifted words.  live = 0x55555555 with KJ_F_FORWARD_ONLY.
// (Doing the shifts as multiplications on the FMA pipe, which is half idle, was measured slower: 1.05 vs 0.84 ms.)
template <int MP>
__device__ __forceinline__ uint32_t kj_chunk_lanes(const KjScanArgs &a, uint32_t c0, uint32_t c1, uint32_t live) {
    uint32_t accf = 0, accr = 0;
#pragma unroll
    for (int i = 0; i < MP; ++i) {
        const uint32_t sh = kj_funnel_r(c0, c1, 2u * i);
        accf |= sh ^ a.pat_f[i];
        accr |= sh ^ a.pat_r[i];
    }
    const uint32_t nf = accf | (accf >> 1);               // bit 0 of a lane: some filter symbol differs from the prefix
    const uint32_t nr = accr | (accr << 1);               // bit 1 of a lane: the same for complement(prefix)
    const uint32_t any = (nf & 0x55555555u) | (nr & 0xAAAAAAAAu);
    return ~any & live;
}

// append {z, loc} to the LANE's queue when z != 0: a predicated store and a predicated add, no atomic, no branch (a
// divergent region costs a compare, a convergence barrier, a branch and a wait for every chunk whether it holds a
// candidate or not; a shared-memory atomic with a result cannot be predicated at all).  qaddr: where the lane's next
// entry goes (shared-memory address on the device, a pointer under tools/cuemu).
#if defined(__CUDA_ARCH__)
typedef uint32_t kj_qaddr;
__device__ __forceinline__ void kj_wt_push(kj_qaddr &qaddr, uint32_t z, uint32_t loc) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.u32 p, %1, 0;\n\t"
                 "@p st.shared.v2.u32 [%0], {%1, %2};\n\t"
                 "@p add.u32 %0, %0, 256;\n\t}"
                 : "+r"(qaddr) : "r"(z), "r"(loc) : "memory");
}
__device__ __forceinline__ kj_qaddr kj_qaddr_of(KjWarpSmem &ws, uint32_t lane) { return kj_smem_u32(&ws.pq[0][lane]); }
__device__ __forceinline__ uint32_t kj_q_count(KjWarpSmem &ws, uint32_t lane, kj_qaddr qaddr) {
    return (qaddr - kj_smem_u32(&ws.pq[0][lane])) >> 8;
}
#else
typedef uint2 *kj_qaddr;
__device__ __forceinline__ void kj_wt_push(kj_qaddr &qaddr, uint32_t z, uint32_t loc) {
    if (z) { *qaddr = make_uint2(z, loc); qaddr += 32; }
}
__device__ __forceinline__ kj_qaddr kj_qaddr_of(KjWarpSmem &ws, uint32_t lane) { return &ws.pq[0][lane]; }
__device__ __forceinline__ uint32_t kj_q_count(KjWarpSmem &ws, uint32_t lane, kj_qaddr qaddr) {
    return (uint32_t)(qaddr - &ws.pq[0][lane]) / 32u;
}
#endif

// per-warp state that lives in registers across tiles
struct KjWarpRegs {
    kj_qaddr qaddr;                    // per lane: where its next queue entry goes
    unsigned long long blk_at;         // entry slots of the block in use: [blk_at, blk_at + blk_left)   (uniform over the lanes)
    uint32_t blk_left;
    unsigned long long next_at;        // the block reserved ahead (its atomic was issued a drain ago: nothing waits for it)
};

// What follows the conversion of a tile, for a lane with its code words (cw[8]: the one behind its last chunk) and its
// packed newline masks: newline bookkeeping, the search, the queue.  MASKED: edge tiles clip the candidates to the positions
// whose windows start inside the owned range.
template <int MP, bool MASKED>
__device__ __forceinline__ void kj_wt_finish_tile(const KjScanArgs &a, KjWarpSmem &ws, const uint32_t (&cw)[9],
                                                  const uint32_t (&nlp)[4], uint64_t t, uint32_t slot, uint32_t lane,
                                                  uint32_t live, kj_qaddr &qaddr) {
    // the lane's masks, '\n' before the lane inside the tile, the tile's count
    ws.bitmap[slot][lane] = make_uint4(nlp[0], nlp[1], nlp[2], nlp[3]);
    const uint32_t cnt = __popc(nlp[0]) + __popc(nlp[1]) + __popc(nlp[2]) + __popc(nlp[3]);
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    ws.lanepre[slot][lane] = incl - cnt;
    if (lane == KJ_WT_OWN_ROWS - 1u) a.tile_cnt[t] = incl;     // the 31 owned rows
    const uint32_t locbase = slot << 3;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        uint32_t z = kj_chunk_lanes<MP>(a, cw[i], cw[i + 1], live);
        if (MASKED) {
            // forward candidates below own_n; reverse ones below own_n + rc_shift (their windows start rc_shift earlier)
            const uint64_t o = t * KJ_WT_BYTES + lane * 128u + (uint32_t)i * 16u;
            const uint64_t lim_f = a.own_n, lim_r = a.own_n + a.rc_shift;
            uint32_t keep = 0;
            if (o + 16 <= lim_f) keep |= 0x55555555u;
            else if (o < lim_f) keep |= ((1u << (2u * (uint32_t)(lim_f - o))) - 1u) & 0x55555555u;
            if (o + 16 <= lim_r) keep |= 0xAAAAAAAAu;
            else if (o < lim_r) keep |= ((1u << (2u * (uint32_t)(lim_r - o))) - 1u) & 0xAAAAAAAAu;
            z &= keep;
        }
        kj_wt_push(qaddr, z, locbase | (uint32_t)i);
    }
    __syncwarp();
}

// ----------------------------------------------------------------------------- drain

// one thread: reserve the next block of entry slots
__device__ __forceinline__ unsigned long long kj_wt_reserve(const KjScanArgs &a) {
    return atomicAdd(&a.ctr->n_cand, (unsigned long long)KJ_WT_BLOCK);
}
// the unused slots [at, at + n) of a block become empty entries (no candidate lanes)
__device__ __forceinline__ void kj_wt_blank(const KjScanArgs &a, unsigned long long at, uint32_t n, uint32_t lane) {
    for (uint32_t i = lane; i < n; i += 32)
        if (at + i < a.cand_cap) reinterpret_cast<uint4 *>(a.cand)[at + i] = make_uint4(0, 0, 0, 0);
}

// The whole warp turns the lanes' queues into entries in global memory: dense again, one entry per lane and round (entry e of
// the concatenated queues belongs to the lane whose inclusive count is the first above e: five shuffle probes find it).
// Entry slots come from blocks of KJ_WT_BLOCK the warp reserves one drain ahead, so that nothing waits for the atomic.
static __device__ __noinline__ void kj_wt_drain(const KjScanArgs &a, KjWarpSmem &ws, uint32_t t_cur, uint32_t slot_cur,
                                                uint32_t G, KjWarpRegs &wr) {
    const uint32_t lane = threadIdx.x & 31;
    __syncwarp();
    const uint32_t mine = kj_q_count(ws, lane, wr.qaddr);
    uint32_t incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if ((int)lane >= d) incl += o;
    }
    const uint32_t n = __shfl_sync(0xFFFFFFFFu, incl, 31);
    for (uint32_t base = 0; base < n; base += 32) {
        const uint32_t e = base + lane;
        uint32_t src = 0;
#pragma unroll
        for (uint32_t st = 16; st > 0; st >>= 1) {
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (src + st - 1u) & 31u);
            if (v <= e) src += st;
        }
        src &= 31u;
        const uint32_t first = __shfl_sync(0xFFFFFFFFu, incl - mine, src);     // entries of the lanes before src
        bool keep = false;
        uint4 rec = make_uint4(0, 0, 0, 0);
        if (e < n) {
            const uint2 qe = ws.pq[e - first][src];
            const uint32_t z = qe.x, loc = qe.y;
            const uint32_t slot = loc >> 3, i = loc & 7u;
            if (src != 31u) {                                       // lane 31 converts the row behind the tile: not owned
                keep = true;
                const uint32_t age = (slot_cur - slot) & (KJ_WT_RING - 1u);
                const uint32_t tile = t_cur - age * G;
                const uint4 bm = ws.bitmap[slot][src];
                const uint32_t w[4] = {bm.x, bm.y, bm.z, bm.w};
                // bits of the lane's 128 below chunk i
                uint32_t before = 0, hi_w = 0, hi_j = 0;
#pragma unroll
                for (uint32_t j = 0; j < 4; ++j) {
                    const uint32_t lim = i * 16u;
                    const uint32_t m = (j * 32u + 32u <= lim) ? 0xFFFFFFFFu : ((j * 32u < lim) ? 0xFFFFu : 0u);
                    const uint32_t x = w[j] & m;
                    before += __popc(x);
                    if (x) { hi_w = x; hi_j = j; }
                }
                uint32_t dist = KJ_ENT_NODIST;
                if (hi_w) {
                    dist = i * 16u - (hi_j * 32u + (31u - __clz(hi_w)) + 1u);
                } else {
                    for (int l = (int)src - 1; l >= 0; --l) {       // the lanes before it, nearest first
                        const uint4 b2 = ws.bitmap[slot][l];
                        const uint32_t v[4] = {b2.x, b2.y, b2.z, b2.w};
                        int jj = -1;
                        for (int j = 3; j >= 0; --j) if (v[j]) { jj = j; break; }
                        if (jj >= 0) {
                            dist = (src - (uint32_t)l) * 128u + i * 16u - ((uint32_t)jj * 32u + (31u - __clz(v[jj])) + 1u);
                            break;
                        }
                    }
                }
                const uint32_t nlmask = (w[i >> 1] >> ((i & 1u) * 16u)) & 0xFFFFu;
                const uint64_t chunk = (uint64_t)tile * KJ_WT_CHUNKS + src * 8u + i;
                const uint64_t word = chunk | ((uint64_t)(ws.lanepre[slot][src] + before) << 40);
                rec = make_uint4((uint32_t)word, (uint32_t)(word >> 32), z, nlmask | (dist << 16));
            }
        }
        const uint32_t kb = __ballot_sync(0xFFFFFFFFu, keep);
        const uint32_t need = __popc(kb);
        if (need > wr.blk_left) {
            // the block is used up: blank its tail, go on in the one reserved ahead, reserve the one after
            kj_wt_blank(a, wr.blk_at, wr.blk_left, lane);
            wr.blk_at = wr.next_at;
            wr.blk_left = KJ_WT_BLOCK;
            unsigned long long nx = 0;
            if (lane == 0) nx = kj_wt_reserve(a);
            wr.next_at = __shfl_sync(0xFFFFFFFFu, nx, 0);
        }
        if (keep) {
            const unsigned long long at = wr.blk_at + __popc(kb & ((1u << lane) - 1u));
            if (at < a.cand_cap) reinterpret_cast<uint4 *>(a.cand)[at] = rec;    // beyond the buffer: the host sees n_cand and repeats the piece
        }
        wr.blk_at += need;
        wr.blk_left -= need;
    }
    __syncwarp();
    wr.qaddr = kj_qaddr_of(ws, lane);
}

// ----------------------------------------------------------------------------- scan kernel

// edge tiles (the last ones of a launch): bounds-checked loads straight from global memory, newlines clipped to the
// owned range.  Out of line: it runs once or twice per launch and must not sit in the hot loop.
template <int MP>
static __device__ __noinline__ void kj_wt_edge_tile(const KjScanArgs &a, KjWarpSmem &ws, uint64_t t, uint32_t slot, uint32_t live,
                                                    kj_qaddr &qaddr) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t cw[9], nlp[4] = {0, 0, 0, 0};
    const uint64_t lo = t * KJ_WT_BYTES + lane * 128u;
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const uint64_t o = lo + (uint32_t)i * 16u;
        const uint4 v = (o < a.n) ? kj_load_chunk(a.buf, o, a.n) : make_uint4(0, 0, 0, 0);
        cw[i] = kj_pack16(v.x, v.y, v.z, v.w);
        if (i < 8) {
            uint32_t m = kj_nl16(v.x, v.y, v.z, v.w);
            if (o >= a.own_n) m = 0;
            else if (o + 16 > a.own_n) m &= (1u << (uint32_t)(a.own_n - o)) - 1u;
            nlp[i >> 1] |= m << (16 * (i & 1));
        }
    }
    kj_wt_finish_tile<MP, true>(a, ws, cw, nlp, t, slot, lane, live, qaddr);
}

template <int MP>
__global__ void __launch_bounds__(KJ_WT_THREADS, 2)
kj_warp_filter_kernel(const __grid_constant__ KjTensorMap tmap, const __grid_constant__ KjScanArgs a) {
    KJ_DYN_SMEM(dyn);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [stages of all warps, 1024-byte aligned: the swizzle is a function of the shared-memory address][barriers][per-warp state]
    uint8_t *base = dyn + ((1024u - (uint32_t)(reinterpret_cast<uintptr_t>(dyn) & 1023u)) & 1023u);
    uint8_t *stage = base + (size_t)warp * KJ_WT_STAGES * KJ_WT_STAGE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(base + (size_t)KJ_WT_WARPS * KJ_WT_STAGES * KJ_WT_STAGE_BYTES) + warp * KJ_WT_STAGES;
    KjWarpSmem &ws = *(reinterpret_cast<KjWarpSmem *>(base + (size_t)KJ_WT_WARPS * KJ_WT_STAGES * (KJ_WT_STAGE_BYTES + 8)) + warp);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < KJ_WT_STAGES; ++s) kj_bar_init(&bars[s], 1);
#if defined(__CUDA_ARCH__)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
    }
    __syncwarp();
    const uint32_t G = gridDim.x * KJ_WT_WARPS, g = blockIdx.x * KJ_WT_WARPS + warp;
    if (g >= a.n_tiles) return;                          // a warp without a tile reserves nothing
    KjWarpRegs wr;
    wr.qaddr = kj_qaddr_of(ws, lane);
    {
        unsigned long long b0 = 0, b1 = 0;
        if (lane == 0) { b0 = kj_wt_reserve(a); b1 = kj_wt_reserve(a); }
        wr.blk_at = __shfl_sync(0xFFFFFFFFu, b0, 0);
        wr.next_at = __shfl_sync(0xFFFFFFFFu, b1, 0);
        wr.blk_left = KJ_WT_BLOCK;
    }
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < KJ_WT_STAGES; ++s) {
            const uint64_t t = (uint64_t)g + (uint64_t)s * G;
            if (t < a.n_fast) kj_tma_tile(stage + s * KJ_WT_STAGE_BYTES, &tmap, (uint32_t)t * KJ_WT_OWN_ROWS, &bars[s]);
        }
    }
    // a lane reads chunk i of its row at (i ^ (row & 7)): what the 128-byte swizzle made of it
    kj_saddr off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) off[i] = kj_saddr_of(stage) + lane * 128u + (((uint32_t)i << 4) ^ ((lane & 7u) << 4));
    const uint32_t c0a = a.c0a, c7f = a.c7f;          // kernel arguments, so that they stay register operands (kj_nl_flags_r)
    const uint32_t live = a.n_strands > 1 ? 0xFFFFFFFFu : 0x55555555u;
    uint32_t it = 0;
    uint32_t t = g;
    // ---- whole tiles through the TMA ring.  One body for every stage (the stage offset is an add per load): unrolled over the
    // stages the loop outgrows the instruction cache, which cost more than the adds.
    for (; t < a.n_fast; t += G, ++it) {
        const uint32_t s = it % KJ_WT_STAGES, slot = it & (KJ_WT_RING - 1u);
        // room in every lane's queue for everything this tile can add to it (8 entries)
        if (__any_sync(0xFFFFFFFFu, kj_q_count(ws, lane, wr.qaddr) + 8u > KJ_WT_PQCAP))
            kj_wt_drain(a, ws, t - G, (slot - 1u) & (KJ_WT_RING - 1u), G, wr);
        kj_bar_wait(&bars[s], (it / KJ_WT_STAGES) & 1u);
        const uint32_t soff = s * KJ_WT_STAGE_BYTES;
        uint32_t cw[9], nlp[4];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const uint4 v0 = kj_lds128(off[i] + soff), v1 = kj_lds128(off[i + 1] + soff);
            cw[i] = kj_pack16(v0.x, v0.y, v0.z, v0.w);
            cw[i + 1] = kj_pack16(v1.x, v1.y, v1.z, v1.w);
            nlp[i >> 1] = kj_nl16_r(v0, c0a, c7f) | (kj_nl16_r(v1, c0a, c7f) << 16);
        }
        __syncwarp();                                        // every lane has read the stage
        if (lane == 0) {
            const uint64_t tn = (uint64_t)t + (uint64_t)KJ_WT_STAGES * G;
            if (tn < a.n_fast) kj_tma_tile(stage + soff, &tmap, (uint32_t)tn * KJ_WT_OWN_ROWS, &bars[s]);
        }
        cw[8] = __shfl_down_sync(0xFFFFFFFFu, cw[0], 1);
        kj_wt_finish_tile<MP, false>(a, ws, cw, nlp, t, slot, lane, live, wr.qaddr);
        // the ring keeps KJ_WT_RING tiles: drain when it is full
        if (slot == KJ_WT_RING - 1u) kj_wt_drain(a, ws, t, slot, G, wr);
    }
    // ---- edge tiles
    for (; t < a.n_tiles; t += G, ++it) {
        const uint32_t slot = it & (KJ_WT_RING - 1u);
        if (__any_sync(0xFFFFFFFFu, kj_q_count(ws, lane, wr.qaddr) + 8u > KJ_WT_PQCAP))
            kj_wt_drain(a, ws, t - G, (slot - 1u) & (KJ_WT_RING - 1u), G, wr);
        kj_wt_edge_tile<MP>(a, ws, t, slot, live, wr.qaddr);
        if (slot == KJ_WT_RING - 1u) kj_wt_drain(a, ws, t, slot, G, wr);
    }
    // what is left in the queues (the ring slot of the last tile is (it - 1) mod ring), then the unused entry slots
    kj_wt_drain(a, ws, t - G, (it - 1u) & (KJ_WT_RING - 1u), G, wr);
    kj_wt_blank(a, wr.blk_at, wr.blk_left, lane);
    kj_wt_blank(a, wr.next_at, KJ_WT_BLOCK, lane);
}

// ----------------------------------------------------------------------------- resolve kernel

// The window at buffer offset j: the KW + 1 aligned 4-byte words from j & ~3 on, out of the (at most) three aligned 16-byte
// chunks that hold buf[j, j + k) (kj_window_load).  KW = 4-byte words of the window the check looks at: 8 covers every
// k <= 32; 4 (k <= 16, the KmerFinder default) halves the work.  The word offset is taken in two binary steps (by 2, by 1).
template <int KW>
__device__ __forceinline__ void kj_window_words(uint64_t j, const uint4 v0, const uint4 v1, const uint4 v2, uint32_t (&w)[KW + 1]) {
    const uint32_t W[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
    const bool by2 = (j & 8u) != 0, by1 = (j & 4u) != 0;
    uint32_t A[KW + 2];
#pragma unroll
    for (int i = 0; i < KW + 2; ++i) A[i] = by2 ? W[i + 2 < 12 ? i + 2 : 11] : W[i];
#pragma unroll
    for (int i = 0; i <= KW; ++i) w[i] = by1 ? A[i + 1] : A[i];
}
// Exact check and key, straight-line SIMD-in-register code.  Returns 0: not an emission (prefix bytes differ, crosses the end
// of the line, ...); 1: key holds the 2k-bit key; 2: irregular (some byte is not A/C/G/T): the byte string is the key.
template <int KW>
__device__ __forceinline__ int kj_window_key(const KjScanArgs &a, uint64_t j, uint32_t strand, const uint32_t (&w)[KW + 1],
                                             uint64_t &key) {
    const uint32_t k = a.k;
    const uint32_t r8 = ((uint32_t)j & 3u) * 8u;
    uint32_t bad = 0, nl = 0, irr = 0, p_lo = 0, p_hi = 0;
#pragma unroll
    for (int i = 0; i < KW; ++i) {
        if (4u * i < k) {
            const uint32_t x = kj_funnel_r(w[i], w[i + 1], r8);
            const uint32_t bm = (4u * i + 4u <= k) ? 0xFFFFFFFFu : ((1u << (8u * (k - 4u * i))) - 1u);   // bytes of the window
            bad |= (x ^ a.want[strand][i]) & a.wmask[strand][i];
            nl |= kj_nl_msb4(x) & bm;
            irr |= kj_not_acgt4(x) & bm;
            const uint32_t c8 = kj_pack4(x & bm);
            if (i < 4) p_lo |= c8 << (8 * i); else p_hi |= c8 << (8 * (i - 4));
        }
    }
    if (nl | bad) return 0;                               // crosses the end of the line / prefix bytes differ
    if (k == 1 && a.line_gate) {                          // lib/kmers.js:151  line.length > 1: a lone byte is not processed
        const bool first = j == 0 ? (a.ctr->carry_last[a.parity] == a.voff) : (a.buf[j - 1] == '\n');
        const bool more = (j + 1 < a.n) && a.buf[j + 1] != '\n';
        if (first && !more) return 0;
    }
    if (irr) return 2;
    const uint64_t P = ((uint64_t)p_hi << 32) | p_lo;     // code of window byte i at bits 2i
    const uint64_t kmask = k == 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    // forward key: first base most significant; reverse key: complement codes, last base first
    key = strand ? ((P ^ 0xAAAAAAAAAAAAAAAAull) & kmask) : (kj_pairrev64(P) >> (64u - 2u * k));
    return 1;
}

__device__ __forceinline__ void kj_prefetch_l2(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// fields of an entry
__device__ __forceinline__ uint64_t kj_ent_chunk(const uint4 rec) { return ((uint64_t)(rec.y & 0xFFu) << 32) | rec.x; }
// the candidate lanes of an entry whose line index is 1 mod 4 (lib/kmers.js:151, i === 1); lines0 = lines before the tile
__device__ __forceinline__ uint32_t kj_ent_line_filter(const uint4 rec, uint64_t lines0) {
    if (!rec.z) return 0u;
    const uint32_t nlmask = rec.w & 0xFFFFu;
    const uint32_t l0 = (uint32_t)lines0 + ((rec.y >> 8) & 0x1FFFu);       // mod 4 is all that matters
    if (!nlmask) return (l0 & 3u) == 1u ? rec.z : 0u;                       // no '\n' inside the chunk: one line for all
    uint32_t keep = 0, lanes = rec.z;
    while (lanes) {
        const uint32_t bit = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        if (((l0 + __popc(nlmask & ((1u << (bit >> 1)) - 1u))) & 3u) == 1u) keep |= 1u << bit;
    }
    return keep;
}

// One entry whose candidate lanes passed the line filter (rec.z; lines0 = lines before its tile): for every candidate the
// window's bytes, the exact check, the key, the first-seen ordinal = (read, strand, column), the hash-table update.  A
// candidate that finds no slot within the probe limit stays in the entry (its bit of the lane mask is kept, the others are
// cleared) and the entry is marked for the retry pass.
template <int KW>
__device__ __forceinline__ void kj_resolve_entry(const KjScanArgs &a, const uint4 rec, uint64_t lines0, unsigned long long idx,
                                                 uint32_t &n_emit, uint32_t &n_fail) {
    uint32_t lanes = rec.z, failed = 0;
    const uint64_t chunk = kj_ent_chunk(rec);
    const uint32_t nlmask = rec.w & 0xFFFFu, dist = rec.w >> 16;
    const uint64_t tile = chunk / KJ_WT_CHUNKS;
    const uint64_t line0 = lines0 + ((rec.y >> 8) & 0x1FFFu);
    while (lanes) {
        const uint32_t bit = __ffs(lanes) - 1;
        lanes &= lanes - 1;
        const uint32_t p = bit >> 1, strand = bit & 1u;
        const uint64_t pos = chunk * 16u + p;                      // where the prefix / complement(prefix) starts
        const uint32_t back = strand ? a.rc_shift : 0u;            // the reverse-strand window starts k - m before
        if (!(pos >= back && pos - back < a.own_n && pos - back + a.k <= a.n)) continue;
        const uint64_t j = pos - back;
        uint4 v0, v1, v2;
        kj_window_load(a, j, v0, v1, v2);
        uint32_t w[KW + 1];
        kj_window_words<KW>(j, v0, v1, v2, w);
        uint64_t key = 0;
        const int st = kj_window_key<KW>(a, j, strand, w, key);    // a '\n' between window start and prefix fails here too
        if (!st) continue;
        uint64_t ord = 0;
        if (a.order || a.k == 1) {
            const uint32_t below = nlmask & ((1u << p) - 1u);
            const uint64_t line = line0 + __popc(below);
            unsigned long long start;                              // first byte of the line (virtual offset)
            if (below) start = a.voff + chunk * 16u + (31u - __clz(below)) + 1ull;
            else if (dist != KJ_ENT_NODIST) start = a.voff + chunk * 16u - dist;
            else start = kj_line_start_global(a, tile * KJ_WT_BYTES);
            const uint64_t col = a.voff + j - start;
            const uint64_t read_idx = line >> 2;
            if (col > KJ_POS_MAX) { atomicOr(&a.ctr->error_flags, KJ_DEV_E_LINE_TOO_LONG); continue; }
            if (read_idx >> 36) { atomicOr(&a.ctr->error_flags, KJ_DEV_E_READS_OVERFLOW); continue; }
            // forward emissions in ascending column, then reverse emissions in descending column
            ord = kj_ordinal(read_idx, strand, strand ? KJ_POS_MAX - col : col);
        }
        bool ok;
        if (st == 1) {
            ok = kj_insert<false>(a.tab, a.ctr, key, ord, 1);
        } else {
            __align__(8) uint8_t key32[32];
            kj_window_bytes(a.buf, j, a.k, strand, key32);
            ok = kj_insert_irr(a.irr, a.ctr, key32, a.k, ord, 1);
        }
        if (ok) ++n_emit;
        else { ++n_fail; failed |= 1u << bit; }
    }
    uint4 *ent = reinterpret_cast<uint4 *>(a.cand);
    const bool marked = (rec.y & KJ_ENT_RETRY32) != 0;
    if (failed) {
        ent[idx].z = failed;
        if (!marked) ent[idx].y = rec.y | KJ_ENT_RETRY32;
    } else if (marked) {
        ent[idx].z = 0u;
        ent[idx].y = rec.y & ~KJ_ENT_RETRY32;
    }
}

#ifndef KJ_RS_LAG
#define KJ_RS_LAG 0u               // entries a warp keeps waiting beyond a full batch (their window prefetches get a round more)
#endif
#ifndef KJ_RS_MINB
#define KJ_RS_MINB 4
#endif
#ifndef KJ_RS_PER
#define KJ_RS_PER 1                // entries per lane and round of the filter stage
#endif
#define KJ_RS_QUEUE 128u           // ring of entries per warp: at most 31 + KJ_RS_LAG left over + 32 new ones
struct KjResolveSmem {
    uint4 rec[8][KJ_RS_QUEUE];
    uint64_t lines0[8][KJ_RS_QUEUE];
    unsigned long long idx[8][KJ_RS_QUEUE];
};

// Entries -> table.  Every round a warp reads 32 entries (a chunk with candidate positions each), looks up the lines before
// their tiles and keeps the candidates whose line index (lines before the launch + lines before the tile + '\n' before the
// position inside the tile) is 1 mod 4 (lib/kmers.js:151, i === 1): that drops every candidate of a header or quality line,
// more than half of them, before anything is fetched.  The entries that are left go to a ring in shared memory and their
// first window is prefetched into L2; whenever the ring holds 32, every lane takes one (kj_resolve_entry): the expensive part
// runs with full warps and finds its window in L2.
// a.resolve_retry == 0: every entry; != 0: the entries an earlier pass marked (the host has grown the table in between);
// nothing is ever dropped, whatever the input looks like.
// Block 0 also closes the stream state of the launch (lines and last '\n' so far) for the next one.
template <int KW>
__global__ void __launch_bounds__(256, KJ_RS_MINB) kj_resolve_kernel(const __grid_constant__ KjScanArgs a) {
    __shared__ KjResolveSmem sm;
    const unsigned long long n_ent = a.ctr->n_cand < a.cand_cap ? a.ctr->n_cand : a.cand_cap;
    const uint64_t base_lines = a.ctr->carry_lines[a.parity];
    if (blockIdx.x == 0 && threadIdx.x == 0 && !a.resolve_retry && a.n_tiles) {
        a.ctr->carry_lines[a.parity ^ 1] = base_lines + a.tile_excl[a.n_tiles - 1] + a.tile_cnt[a.n_tiles - 1];
        a.ctr->carry_last[a.parity ^ 1] = kj_line_start_global(a, a.own_n);
    }
    if (a.ctr->n_cand > a.cand_cap) return;                   // the entry buffer was too small: nothing is touched, the host repeats the piece
    uint32_t n_emit = 0, n_fail = 0;
    const uint4 *ent = reinterpret_cast<const uint4 *>(a.cand);
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    // KJ_RS_PER entries per lane and round: their loads (entry, then lines before the tile) are in flight together, and the
    // fixed cost of a round is shared
    const unsigned long long per_round = stride * KJ_RS_PER;
    const unsigned long long rounds = (n_ent + per_round - 1) / per_round;    // the same trip count for every thread (warp collectives inside)
    unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const uint4 none = make_uint4(0, 0, 0, 0);
    uint32_t q_head = 0, q_n = 0;
    uint4 rec_next[KJ_RS_PER];
#pragma unroll
    for (int j = 0; j < KJ_RS_PER; ++j) rec_next[j] = i + j * stride < n_ent ? ent[i + j * stride] : none;
    for (unsigned long long r = 0; r < rounds; ++r, i += per_round) {
        uint4 rec[KJ_RS_PER];
        uint64_t lines0[KJ_RS_PER];
#pragma unroll
        for (int j = 0; j < KJ_RS_PER; ++j) {
            rec[j] = rec_next[j];
            const unsigned long long nx = i + per_round + j * stride;
            rec_next[j] = nx < n_ent ? ent[nx] : none;                    // in flight while this round works
            if (a.resolve_retry && !(rec[j].y & KJ_ENT_RETRY32)) rec[j].z = 0;
        }
#pragma unroll
        for (int j = 0; j < KJ_RS_PER; ++j)
            lines0[j] = rec[j].z ? base_lines + a.tile_excl[kj_ent_chunk(rec[j]) / KJ_WT_CHUNKS] : 0;
#pragma unroll
        for (int j = 0; j < KJ_RS_PER; ++j) {
            rec[j].z = kj_ent_line_filter(rec[j], lines0[j]);
            const uint32_t m = __ballot_sync(0xFFFFFFFFu, rec[j].z != 0u);
            if (rec[j].z) {
                const uint32_t at = (q_head + q_n + __popc(m & ((1u << lane) - 1u))) & (KJ_RS_QUEUE - 1u);
                sm.rec[warp][at] = rec[j];
                sm.lines0[warp][at] = lines0[j];
                sm.idx[warp][at] = i + j * stride;
                const uint32_t bit = __ffs(rec[j].z) - 1;
                const uint64_t pos = kj_ent_chunk(rec[j]) * 16u + (bit >> 1);
                const uint64_t jw = (bit & 1u) ? (pos >= a.rc_shift ? pos - a.rc_shift : 0) : pos;
                if (jw < a.n) {
                    kj_prefetch_l2(a.buf + (jw & ~31ull));
                    if ((jw & 31u) + a.k > 32u && (jw | 31ull) + 1 < a.n) kj_prefetch_l2(a.buf + (jw | 31ull) + 1);
                }
            }
            q_n += __popc(m);
            __syncwarp();
            if (q_n >= 32u + KJ_RS_LAG) {
                const uint32_t at = (q_head + lane) & (KJ_RS_QUEUE - 1u);
                kj_resolve_entry<KW>(a, sm.rec[warp][at], sm.lines0[warp][at], sm.idx[warp][at], n_emit, n_fail);
                q_head = (q_head + 32u) & (KJ_RS_QUEUE - 1u);
                q_n -= 32u;
                __syncwarp();
            }
        }
    }
    for (; q_n; q_n -= min(q_n, 32u), q_head = (q_head + 32u) & (KJ_RS_QUEUE - 1u)) {
        if (lane < q_n) {
            const uint32_t at = (q_head + lane) & (KJ_RS_QUEUE - 1u);
            kj_resolve_entry<KW>(a, sm.rec[warp][at], sm.lines0[warp][at], sm.idx[warp][at], n_emit, n_fail);
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        n_emit += __shfl_xor_sync(0xFFFFFFFFu, n_emit, d);
        n_fail += __shfl_xor_sync(0xFFFFFFFFu, n_fail, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n_emit) atomicAdd(&a.ctr->n_occ, (unsigned long long)n_emit);
        if (n_fail) atomicAdd(&a.ctr->n_overflow, (unsigned long long)n_fail);
    }
}

// ----------------------------------------------------------------------------- KJ_F_COUNT_BASES

// Sum of the lengths of the sequence lines (index 1 mod 4): a '\n' that ends such a line adds its offset, one
// that ends the line before it subtracts offset + 1; the sums telescope over tiles and launches and the host
// closes the two ends of the stream (kj_counts_finish).  A second pass over the input, one warp per tile: the
// statistic is not part of the reference and off by default.
__global__ void __launch_bounds__(256) kj_bases_kernel(const __grid_constant__ KjScanArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t base_lines = a.ctr->carry_lines[a.parity];
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    long long sum = 0;
    for (uint64_t t = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < a.n_tiles; t += warps) {
        uint32_t m[8], cnt = 0;
        const uint64_t lo = t * KJ_WT_BYTES + lane * 128u;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint64_t o = lo + (uint32_t)i * 16u;
            m[i] = 0;
            if (lane < KJ_WT_OWN_ROWS && o < a.own_n) {
                const uint4 v = kj_load_chunk(a.buf, o, a.n);
                m[i] = kj_nl16(v.x, v.y, v.z, v.w);
                if (o + 16 > a.own_n) m[i] &= (1u << (uint32_t)(a.own_n - o)) - 1u;
            }
            cnt += __popc(m[i]);
        }
        uint32_t incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        uint64_t line = base_lines + a.tile_excl[t] + (incl - cnt);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            uint32_t mk = m[i];
            const long long at = (long long)(a.voff + lo + (uint32_t)i * 16u);
            while (mk) {
                const uint32_t bit = __ffs(mk) - 1; mk &= mk - 1;
                const uint32_t ph = (uint32_t)line & 3u;
                if (ph == 1u) sum += at + bit;
                else if (ph == 0u) sum -= at + bit + 1;
                ++line;
            }
        }
    }
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
    if (lane == 0 && sum) atomicAdd(&a.ctr->n_bases, (unsigned long long)sum);
}
