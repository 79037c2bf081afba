// kj_score.cu -- template scoring on the GPU (K4-K6 of SURVEY.md appendix B):
//
//   kj_db_*         k-mer -> ordered template list as a device hash index + CSR lists; replaces the
//                   Redis LRANGE store of lib/kmerFinderServer.js:171-226 (schema :68-92,184-199)
//   kj_first_match  per-template uScore / tScore / hits over the query (lib/kmerFinderServer.js:180-201)
//   kj_wta_next     one round of the findMatches generator (lib/kmerFinderClient.js:174-290):
//                   argmax over uScore (ties: first-encounter order), gate, removal of the winner's
//                   k-mers with incremental score decrements
//
// Layout in HBM.  DB: keys[cap] u64 + vals[cap] u32 (open addressing, linear probing), list_off
// [n_kmers+1] u64, tmpl[pairs] u32 (DB list order), ulen[T] u64.  Match: score vector
// S = {u[T], tau[T], H} u64 (partial sums of this rank) and G (global sums; G == S on one GPU),
// first-encounter keys first_ord[T], first_idx[T] u64, the per-template CSR of
// matched query entries toff[T+1] u64 / tq[hits] u32, qkmer[Q] u32 (DB k-mer id per query entry).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>
#include "kj_internal.hpp"
#ifndef KJ_CPU_EMU
#include <cub/device/device_scan.cuh>
#endif
#include "kj_stats.hpp"

#define KJ_SCORE_THREADS 256
#define KJ_SMEM_T_MAX 8192u     // templates whose u (u32) and tau (u64) are privatised in shared memory

struct KjWtaResult {        // written by the argmax / loop kernels, copied to the host
    uint32_t winner;        // template id, KJ_NONE32 when every uScore is zero
    uint32_t pad;
    uint64_t u, tau, hits;
    double z, p;            // double-precision zScore / fastp * templates (the device gate)
    uint64_t u0, tau0;      // the winner's first-round scores (loop kernel only)
};

struct kj_match {
    kj_ctx *ctx = nullptr;
    kj_counts *q = nullptr;
    const kj_db *db = nullptr;
    uint32_t T = 0;
    uint64_t Q = 0;
    // the query entries and the template lists the kernels read: the counts handle's and the database's
    // arrays, or (kj_match_from_matched) copies of the matched entries of every rank, owned by the match
    const uint64_t *qcount = nullptr, *qord = nullptr;
    uint8_t *alive = nullptr;
    KjDbDev d{};
    uint64_t *own_count = nullptr, *own_ord = nullptr, *own_off = nullptr;
    uint8_t *own_alive = nullptr;
    uint32_t *own_tmpl = nullptr;
    uint32_t *d_qkmer = nullptr;
    unsigned long long *d_msize = nullptr;   // {matched entries, template-list pairs} of kj_match_matched_size
    uint64_t *d_part = nullptr;      // {u[T], tau[T], H}: this rank's sums
    uint64_t *d_glob = nullptr;      // global sums (== d_part unless the host layer reduces over ranks)
    uint64_t *d_first_ord = nullptr, *d_first_idx = nullptr;
    uint64_t *d_toff = nullptr;
    unsigned long long *d_tcur = nullptr;
    uint32_t *d_tq = nullptr;
    KjWtaResult *d_res = nullptr, *h_res = nullptr;
    void *d_loop = nullptr;          // kj_wta_loop_kernel: {records, status} + one KjWtaResult per round
    bool committed = false;
    bool distributed = false;        // d_glob is separate and maintained by the host layer
    bool from_segments = false;      // kj_match_from_segments: sizes / query size / flags arrive with the commit
    uint64_t *d_glob0 = nullptr;     // {u[T], tau[T]} as the first match left them (lib/kmerFinderClient.js:44-46)
    unsigned int *d_sync = nullptr;  // kj_wta_loop_kernel: grid barrier counter + one control word per round
    bool host_first = false;         // u0 / t0 / order / toff_h below have been fetched (on demand: the hot path never needs them)
    std::vector<uint64_t> u0, t0;    // first-round scores
    std::vector<uint32_t> order;     // matched templates in first-encounter order
    std::vector<uint64_t> toff_h;    // host copy of toff (range of a winner's matched entries)
    uint64_t hits0 = 0;
    uint64_t kmer_map_size = 0;
    uint64_t seg_entries = 0, seg_pairs = 0;   // kj_match_from_segments: what the ranks really sent
    uint64_t pair_bound = 0;         // upper bound of the matched template-list entries, known without asking the device (0: unknown)
    uint32_t max_hits = 100, hit_counter = 0;
    bool ended = false;
    bool inflight = false;           // the argmax of the next round has been launched ahead
    bool defer_rows = false;         // kj_wta_next may return 2: the exact row is finished by kj_wta_row
    bool row_pending = false;
    KjWtaResult pending{};           // integers of the row to finish
};

// ------------------------------------------------------------------------------------ kernels

__device__ __forceinline__ uint32_t kj_db_lookup(const KjDbDev &d, uint64_t key) {
    uint64_t slot = kj_mix64(key) & d.mask;
    for (;;) {
        uint64_t cur = d.keys[slot];
        if (cur == key) return d.vals[slot];
        if (cur == KJ_EMPTY) return KJ_NONE32;
        slot = (slot + 1) & d.mask;
    }
}

__global__ void kj_db_build_kernel(uint64_t *keys, uint32_t *vals, uint64_t mask, const uint64_t *in_keys,
                                   const uint32_t *in_ids, uint64_t n, unsigned int *dup_flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t key = in_keys[i];
        uint64_t slot = kj_mix64(key) & mask;
        for (;;) {
            uint64_t cur = atomicCAS((unsigned long long *)&keys[slot], (unsigned long long)KJ_EMPTY,
                                     (unsigned long long)key);
            if (cur == KJ_EMPTY) { vals[slot] = in_ids[i]; break; }
            if (cur == key) { atomicOr(dup_flag, 1u); break; }     // duplicate k-mer in the DB
            slot = (slot + 1) & mask;
        }
    }
}

// query entry -> DB k-mer id (table entries only; the special and irregular ones are resolved on the host)
__global__ void kj_probe_kernel(KjDbDev d, const uint64_t *qkeys, uint64_t n_tab, uint32_t *qkmer) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_tab;
         i += (uint64_t)gridDim.x * blockDim.x)
        qkmer[i] = kj_db_lookup(d, qkeys[i]);
}

struct KjWalkArgs {
    KjDbDev d;
    const uint32_t *qkmer;
    const uint64_t *qcount, *qord;
    const uint8_t *alive;
    uint64_t Q;
    uint32_t T;
    uint64_t *part;              // u[T], tau[T], H
    uint64_t *first_ord, *first_idx;
    const uint64_t *toff;
    unsigned long long *tcur;
    uint32_t *tq;
};

enum { KJ_WALK_ACCUM = 0, KJ_WALK_FIRST = 1, KJ_WALK_FILL = 2, KJ_WALK_FIRST_FILL = 3 };   // 3: both in one pass

// Every warp takes 32 consecutive query entries; entries that hit the DB are walked one after the
// other by the whole warp (lanes stride over the template list: coalesced reads of tmpl[]).
//   ACCUM: u[t] += 1, tau[t] += count, first_ord[t] = min(ord), H += list length.  With SMEM the
//          block keeps u (u32) / tau (u64) in shared memory and flushes once.
//   FIRST: first_idx[t] = min(list index) over the entries whose ord equals first_ord[t].
//   FILL : tq[toff[t] + cursor[t]++] = q  (per-template list of matched query entries).
template <int MODE, bool SMEM>
__global__ void __launch_bounds__(KJ_SCORE_THREADS) kj_walk_kernel(const KjWalkArgs a) {
    KJ_DYN_SMEM(dyn);
    unsigned long long *s_tau = reinterpret_cast<unsigned long long *>(dyn);
    uint32_t *s_u = reinterpret_cast<uint32_t *>(dyn + (size_t)a.T * 8);
    const uint32_t lane = threadIdx.x & 31;
    if (MODE == KJ_WALK_ACCUM && SMEM) {
        for (uint32_t t = threadIdx.x; t < a.T; t += blockDim.x) { s_tau[t] = 0; s_u[t] = 0; }
        __syncthreads();
    }
    unsigned long long my_hits = 0;
    const uint64_t n_groups = (a.Q + 31) / 32;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const uint64_t q = g * 32 + lane;
        uint32_t id = KJ_NONE32;
        uint64_t cnt = 0, ord = 0;
        if (q < a.Q && a.alive[q]) {
            id = a.qkmer[q];
            if (id != KJ_NONE32) { cnt = a.qcount[q]; ord = a.qord[q]; }
        }
        // the 32 template lists as one sequence of (entry, position) items, 32 items per round: every lane has work whatever
        // the list lengths are, and no round waits for a list header (they were all requested at once)
        uint64_t lo = 0;
        uint32_t len = 0;
        if (id != KJ_NONE32) { lo = a.d.list_off[id]; len = (uint32_t)(a.d.list_off[id + 1] - lo); }
        uint32_t incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if ((int)lane >= d) incl += o;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (MODE == KJ_WALK_ACCUM && lane == 0) my_hits += total;
        for (uint32_t base = 0; base < total; base += 32) {
            const uint32_t item = base + lane;
            uint32_t src = 0;
#pragma unroll
            for (uint32_t st = 16; st > 0; st >>= 1) {
                const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (src + st - 1u) & 31u);
                if (v <= item) src += st;
            }
            src &= 31u;                                        // lanes beyond the total read lane 0's values and do nothing
            const uint32_t first = __shfl_sync(0xFFFFFFFFu, incl - len, src);
            const uint64_t slo = __shfl_sync(0xFFFFFFFFu, lo, src);
            const uint64_t c = __shfl_sync(0xFFFFFFFFu, cnt, src);
            const uint64_t o = __shfl_sync(0xFFFFFFFFu, ord, src);
            if (item < total) {
                const uint64_t i = slo + (item - first);
                const uint32_t t = a.d.tmpl[i];
                if (MODE == KJ_WALK_ACCUM) {
                    if (SMEM) {
                        atomicAdd(&s_u[t], 1u);
                        atomicAdd(&s_tau[t], (unsigned long long)c);
                    } else {
                        atomicAdd((unsigned long long *)&a.part[t], 1ull);
                        atomicAdd((unsigned long long *)&a.part[a.T + t], (unsigned long long)c);
                    }
                    if (a.first_ord[t] > o) atomicMin((unsigned long long *)&a.first_ord[t], (unsigned long long)o);
                } else {
                    if (MODE == KJ_WALK_FIRST || MODE == KJ_WALK_FIRST_FILL) {
                        if (a.first_ord[t] == o)
                            atomicMin((unsigned long long *)&a.first_idx[t], (unsigned long long)(item - first));
                    }
                    if (MODE == KJ_WALK_FILL || MODE == KJ_WALK_FIRST_FILL) {
                        unsigned long long pos = atomicAdd(&a.tcur[t], 1ull);
                        a.tq[a.toff[t] + pos] = (uint32_t)(g * 32 + src);
                    }
                }
            }
        }
    }
    if (MODE == KJ_WALK_ACCUM) {
        for (int d = 16; d > 0; d >>= 1) my_hits += __shfl_xor_sync(0xFFFFFFFFu, my_hits, d);
        if (lane == 0 && my_hits) atomicAdd((unsigned long long *)&a.part[2 * (uint64_t)a.T], my_hits);
        if (SMEM) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < a.T; t += blockDim.x) {
                uint32_t u = s_u[t];
                if (u) {
                    atomicAdd((unsigned long long *)&a.part[t], (unsigned long long)u);
                    atomicAdd((unsigned long long *)&a.part[a.T + t], s_tau[t]);
                }
            }
        }
    }
}

// Matched entries of this rank as self-contained records {count, ord, off, len} + their template lists
// (copied in DB order), for the all-gather that lets every rank run the winner-takes-all loop on the whole
// matched set without a collective per round.  COUNT only sizes the two outputs.
template <bool COUNT>
__global__ void __launch_bounds__(KJ_SCORE_THREADS)
kj_matched_export_kernel(KjDbDev d, const uint32_t *qkmer, const uint64_t *qcount, const uint64_t *qord,
                         const uint8_t *alive, uint64_t Q, uint64_t *entries, uint64_t cap_entries,
                         uint32_t *tmpl_out, uint64_t cap_pairs, unsigned long long *ctr) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t n_groups = (Q + 31) / 32;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t g = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
        const uint64_t q = g * 32 + lane;
        uint32_t id = KJ_NONE32;
        if (q < Q && alive[q]) id = qkmer[q];
        const bool hit = id != KJ_NONE32;
        uint64_t lo = 0, len = 0;
        if (hit) { lo = d.list_off[id]; len = d.list_off[id + 1] - lo; }
        uint32_t hitmask = __ballot_sync(0xFFFFFFFFu, hit);
        if (!hitmask) continue;
        uint64_t incl = len;
        for (int s = 1; s < 32; s <<= 1) {
            uint64_t o = __shfl_up_sync(0xFFFFFFFFu, incl, s);
            if ((int)lane >= s) incl += o;
        }
        const uint64_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long ebase = 0, pbase = 0;
        if (lane == 0) {
            ebase = atomicAdd(&ctr[0], (unsigned long long)__popc(hitmask));
            pbase = atomicAdd(&ctr[1], (unsigned long long)total);
        }
        if (COUNT) continue;
        ebase = __shfl_sync(0xFFFFFFFFu, ebase, 0);
        pbase = __shfl_sync(0xFFFFFFFFu, pbase, 0);
        const uint64_t my_off = pbase + incl - len;
        if (hit) {
            const uint64_t e = ebase + __popc(hitmask & ((1u << lane) - 1));
            if (e < cap_entries) {
                entries[4 * e + 0] = qcount[q];
                entries[4 * e + 1] = qord[q];
                entries[4 * e + 2] = my_off;
                entries[4 * e + 3] = len;
            }
        }
        while (hitmask) {
            const int src = __ffs(hitmask) - 1;
            hitmask &= hitmask - 1;
            const uint64_t slo = __shfl_sync(0xFFFFFFFFu, lo, src);
            const uint64_t slen = __shfl_sync(0xFFFFFFFFu, len, src);
            const uint64_t dst = __shfl_sync(0xFFFFFFFFu, my_off, src);
            for (uint64_t i = lane; i < slen; i += 32)
                if (dst + i < cap_pairs) tmpl_out[dst + i] = d.tmpl[slo + i];
        }
    }
}

// One gathered segment -> the arrays of a gathered match.  Entry e of the segment becomes query entry
// e0 + e with "k-mer id" 2 * (e0 + e): list_off[id] / list_off[id + 1] are its own {begin, end}.
__global__ void kj_matched_import_kernel(const uint64_t *entries, uint64_t n, uint64_t e0, uint64_t p0,
                                         uint64_t p_end, uint64_t *count, uint64_t *ord, uint8_t *alive,
                                         uint32_t *qkmer, uint64_t *off, unsigned int *bad) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t e = e0 + i;
        const uint64_t o = entries[4 * i + 2], l = entries[4 * i + 3];
        count[e] = entries[4 * i + 0];
        ord[e] = entries[4 * i + 1];
        alive[e] = 1;
        qkmer[e] = (uint32_t)(2 * e);
        off[2 * e] = p0 + o;
        off[2 * e + 1] = p0 + o + l;
        if (p0 + o + l > p_end || o + l < o) atomicOr(bad, 1u);
    }
}

// Fixed-capacity variant (no size round trip): a segment = {n_entries, n_pairs, query size, flags | entries[cap_e] x 32 B |
// template ids[cap_p]}; all segments of the all-gather are imported by one launch, sizes read on the device, and the
// template lists stay where the collective put them (list offsets index the gathered buffer as u32).
struct KjMSegHeader { unsigned long long n_entries, n_pairs, qsize, flags; };
__global__ void kj_matched_header_kernel(KjMSegHeader *h, unsigned long long qsize, unsigned long long flags) {
    h->n_entries = 0; h->n_pairs = 0; h->qsize = qsize; h->flags = flags;
}
__global__ void kj_matched_import_segments_kernel(const uint8_t *segs, uint64_t seg_bytes, uint32_t n_seg, uint32_t cap_e,
                                                  uint32_t cap_p, uint64_t *count, uint64_t *ord, uint8_t *alive,
                                                  uint32_t *qkmer, uint64_t *off, unsigned long long *info) {
    const uint64_t total = (uint64_t)n_seg * cap_e;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t sgm = (uint32_t)(i / cap_e), j = (uint32_t)(i % cap_e);
        const KjMSegHeader *h = reinterpret_cast<const KjMSegHeader *>(segs + (uint64_t)sgm * seg_bytes);
        const uint64_t n_s = h->n_entries < cap_e ? h->n_entries : cap_e;
        if (j >= n_s) continue;
        uint64_t e0 = 0;
        for (uint32_t q = 0; q < sgm; ++q) {
            const KjMSegHeader *hq = reinterpret_cast<const KjMSegHeader *>(segs + (uint64_t)q * seg_bytes);
            e0 += hq->n_entries < cap_e ? hq->n_entries : cap_e;
        }
        const uint64_t e = e0 + j;
        const uint64_t *ent = reinterpret_cast<const uint64_t *>(segs + (uint64_t)sgm * seg_bytes + sizeof(KjMSegHeader)) + 4 * (uint64_t)j;
        const uint64_t o = ent[2], l = ent[3];
        const uint64_t tbase = ((uint64_t)sgm * seg_bytes + sizeof(KjMSegHeader) + (uint64_t)cap_e * 32u) / 4u;
        count[e] = ent[0];
        ord[e] = ent[1];
        alive[e] = 1;
        qkmer[e] = (uint32_t)(2 * e);
        // a list that does not lie inside the segment (the segment overflowed) is flagged and left empty: the walk that is
        // queued right behind this kernel must not follow it; kj_match_commit then fails with KJ_E_RANGE
        const bool inside = o + l <= cap_p && o + l >= o;
        off[2 * e] = tbase + (inside ? o : 0);
        off[2 * e + 1] = tbase + (inside ? o + l : 0);
        if (!inside) atomicOr(&info[3], 2ull);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long ne = 0, np = 0, qs = 0, fl = 0;
        for (uint32_t q = 0; q < n_seg; ++q) {
            const KjMSegHeader *hq = reinterpret_cast<const KjMSegHeader *>(segs + (uint64_t)q * seg_bytes);
            ne += hq->n_entries; np += hq->n_pairs; qs += hq->qsize;
            if (hq->n_entries > cap_e || hq->n_pairs > cap_p) fl |= 2ull;
            if (hq->flags) fl |= 1ull;
        }
        info[0] = ne; info[1] = np; info[2] = qs;
        if (fl) atomicOr(&info[3], fl);
    }
}

// double-precision restatement of lib/stats.js:19-45 (the device gate; rows are finished exactly on the host)
__device__ __forceinline__ double kj_zscore_f64(double r1, double n1, double r2, double n2) {
    const double eta = 1.0e-8;
    double p1 = r1 / n1 + eta, p2 = r2 / n2 + eta;
    double p = (r1 + r2) / (n1 + n2 + eta);
    double q = 1.0 - p;
    double s = sqrt(p * q * (1.0 / (n1 + eta) + 1.0 / (n2 + eta)) + eta);
    return (p1 - p2) / s;
}
__device__ __forceinline__ double kj_fastp_f64(double z) {
    const double thr[27] = {10.7016, 10.4862, 10.2663, 10.0416, 9.81197, 9.5769, 9.33604, 9.08895, 8.83511,
                            8.57394, 8.30479, 8.02686, 7.73926, 7.4409,  7.13051, 6.8065,  6.46695, 6.10941,
                            5.73073, 5.32672, 4.89164, 4.41717, 3.89059, 3.29053, 2.57583, 1.95996, 1.64485};
    const double pv[27] = {1e-26, 1e-25, 1e-24, 1e-23, 1e-22, 1e-21, 1e-20, 1e-19, 1e-18,
                           1e-17, 1e-16, 1e-15, 1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9,
                           1e-8,  1e-7,  1e-6,  1e-5,  1e-4,  1e-3,  0.01,  0.05,  0.1};
    for (int i = 0; i < 27; ++i)
        if (z > thr[i]) return pv[i];
    return 1.0;
}

__global__ void kj_stats_kernel(uint64_t n, const uint64_t *r1, const uint64_t *n1, const uint64_t *r2,
                                const uint64_t *n2, double *z, double *p) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        double zz = kj_zscore_f64((double)r1[i], (double)n1[i], (double)r2[i], (double)n2[i]);
        z[i] = zz;
        p[i] = kj_fastp_f64(zz);
    }
}

// winner = argmax over (uScore desc, first-encounter order asc)  (lib/kmerFinderClient.js:100-109,181; the stable-sort tie
// rule of SURVEY.md 7.5).  First-encounter order = ascending (first ordinal, list index, id) (lib/kmerFinderServer.js:
// 180-201: query k-mers in Map order, each list in DB order); it is only ever needed between templates that tie on the
// score, so the block compares (score, order key) directly -- no rank of all templates against all templates is
// computed.  Every thread returns the winner (KJ_NONE32: none).
__device__ __forceinline__ bool kj_order_less(unsigned long long o, unsigned long long x, uint32_t t,
                                              unsigned long long fo, unsigned long long fi, uint32_t who) {
    return o < fo || (o == fo && (x < fi || (x == fi && t < who)));
}
__device__ __forceinline__ uint32_t kj_block_winner(const uint64_t *glob, const uint64_t *ford, const uint64_t *fidx, uint32_t T) {
    __shared__ unsigned long long s_u[32], s_a[32], s_b[32];
    __shared__ uint32_t s_w[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    unsigned long long bu = 0, fo = ~0ull, fi = ~0ull;
    uint32_t who = KJ_NONE32;
    // eight scores in flight per thread (the scores sit in L2: a thread that waited for each would spend the whole round here);
    // the order key of a template is only fetched when its score reaches the thread's best so far
    for (uint32_t t0 = threadIdx.x; t0 < T; t0 += blockDim.x * 8u) {
        unsigned long long u8[8];
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const uint32_t t = t0 + j * blockDim.x;
            u8[j] = t < T ? kj_ld_volatile(&glob[t]) : 0ull;
        }
#pragma unroll
        for (uint32_t j = 0; j < 8; ++j) {
            const uint32_t t = t0 + j * blockDim.x;
            if (u8[j] == 0ull || u8[j] < bu) continue;
            const unsigned long long o = ford[t], x = fidx[t];
            if (u8[j] > bu || kj_order_less(o, x, t, fo, fi, who)) { bu = u8[j]; fo = o; fi = x; who = t; }
        }
    }
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, bu, d);
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, fo, d), x = __shfl_xor_sync(0xFFFFFFFFu, fi, d);
        const uint32_t w = __shfl_xor_sync(0xFFFFFFFFu, who, d);
        if (u > bu || (u == bu && kj_order_less(o, x, w, fo, fi, who))) { bu = u; fo = o; fi = x; who = w; }
    }
    __syncthreads();                                   // the arrays may still be read by a previous call
    if (lane == 0) { s_u[warp] = bu; s_a[warp] = fo; s_b[warp] = fi; s_w[warp] = who; }
    __syncthreads();
    bu = s_u[0]; fo = s_a[0]; fi = s_b[0]; who = s_w[0];
    for (uint32_t i = 1; i < nw; ++i) {
        const unsigned long long u = s_u[i], o = s_a[i], x = s_b[i];
        const uint32_t w = s_w[i];
        if (u > bu || (u == bu && kj_order_less(o, x, w, fo, fi, who))) { bu = u; fo = o; fi = x; who = w; }
    }
    return bu ? who : KJ_NONE32;
}

// One block.
__global__ void __launch_bounds__(1024) kj_argmax_kernel(const uint64_t *glob, const uint64_t *ford, const uint64_t *fidx, uint32_t T,
                                                         const uint64_t *ulen, double unique_lens,
                                                         double n_templates, KjWtaResult *res) {
    const uint32_t who = kj_block_winner(glob, ford, fidx, T);
    {
        if (threadIdx.x == 0) {
            KjWtaResult r;
            r.winner = who; r.pad = 0;
            r.hits = glob[2 * (uint64_t)T];
            r.u = 0; r.tau = 0; r.z = 0.0; r.p = 1.0;
            if (who != KJ_NONE32) {
                r.u = glob[who];
                r.tau = glob[(uint64_t)T + who];
                r.z = kj_zscore_f64((double)r.u, (double)ulen[who], (double)r.hits, unique_lens);
                r.p = kj_fastp_f64(r.z) * n_templates;
            }
            *res = r;
        }
    }
}

// 32 matched entries [i0, i0 + 32) of a winner's list, one per lane: mark them dead and take their contribution out of
// every template that shares them.  The lanes fetch their entry's count and list header side by side, then the 32 template
// lists are walked as one sequence of (entry, position) items, 32 per round.  Returns (lane 0) the list entries removed.
__device__ __forceinline__ unsigned long long kj_remove_group(const KjDbDev &d, const uint32_t *tq, uint64_t i0, uint64_t hi,
                                                              const uint32_t *qkmer, const uint64_t *qcount, uint8_t *alive,
                                                              uint64_t *part, uint32_t T, uint32_t lane) {
    const uint64_t i = i0 + lane;
    uint64_t lo = 0;
    uint32_t len = 0;
    unsigned long long c = 0;
    if (i < hi) {
        const uint32_t q = tq[i];
        // each q occurs once in a template's list (DB lists are deduplicated): no other lane or warp looks at alive[q] now
        if (*reinterpret_cast<volatile uint8_t *>(&alive[q])) {
            alive[q] = 0;
            const uint32_t kid = qkmer[q];
            c = qcount[q];
            lo = d.list_off[kid];
            len = (uint32_t)(d.list_off[kid + 1] - lo);
        }
    }
    uint32_t incl = len;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
        const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, incl, s);
        if ((int)lane >= s) incl += o;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    for (uint32_t base = 0; base < total; base += 32) {
        const uint32_t item = base + lane;
        uint32_t src = 0;
#pragma unroll
        for (uint32_t st = 16; st > 0; st >>= 1) {
            const uint32_t v = __shfl_sync(0xFFFFFFFFu, incl, (src + st - 1u) & 31u);
            if (v <= item) src += st;
        }
        src &= 31u;
        const uint32_t first = __shfl_sync(0xFFFFFFFFu, incl - len, src);
        const uint64_t slo = __shfl_sync(0xFFFFFFFFu, lo, src);
        const unsigned long long sc = __shfl_sync(0xFFFFFFFFu, c, src);
        if (item < total) {
            const uint32_t t = d.tmpl[slo + (item - first)];
            atomicAdd((unsigned long long *)&part[t], ~0ull);                     // u[t] -= 1
            atomicAdd((unsigned long long *)&part[(uint64_t)T + t], 0ull - sc);   // tau[t] -= count
        }
    }
    return total;
}

// Remove the winner's k-mers from the query (kmerMap.delete, lib/kmerFinderClient.js:220-230) and
// take their contribution out of every template that shares them.  One warp per 32 matched entries.
__global__ void kj_remove_kernel(KjDbDev d, const uint32_t *tq, uint64_t lo, uint64_t hi, const uint32_t *qkmer,
                                 const uint64_t *qcount, uint8_t *alive, uint64_t *part, uint32_t T) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long gone = 0;
    for (uint64_t i = lo + 32 * (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); i < hi; i += 32 * warps)
        gone += kj_remove_group(d, tq, i, hi, qkmer, qcount, alive, part, T, lane);
    if (lane == 0 && gone) atomicAdd((unsigned long long *)&part[2 * (uint64_t)T], 0ull - gone);
}

// The whole findMatches loop (lib/kmerFinderClient.js:273-286) on the device.  Block 0 takes the argmax and the gate of a
// round and appends {winner, u, tau, H, z, p, u0, tau0} to `res`; every block then removes its share of the winner's
// k-mers; a grid barrier closes the round.  The host waits once and finishes the rows in exact decimal arithmetic.  The
// gate is the double-precision one: the loop stops BEFORE the removal at the first round whose z is within 1e-6 of a fastp
// threshold (or whose p * templates is within 1e-9 of the evalue) -- the exact arithmetic decides that round on the host
// and the loop is resumed.  Scores are changed by atomics (L2) and read back with volatile loads.  The blocks must be
// co-resident (cooperative launch, grid <= what the device holds at once).
struct KjWtaLoopArgs {
    KjDbDev d;
    uint64_t *glob;               // u[T], tau[T], H  (== part on one GPU / in a gathered match)
    const uint64_t *glob0;        // u[T], tau[T] of the first match
    const uint64_t *ford, *fidx;  // first-encounter keys (ties of the score)
    const uint64_t *ulen;
    const uint64_t *toff;
    const uint32_t *tq, *qkmer;
    const uint64_t *qcount;
    uint8_t *alive;
    uint32_t T;
    uint32_t max_rounds;          // rounds this launch may accept
    double unique_lens, n_templates;
    KjWtaResult *res;             // max_rounds + 1 records
    uint32_t *head;               // {records written, status}
    unsigned int *sync;           // [0]: barrier counter, [1 + round]: control word of the round (0: not yet, 1: stop, 2 + w: remove w)
};
enum { KJ_LOOP_MORE = 0,          // max_rounds accepted: the caller decides whether maxHits is reached
       KJ_LOOP_NO_HITS = 1,       // nHits === 0 / no template left (last record is not a row)
       KJ_LOOP_REJECTED = 2,      // the gate rejected the winner away from every threshold (last record is not a row)
       KJ_LOOP_UNDECIDED = 3 };   // the last record's winner has NOT been removed: exact arithmetic decides

__device__ __forceinline__ bool kj_gate_decisive(double z, double p) {
    const double thr[27] = {10.7016, 10.4862, 10.2663, 10.0416, 9.81197, 9.5769, 9.33604, 9.08895, 8.83511,
                            8.57394, 8.30479, 8.02686, 7.73926, 7.4409,  7.13051, 6.8065,  6.46695, 6.10941,
                            5.73073, 5.32672, 4.89164, 4.41717, 3.89059, 3.29053, 2.57583, 1.95996, 1.64485};
    if (!(z == z)) return false;
    for (int i = 0; i < 27; ++i) if (fabs(z - thr[i]) <= 1e-6) return false;
    if (fabs(p - 0.05) <= 1e-9) return false;
    return true;
}

__global__ void __launch_bounds__(256) kj_wta_loop_kernel(const KjWtaLoopArgs a) {
    __shared__ uint32_t s_ctl;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint64_t gwarp = (uint64_t)blockIdx.x * nw + warp, gwarps = (uint64_t)gridDim.x * nw;
    uint32_t n_rec = 0, status = KJ_LOOP_MORE, done_rounds = 0;
    for (uint32_t round = 0; round < a.max_rounds; ++round) {
        if (blockIdx.x == 0) {
            const uint32_t who = kj_block_winner(a.glob, a.ford, a.fidx, a.T);
            if (threadIdx.x == 0) {
                KjWtaResult r;
                r.winner = who; r.pad = 0;
                r.hits = kj_ld_volatile(&a.glob[2 * (uint64_t)a.T]);
                r.u = 0; r.tau = 0; r.z = 0.0; r.p = 1.0; r.u0 = 0; r.tau0 = 0;
                uint32_t ctl = 1;
                if (who == KJ_NONE32 || r.hits == 0) {
                    status = KJ_LOOP_NO_HITS;
                } else {
                    r.u = kj_ld_volatile(&a.glob[who]);
                    r.tau = kj_ld_volatile(&a.glob[(uint64_t)a.T + who]);
                    r.u0 = a.glob0[who];
                    r.tau0 = a.glob0[(uint64_t)a.T + who];
                    r.z = kj_zscore_f64((double)r.u, (double)a.ulen[who], (double)r.hits, a.unique_lens);
                    r.p = kj_fastp_f64(r.z) * a.n_templates;
                    if (!kj_gate_decisive(r.z, r.p)) status = KJ_LOOP_UNDECIDED;
                    else if (r.u > 0 && r.p <= 0.05) ctl = 2u + who;
                    else status = KJ_LOOP_REJECTED;
                }
                a.res[n_rec] = r;
                __threadfence();
                *reinterpret_cast<volatile unsigned int *>(&a.sync[1 + round]) = ctl;
            }
        }
        if (threadIdx.x == 0) {
            unsigned int c;
            do { c = kj_ld_volatile(&a.sync[1 + round]); } while (c == 0);
            s_ctl = c;
        }
        __syncthreads();
        const uint32_t ctl = s_ctl;
        ++n_rec;
        if (ctl < 2u) break;
        // removeWinnerKmers (lib/kmerFinderClient.js:220-230): a warp per 32 matched entries of the winner, the whole grid
        const uint32_t w = ctl - 2u;
        const uint64_t lo = a.toff[w], hi = a.toff[w + 1];
        unsigned long long gone = 0;
        for (uint64_t i = lo + 32 * gwarp; i < hi; i += 32 * gwarps)
            gone += kj_remove_group(a.d, a.tq, i, hi, a.qkmer, a.qcount, a.alive, a.glob, a.T, lane);
        if (lane == 0 && gone) atomicAdd((unsigned long long *)&a.glob[2 * (uint64_t)a.T], 0ull - gone);
        // grid barrier: every block's removals are in L2 before block 0 takes the next argmax
        __threadfence();
        __syncthreads();
        ++done_rounds;
        if (threadIdx.x == 0) {
            atomicAdd(&a.sync[0], 1u);
            const unsigned int target = done_rounds * gridDim.x;
            while (kj_ld_volatile(&a.sync[0]) < target) { }
            __threadfence();
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { a.head[0] = n_rec; a.head[1] = status; }
}

// ------------------------------------------------------------------------------------ database

static bool all_acgt(const uint8_t *p, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) if (!kj_is_acgt(p[i])) return false;
    return true;
}
static uint64_t pack_key(const uint8_t *p, uint32_t n) {
    uint64_t key = 0;
    for (uint32_t i = 0; i < n; ++i) key = (key << 2) | kj_code(p[i]);
    return key;
}
static uint64_t pow2_at_least(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

extern "C" void kj_db_free(kj_db *db) {
    if (!db) return;
    kj_ctx *ctx = db->ctx;
    if (ctx) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        cudaSetDevice(ctx->device);
        kj_dfree(ctx, db->d_keys); kj_dfree(ctx, db->d_vals); kj_dfree(ctx, db->d_list_off);
        kj_dfree(ctx, db->d_tmpl); kj_dfree(ctx, db->d_ulen);
    }
    delete db;
}

extern "C" int kj_db_create(kj_ctx *ctx, const kj_db_desc *d, kj_db **out) {
    if (!ctx || !d || !out) return kj_fail(ctx, KJ_E_INVALID, "kj_db_create: null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (d->n_kmers && (!d->kmer_bytes || !d->kmer_len || !d->list_off || !d->tmpl_ids))
        return kj_fail(ctx, KJ_E_INVALID, "kj_db_create: null k-mer arrays");
    if (d->n_templates && (!d->lengths || !d->ulengths))
        return kj_fail(ctx, KJ_E_INVALID, "kj_db_create: null template arrays");
    if (d->n_kmers > 0xFFFFFFF0ull) return kj_fail(ctx, KJ_E_RANGE, "more than 2^32 DB k-mers on one GPU");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    kj_db *db = new kj_db();
    db->ctx = ctx;
    db->n_templates = d->n_templates;
    db->lengths.assign(d->lengths, d->lengths + d->n_templates);
    db->ulengths.assign(d->ulengths, d->ulengths + d->n_templates);
    db->s_templates = d->summary_templates;
    db->s_unique_lens = d->summary_unique_lens;
    db->s_total_len = d->summary_total_len;
    const uint32_t n_parts = d->n_parts ? d->n_parts : 1;

    // regular k-mer length = length of the first ACGT-only k-mer
    uint64_t boff = 0;
    for (uint64_t i = 0; i < d->n_kmers && !db->k; ++i) {
        if (d->kmer_len[i] >= 1 && d->kmer_len[i] <= 32 && all_acgt(d->kmer_bytes + boff, d->kmer_len[i]))
            db->k = d->kmer_len[i];
        boff += d->kmer_len[i];
    }
    // keep the k-mers this part owns; rebuild the CSR with duplicate templates removed
    // (getMatches counts Set members, lib/kmerFinderClient.js:238-259: a template counts once per k-mer)
    std::vector<uint64_t> keys;
    std::vector<uint32_t> ids;
    std::vector<uint64_t> off(1, 0);
    std::vector<uint32_t> tm;
    std::vector<uint32_t> seen_stamp(d->n_templates, KJ_NONE32);
    boff = 0;
    uint32_t next_id = 0;
    for (uint64_t i = 0; i < d->n_kmers; ++i) {
        const uint8_t *kb = d->kmer_bytes + boff;
        const uint32_t len = d->kmer_len[i];
        boff += len;
        if (len > 32) { delete db; return kj_fail(ctx, KJ_E_RANGE, "DB k-mer longer than 32 bytes"); }
        const bool regular = db->k && len == db->k && all_acgt(kb, len);
        uint8_t pad[32] = {0};
        memcpy(pad, kb, len);
        const uint32_t owner = regular ? kj_owner_key(pack_key(kb, len), n_parts) : kj_owner_bytes(pad, len, n_parts);
        if (owner != d->part % n_parts) continue;
        const uint32_t id = next_id;
        if (regular) {
            uint64_t key = pack_key(kb, len);
            if (key == KJ_EMPTY) {
                if (db->special_id != KJ_NONE32) continue;      // duplicate k-mer: the first list wins
                db->special_id = id;
            } else { keys.push_back(key); ids.push_back(id); }
        } else {
            if (!db->other.emplace(std::string((const char *)kb, len), id).second) continue;
        }
        for (uint64_t j = d->list_off[i]; j < d->list_off[i + 1]; ++j) {
            uint32_t t = d->tmpl_ids[j];
            if (t >= d->n_templates) { delete db; return kj_fail(ctx, KJ_E_INVALID, "template id out of range"); }
            if (seen_stamp[t] == id) continue;
            seen_stamp[t] = id;
            tm.push_back(t);
        }
        off.push_back(tm.size());
        ++next_id;
    }
    db->n_kmers = next_id;
    db->n_pairs = tm.size();
    db->cap = pow2_at_least(std::max<uint64_t>(2 * keys.size(), 1024));

    cudaError_t e = kj_dmalloc(ctx, &db->d_keys, db->cap * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &db->d_vals, db->cap * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &db->d_list_off, off.size() * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &db->d_tmpl, std::max<size_t>(tm.size(), 1) * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &db->d_ulen, std::max<uint32_t>(d->n_templates, 1) * 8);
    uint64_t *d_in_keys = nullptr;
    uint32_t *d_in_ids = nullptr;
    unsigned int *d_dup = nullptr;
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_in_keys, std::max<size_t>(keys.size(), 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_in_ids, std::max<size_t>(ids.size(), 1) * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_dup, 4);
    if (e == cudaSuccess) e = cudaMemsetAsync(db->d_keys, 0xFF, db->cap * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(db->d_vals, 0xFF, db->cap * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_dup, 0, 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(db->d_list_off, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !tm.empty()) e = cudaMemcpyAsync(db->d_tmpl, tm.data(), tm.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && d->n_templates) e = cudaMemcpyAsync(db->d_ulen, d->ulengths, (size_t)d->n_templates * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !keys.empty()) e = cudaMemcpyAsync(d_in_keys, keys.data(), keys.size() * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !ids.empty()) e = cudaMemcpyAsync(d_in_ids, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
    unsigned int dup = 0;
    if (e == cudaSuccess && !keys.empty()) {
        KJ_LAUNCH(kj_db_build_kernel, kj_grid_for(ctx, keys.size()), 256, 0, ctx->stream, db->d_keys, db->d_vals,
                  db->cap - 1, d_in_keys, d_in_ids, (uint64_t)keys.size(), d_dup);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(&dup, d_dup, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);   // host vectors go out of scope
    kj_dfree(ctx, d_in_keys); kj_dfree(ctx, d_in_ids); kj_dfree(ctx, d_dup);
    if (e != cudaSuccess) {
        kj_db_free(db);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_db_create: ") + cudaGetErrorString(e));
    }
    if (dup) {
        kj_db_free(db);
        return kj_fail(ctx, KJ_E_INVALID, "kj_db_create: the same k-mer appears more than once in the DB");
    }
    *out = db;
    return KJ_OK;
}

extern "C" uint64_t kj_db_n_kmers(const kj_db *db) { return db ? db->n_kmers : 0; }
extern "C" uint64_t kj_db_n_pairs(const kj_db *db) { return db ? db->n_pairs : 0; }
extern "C" uint32_t kj_db_n_templates(const kj_db *db) { return db ? db->n_templates : 0; }

// ------------------------------------------------------------------------------------ first match

extern "C" void kj_match_free(kj_match *m) {
    if (!m) return;
    kj_ctx *ctx = m->ctx;
    if (ctx) {
        std::lock_guard<std::recursive_mutex> lk(ctx->mu);
        cudaSetDevice(ctx->device);
        kj_dfree(ctx, m->d_qkmer); kj_dfree(ctx, m->d_msize);
        kj_dfree(ctx, m->own_count); kj_dfree(ctx, m->own_ord); kj_dfree(ctx, m->own_off);
        kj_dfree(ctx, m->own_alive); kj_dfree(ctx, m->own_tmpl);
        if (m->d_glob != m->d_part) kj_dfree(ctx, m->d_glob);
        kj_dfree(ctx, m->d_part);
        kj_dfree(ctx, m->d_first_ord); kj_dfree(ctx, m->d_first_idx);
        kj_dfree(ctx, m->d_toff); kj_dfree(ctx, m->d_tcur); kj_dfree(ctx, m->d_tq); kj_dfree(ctx, m->d_res);
        kj_dfree(ctx, m->d_loop); kj_dfree(ctx, m->d_glob0); kj_dfree(ctx, m->d_sync);
        kj_pinned_put(ctx, m->h_res);
    }
    delete m;
}

static KjWalkArgs walk_args(const kj_match *m) {
    KjWalkArgs a{};
    a.d = m->d;
    a.qkmer = m->d_qkmer;
    a.qcount = m->qcount;
    a.qord = m->qord;
    a.alive = m->alive;
    a.Q = m->Q;
    a.T = m->T;
    a.part = m->d_part;
    a.first_ord = m->d_first_ord;
    a.first_idx = m->d_first_idx;
    a.toff = m->d_toff;
    a.tcur = m->d_tcur;
    a.tq = m->d_tq;
    return a;
}

template <int MODE>
static int launch_walk(kj_match *m) {
    kj_ctx *ctx = m->ctx;
    if (!m->Q) return KJ_OK;
    KjWalkArgs a = walk_args(m);
    const uint64_t groups = (m->Q + 31) / 32;
    const int warps_per_block = KJ_SCORE_THREADS / 32;
    int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + warps_per_block - 1) / warps_per_block,
                                                               (uint64_t)ctx->sm_count * 4));
    const bool smem = MODE == KJ_WALK_ACCUM && m->T <= KJ_SMEM_T_MAX && m->T > 0;
    if (smem) {
        // privatised histograms: fewer, fatter blocks (one flush of T atomics per block)
        grid = std::min(grid, ctx->sm_count * 2);
        const size_t bytes = (size_t)m->T * 12 + 16;
        KJ_CUDA(ctx, cudaFuncSetAttribute(kj_walk_kernel<MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
        KJ_LAUNCH((kj_walk_kernel<MODE, true>), grid, KJ_SCORE_THREADS, bytes, ctx->stream, a);
    } else {
        KJ_LAUNCH((kj_walk_kernel<MODE, false>), grid, KJ_SCORE_THREADS, 0, ctx->stream, a);
    }
    ctx->launches++;
    KJ_CUDA(ctx, cudaGetLastError());
    return KJ_OK;
}

// resolve the query entries, accumulate this rank's partial scores and first-encounter ordinals
extern "C" int kj_first_match_local(kj_ctx *ctx, kj_counts *q, const kj_db *db, kj_match **out) {
    if (!ctx || !q || !db || !out) return kj_fail(ctx, KJ_E_INVALID, "kj_first_match: null argument");
    int rc = kj_counts_check_finished(q);
    if (rc) return rc;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    kj_match *m = new kj_match();
    m->ctx = ctx; m->q = q; m->db = db;
    m->qcount = q->reg.counts; m->qord = q->reg.ords; m->alive = q->reg.alive;
    m->d = db->dev();
    m->T = db->n_templates;
    m->Q = q->reg.n;
    m->kmer_map_size = q->reg.n;
    m->pair_bound = db->n_pairs;                      // every matched list is one of the DB's
    const uint64_t T = m->T;
    cudaError_t e = kj_dmalloc(ctx, &m->d_qkmer, std::max<uint64_t>(m->Q, 1) * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_part, (2 * T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_ord, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_idx, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_toff, (T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_tcur, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_res, sizeof(KjWtaResult));
    if (e == cudaSuccess) {
        m->h_res = (KjWtaResult *)kj_pinned_get(ctx);
        if (!m->h_res) e = cudaErrorMemoryAllocation;
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_qkmer, 0xFF, std::max<uint64_t>(m->Q, 1) * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_part, 0, (2 * T + 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_ord, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_idx, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e != cudaSuccess) {
        kj_match_free(m);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_first_match: ") + cudaGetErrorString(e));
    }
    m->d_glob = m->d_part;

    // query entry -> DB k-mer id
    const uint64_t n_tab = q->reg.n_tab, n_reg = q->reg.n_reg;
    if (q->k == db->k) {
        if (n_tab) {
            KJ_LAUNCH(kj_probe_kernel, kj_grid_for(ctx, n_tab), 256, 0, ctx->stream, db->dev(), q->reg.keys, n_tab,
                      m->d_qkmer);
            ctx->launches++;
        }
        if (n_reg > n_tab && db->special_id != KJ_NONE32)
            e = cudaMemcpyAsync(m->d_qkmer + n_tab, &db->special_id, 4, cudaMemcpyHostToDevice, ctx->stream);
    } else if (!db->other.empty() && n_reg) {
        // query k differs from the DB's regular length: only byte-string DB entries can match
        std::vector<uint64_t> hk(n_reg);
        std::vector<uint32_t> ids(n_reg, KJ_NONE32);
        e = cudaMemcpyAsync(hk.data(), q->reg.keys, n_reg * 8, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        uint8_t buf[32];
        for (uint64_t i = 0; i < n_reg && e == cudaSuccess; ++i) {
            kj_decode_key(hk[i], q->k, buf);
            auto it = db->other.find(std::string((const char *)buf, q->k));
            if (it != db->other.end()) ids[i] = it->second;
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_qkmer, ids.data(), n_reg * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    const uint64_t n_irr = m->Q - n_reg;
    if (e == cudaSuccess && n_irr && !db->other.empty()) {
        const KjIrrRecord *ir = reinterpret_cast<const KjIrrRecord *>(q->irr_host.data());
        std::vector<uint32_t> ids(n_irr, KJ_NONE32);
        for (uint64_t i = 0; i < n_irr; ++i) {
            auto it = db->other.find(std::string((const char *)ir[i].key, (size_t)ir[i].len));
            if (it != db->other.end()) ids[i] = it->second;
        }
        e = cudaMemcpyAsync(m->d_qkmer + n_reg, ids.data(), n_irr * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    }
    if (e != cudaSuccess) {
        kj_match_free(m);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_first_match: ") + cudaGetErrorString(e));
    }
    rc = launch_walk<KJ_WALK_ACCUM>(m);
    if (rc) { kj_match_free(m); return rc; }
    *out = m;
    return KJ_OK;
}

// ------------------------------------------------------------------------------------ gathered matches

static int launch_export(kj_match *m, bool count, void *dev_entries, uint64_t cap_entries, void *dev_tmpl,
                         uint64_t cap_pairs) {
    kj_ctx *ctx = m->ctx;
    if (!m->d_msize) KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_msize, 16));
    KJ_CUDA(ctx, cudaMemsetAsync(m->d_msize, 0, 16, ctx->stream));
    if (!m->Q) return KJ_OK;
    const uint64_t groups = (m->Q + 31) / 32;
    const int wpb = KJ_SCORE_THREADS / 32;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + wpb - 1) / wpb, (uint64_t)ctx->sm_count * 8));
    if (count)
        KJ_LAUNCH((kj_matched_export_kernel<true>), grid, KJ_SCORE_THREADS, 0, ctx->stream, m->d, m->d_qkmer, m->qcount,
                  m->qord, m->alive, m->Q, (uint64_t *)nullptr, (uint64_t)0, (uint32_t *)nullptr, (uint64_t)0, m->d_msize);
    else
        KJ_LAUNCH((kj_matched_export_kernel<false>), grid, KJ_SCORE_THREADS, 0, ctx->stream, m->d, m->d_qkmer, m->qcount,
                  m->qord, m->alive, m->Q, (uint64_t *)dev_entries, cap_entries, (uint32_t *)dev_tmpl, cap_pairs, m->d_msize);
    ctx->launches++;
    KJ_CUDA(ctx, cudaGetLastError());
    return KJ_OK;
}

extern "C" int kj_match_matched_size(kj_match *m, uint64_t *n_entries, uint64_t *n_pairs) {
    if (!m || !n_entries || !n_pairs) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = launch_export(m, true, nullptr, 0, nullptr, 0);
    if (rc) return rc;
    unsigned long long *h = (unsigned long long *)m->h_res;      // pinned scratch of the match
    KJ_CUDA(ctx, cudaMemcpyAsync(h, m->d_msize, 16, cudaMemcpyDeviceToHost, ctx->stream));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *n_entries = h[0];
    *n_pairs = h[1];
    return KJ_OK;
}

extern "C" int kj_match_export_matched(kj_match *m, void *dev_entries, uint64_t cap_entries, void *dev_tmpl,
                                       uint64_t cap_pairs) {
    if (!m || !dev_entries || !dev_tmpl) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    return launch_export(m, false, dev_entries, cap_entries, dev_tmpl, cap_pairs);     // stream-ordered, no sync
}

extern "C" int kj_match_from_matched(kj_ctx *ctx, const kj_db *db, uint32_t n_segments, const uint64_t *seg_entries,
                                     const uint64_t *seg_n_entries, const uint64_t *seg_tmpl, const uint64_t *seg_n_pairs,
                                     uint64_t kmer_map_size, kj_match **out) {
    if (!ctx || !db || !out || (n_segments && (!seg_entries || !seg_n_entries || !seg_tmpl || !seg_n_pairs)))
        return kj_fail(ctx, KJ_E_INVALID, "kj_match_from_matched: null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t Q = 0, P = 0;
    for (uint32_t s = 0; s < n_segments; ++s) { Q += seg_n_entries[s]; P += seg_n_pairs[s]; }
    if (2 * Q >= KJ_NONE32) return kj_fail(ctx, KJ_E_RANGE, "too many matched entries for a gathered match");
    kj_match *m = new kj_match();
    m->ctx = ctx; m->q = nullptr; m->db = db;
    m->T = db->n_templates;
    m->Q = Q;
    m->kmer_map_size = kmer_map_size;
    const uint64_t T = m->T;
    unsigned int *d_bad = nullptr;
    cudaError_t e = kj_dmalloc(ctx, &m->d_qkmer, std::max<uint64_t>(Q, 1) * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_count, std::max<uint64_t>(Q, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_ord, std::max<uint64_t>(Q, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_off, std::max<uint64_t>(2 * Q, 2) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_alive, std::max<uint64_t>(Q, 1));
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_tmpl, std::max<uint64_t>(P, 1) * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_part, (2 * T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_ord, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_idx, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_toff, (T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_tcur, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_res, sizeof(KjWtaResult));
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &d_bad, 4);
    if (e == cudaSuccess) {
        m->h_res = (KjWtaResult *)kj_pinned_get(ctx);
        if (!m->h_res) e = cudaErrorMemoryAllocation;
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_part, 0, (2 * T + 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_ord, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_idx, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_bad, 0, 4, ctx->stream);
    uint64_t e0 = 0, p0 = 0;
    for (uint32_t s = 0; s < n_segments && e == cudaSuccess; ++s) {
        const uint64_t n = seg_n_entries[s], np = seg_n_pairs[s];
        if (n) {
            KJ_LAUNCH(kj_matched_import_kernel, kj_grid_for(ctx, n), 256, 0, ctx->stream,
                      (const uint64_t *)(uintptr_t)seg_entries[s], n, e0, p0, p0 + np, m->own_count, m->own_ord,
                      m->own_alive, m->d_qkmer, m->own_off, d_bad);
            ctx->launches++;
        }
        if (np) e = cudaMemcpyAsync(m->own_tmpl + p0, (const void *)(uintptr_t)seg_tmpl[s], np * 4,
                                    cudaMemcpyDeviceToDevice, ctx->stream);
        e0 += n; p0 += np;
    }
    unsigned int bad = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    kj_dfree(ctx, d_bad);
    if (e != cudaSuccess || bad) {
        kj_match_free(m);
        if (bad) return kj_fail(ctx, KJ_E_INVALID, "kj_match_from_matched: a record's template list lies outside its segment");
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_match_from_matched: ") + cudaGetErrorString(e));
    }
    m->d_glob = m->d_part;
    m->qcount = m->own_count; m->qord = m->own_ord; m->alive = m->own_alive;
    m->d = KjDbDev{nullptr, nullptr, 0, m->own_off, m->own_tmpl};
    int rc = launch_walk<KJ_WALK_ACCUM>(m);
    if (rc) { kj_match_free(m); return rc; }
    *out = m;
    return KJ_OK;
}

extern "C" uint64_t kj_matched_segment_bytes(uint32_t cap_entries, uint32_t cap_pairs) {
    const uint64_t b = sizeof(KjMSegHeader) + (uint64_t)cap_entries * 32u + (uint64_t)cap_pairs * 4u;
    return (b + 255) / 256 * 256;
}

// the matched entries of this rank into one fixed-capacity segment (stream-ordered, no host wait); query_size = the
// number of k-mers this rank owns, flags != 0 marks the rank's part of the job as unusable (the gathered match fails)
extern "C" int kj_match_export_segment(kj_match *m, void *dev_segment, uint32_t cap_entries, uint32_t cap_pairs,
                                       uint64_t query_size, uint64_t flags) {
    if (!m || !dev_segment) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *seg = reinterpret_cast<uint8_t *>(dev_segment);
    KJ_LAUNCH(kj_matched_header_kernel, 1, 1, 0, ctx->stream, reinterpret_cast<KjMSegHeader *>(seg),
              (unsigned long long)query_size, (unsigned long long)flags);
    ctx->launches++;
    if (m->Q) {
        const uint64_t groups = (m->Q + 31) / 32;
        const int wpb = KJ_SCORE_THREADS / 32;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + wpb - 1) / wpb, (uint64_t)ctx->sm_count * 8));
        KJ_LAUNCH((kj_matched_export_kernel<false>), grid, KJ_SCORE_THREADS, 0, ctx->stream, m->d, m->d_qkmer, m->qcount,
                  m->qord, m->alive, m->Q, reinterpret_cast<uint64_t *>(seg + sizeof(KjMSegHeader)), (uint64_t)cap_entries,
                  reinterpret_cast<uint32_t *>(seg + sizeof(KjMSegHeader) + (uint64_t)cap_entries * 32u), (uint64_t)cap_pairs,
                  reinterpret_cast<unsigned long long *>(seg));
        ctx->launches++;
    }
    KJ_CUDA(ctx, cudaGetLastError());
    return KJ_OK;
}

// The same segment straight from a handle's hash table, without kj_counts_finish in between (the owner of a fixed-capacity
// exchange: its finish -- compaction, counters back, a host wait -- then leaves the chain count -> exchange -> gather ->
// winner-takes-all and only has to be done before the k-mer map itself is read).  Table slots play the query entries.
// Only regular k-mers can hit here, so the short cut applies when the DB holds nothing else (no byte-string k-mers, no
// all-G 32-mer) and k matches; returns 1 (not an error) when it does not and the caller finishes first.  The query size and
// the "this rank's exchange did not fit" flag are read from the handle's device counters.
__global__ void kj_probe_table_kernel(KjDbDev d, const uint64_t *tkeys, uint64_t cap, uint32_t *qkmer, uint8_t *alive) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < cap; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t key = tkeys[i];
        qkmer[i] = key == KJ_EMPTY ? KJ_NONE32 : kj_db_lookup(d, key);
        alive[i] = 1;
    }
}
__global__ void kj_matched_header_counts_kernel(KjMSegHeader *h, const KjCounters *ctr) {
    unsigned long long n = 0;
    for (int i = 0; i < 64; ++i) n += ctr->n_unique_part[i];
    h->n_entries = 0; h->n_pairs = 0;
    h->qsize = n + ctr->n_irr_unique + (ctr->special_count ? 1ull : 0ull);
    h->flags = (ctr->error_flags || ctr->n_overflow || ctr->n_irr_overflow) ? 1ull : 0ull;
}
extern "C" int kj_counts_export_matched_segment(kj_counts *c, const kj_db *db, void *dev_segment, uint32_t cap_entries,
                                                uint32_t cap_pairs) {
    if (!c || !db || !dev_segment) return KJ_E_INVALID;
    kj_ctx *ctx = c->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (c->k != db->k || c->k >= 32 || !db->other.empty() || db->special_id != KJ_NONE32 || c->pending) return 1;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint8_t *seg = reinterpret_cast<uint8_t *>(dev_segment);
    KJ_LAUNCH(kj_matched_header_counts_kernel, 1, 1, 0, ctx->stream, reinterpret_cast<KjMSegHeader *>(seg), c->ctr);
    ctx->launches++;
    if (c->cap) {
        uint32_t *d_qkmer = nullptr;
        uint8_t *d_alive = nullptr;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &d_qkmer, c->cap * 4));
        cudaError_t e = kj_dmalloc(ctx, &d_alive, c->cap);
        if (e != cudaSuccess) { kj_dfree(ctx, d_qkmer); return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e)); }
        KJ_LAUNCH(kj_probe_table_kernel, kj_grid_for(ctx, c->cap), 256, 0, ctx->stream, db->dev(), c->tab.keys, c->cap, d_qkmer, d_alive);
        const uint64_t groups = (c->cap + 31) / 32;
        const int wpb = KJ_SCORE_THREADS / 32;
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((groups + wpb - 1) / wpb, (uint64_t)ctx->sm_count * 8));
        KJ_LAUNCH((kj_matched_export_kernel<false>), grid, KJ_SCORE_THREADS, 0, ctx->stream, db->dev(), d_qkmer, c->tab.counts,
                  c->tab.ords, d_alive, c->cap, reinterpret_cast<uint64_t *>(seg + sizeof(KjMSegHeader)), (uint64_t)cap_entries,
                  reinterpret_cast<uint32_t *>(seg + sizeof(KjMSegHeader) + (uint64_t)cap_entries * 32u), (uint64_t)cap_pairs,
                  reinterpret_cast<unsigned long long *>(seg));
        ctx->launches += 2;
        kj_dfree(ctx, d_qkmer); kj_dfree(ctx, d_alive);       // stream-ordered: after the kernels
    }
    KJ_CUDA(ctx, cudaGetLastError());
    return KJ_OK;
}

// a match over the gathered segments of every rank; the buffer must stay valid until kj_match_free.  Sizes, the query
// size and the flags are read on the device: they surface in kj_match_commit (KJ_E_RANGE when a segment overflowed or
// a rank flagged its part).
extern "C" int kj_match_from_segments(kj_ctx *ctx, const kj_db *db, uint32_t n_segments, const void *dev_segments,
                                      uint32_t cap_entries, uint32_t cap_pairs, kj_match **out) {
    if (!ctx || !db || !out || !dev_segments || !n_segments || !cap_entries)
        return kj_fail(ctx, KJ_E_INVALID, "kj_match_from_segments: bad argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t Q = (uint64_t)n_segments * cap_entries;
    const uint64_t seg_bytes = kj_matched_segment_bytes(cap_entries, cap_pairs);
    if (2 * Q >= KJ_NONE32 || (uint64_t)n_segments * seg_bytes / 4 >= (1ull << 40))
        return kj_fail(ctx, KJ_E_RANGE, "too many matched entries for a gathered match");
    kj_match *m = new kj_match();
    m->ctx = ctx; m->q = nullptr; m->db = db;
    m->T = db->n_templates;
    m->Q = Q;
    m->from_segments = true;
    m->pair_bound = (uint64_t)n_segments * cap_pairs;
    const uint64_t T = m->T;
    cudaError_t e = kj_dmalloc(ctx, &m->d_qkmer, Q * 4);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_count, Q * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_ord, Q * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_off, 2 * Q * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->own_alive, Q);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_msize, 32);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_part, (2 * T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_ord, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_first_idx, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_toff, (T + 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_tcur, std::max<uint64_t>(T, 1) * 8);
    if (e == cudaSuccess) e = kj_dmalloc(ctx, &m->d_res, sizeof(KjWtaResult));
    if (e == cudaSuccess) {
        m->h_res = (KjWtaResult *)kj_pinned_get(ctx);
        if (!m->h_res) e = cudaErrorMemoryAllocation;
    }
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_part, 0, (2 * T + 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_ord, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_first_idx, 0xFF, std::max<uint64_t>(T, 1) * 8, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->own_alive, 0, Q, ctx->stream);                // the unused tail stays dead
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_qkmer, 0xFF, Q * 4, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(m->d_msize, 0, 32, ctx->stream);
    if (e != cudaSuccess) {
        kj_match_free(m);
        return kj_fail(ctx, KJ_E_CUDA, std::string("kj_match_from_segments: ") + cudaGetErrorString(e));
    }
    KJ_LAUNCH(kj_matched_import_segments_kernel, kj_grid_for(ctx, Q), 256, 0, ctx->stream,
              reinterpret_cast<const uint8_t *>(dev_segments), seg_bytes, n_segments, cap_entries, cap_pairs, m->own_count,
              m->own_ord, m->own_alive, m->d_qkmer, m->own_off, m->d_msize);
    ctx->launches++;
    m->d_glob = m->d_part;
    m->qcount = m->own_count; m->qord = m->own_ord; m->alive = m->own_alive;
    m->d = KjDbDev{nullptr, nullptr, 0, m->own_off, reinterpret_cast<const uint32_t *>(dev_segments)};
    int rc = launch_walk<KJ_WALK_ACCUM>(m);
    if (rc) { kj_match_free(m); return rc; }
    *out = m;
    return KJ_OK;
}

// vectors the host layer reduces over ranks (device memory, caller-provided buffers)
static int vec_info(kj_match *m, int which, uint64_t **ptr, uint64_t *n) {
    switch (which) {
        case KJ_VEC_SCORES: *ptr = m->d_part; *n = 2 * (uint64_t)m->T + 1; return KJ_OK;
        case KJ_VEC_FIRST_ORD: *ptr = m->d_first_ord; *n = m->T; return KJ_OK;
        case KJ_VEC_FIRST_IDX: *ptr = m->d_first_idx; *n = m->T; return KJ_OK;
        default: return kj_fail(m->ctx, KJ_E_INVALID, "unknown vector id");
    }
}

extern "C" uint64_t kj_match_vec_len(kj_match *m, int which) {
    uint64_t *p = nullptr, n = 0;
    if (!m || vec_info(m, which, &p, &n)) return 0;
    return n;
}

static int match_get(kj_match *m, int which, void *dev_out, bool sync) {
    if (!m || !dev_out) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *p = nullptr, n = 0;
    int rc = vec_info(m, which, &p, &n);
    if (rc) return rc;
    if (which == KJ_VEC_FIRST_IDX) {
        // second stage of the (ordinal, list index) minimum: only entries whose ordinal equals the
        // (by now global) first_ord contribute
        KJ_CUDA(ctx, cudaMemsetAsync(m->d_first_idx, 0xFF, std::max<uint64_t>(m->T, 1) * 8, ctx->stream));
        rc = launch_walk<KJ_WALK_FIRST>(m);
        if (rc) return rc;
    }
    if (n) KJ_CUDA(ctx, cudaMemcpyAsync(dev_out, p, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (sync) KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KJ_OK;
}
extern "C" int kj_match_get(kj_match *m, int which, void *dev_out) { return match_get(m, which, dev_out, true); }
extern "C" int kj_match_get_async(kj_match *m, int which, void *dev_out) { return match_get(m, which, dev_out, false); }

static int match_set(kj_match *m, int which, const void *dev_in, bool sync) {
    if (!m || !dev_in) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *p = nullptr, n = 0;
    int rc = vec_info(m, which, &p, &n);
    if (rc) return rc;
    if (which == KJ_VEC_SCORES) {
        // global sums live beside the partial ones from now on
        if (m->d_glob == m->d_part) {
            uint64_t *g = nullptr;
            KJ_CUDA(ctx, kj_dmalloc(ctx, &g, n * 8));
            m->d_glob = g;
            m->distributed = true;
        }
        p = m->d_glob;
    }
    if (n) KJ_CUDA(ctx, cudaMemcpyAsync(p, dev_in, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    if (sync) KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KJ_OK;
}
extern "C" int kj_match_set(kj_match *m, int which, const void *dev_in) { return match_set(m, which, dev_in, true); }
extern "C" int kj_match_set_async(kj_match *m, int which, const void *dev_in) { return match_set(m, which, dev_in, false); }

extern "C" int kj_match_set_query_size(kj_match *m, uint64_t kmer_map_size) {
    if (!m) return KJ_E_INVALID;
    m->kmer_map_size = kmer_map_size;
    return KJ_OK;
}

// u0 / t0 / order / toff_h on the host: what the stepwise calls and the result accessors need.  The hot path
// (kj_first_match -> kj_wta_all) never does.
static int ensure_host_first(kj_match *m) {
    if (m->host_first) return KJ_OK;
    kj_ctx *ctx = m->ctx;
    const uint64_t T = m->T;
    std::vector<uint64_t> ford(T), fidx(T);
    m->u0.assign(T, 0); m->t0.assign(T, 0); m->toff_h.assign(T + 1, 0);
    if (T) {
        KJ_CUDA(ctx, cudaMemcpyAsync(m->u0.data(), m->d_glob0, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(m->t0.data(), m->d_glob0 + T, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(ford.data(), m->d_first_ord, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaMemcpyAsync(fidx.data(), m->d_first_idx, T * 8, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KJ_CUDA(ctx, cudaMemcpyAsync(m->toff_h.data(), m->d_toff, (T + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    // templates in order of first encounter (lib/kmerFinderServer.js:180-201) = ascending (first ordinal, list index)
    m->order.clear();
    for (uint32_t t = 0; t < T; ++t) if (m->u0[t]) m->order.push_back(t);
    std::sort(m->order.begin(), m->order.end(), [&](uint32_t a, uint32_t b) {
        if (ford[a] != ford[b]) return ford[a] < ford[b];
        if (fidx[a] != fidx[b]) return fidx[a] < fidx[b];
        return a < b;
    });
    m->host_first = true;
    return KJ_OK;
}

__global__ void kj_toff_first_kernel(uint64_t *toff) { toff[0] = 0; }

// first-encounter rank, first-round scores, per-template lists of matched query entries: all on the device; the host
// waits once, for the number of hits (it sizes the lists and decides 'No hits were found!')
extern "C" int kj_match_commit(kj_match *m) {
    if (!m) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    if (m->committed) return kj_fail(ctx, KJ_E_STATE, "kj_match_commit called twice");
    const uint64_t T = m->T;
    int rc;
    // {global hits, this rank's hits, gathered-segment info[4]} come back in one pinned block
    unsigned long long *h = reinterpret_cast<unsigned long long *>(m->h_res);
    KJ_CUDA(ctx, cudaMemcpyAsync(&h[0], m->d_glob + 2 * T, 8, cudaMemcpyDeviceToHost, ctx->stream));
    KJ_CUDA(ctx, cudaMemcpyAsync(&h[1], m->d_part + 2 * T, 8, cudaMemcpyDeviceToHost, ctx->stream));
    h[2] = h[3] = h[4] = h[5] = 0;
    if (m->from_segments) KJ_CUDA(ctx, cudaMemcpyAsync(&h[2], m->d_msize, 32, cudaMemcpyDeviceToHost, ctx->stream));
    // meanwhile: the first-round scores aside, the offsets of the per-template lists
    KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_glob0, std::max<uint64_t>(2 * T, 1) * 8));
    if (T) {
        KJ_CUDA(ctx, cudaMemcpyAsync(m->d_glob0, m->d_glob, 2 * T * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        KJ_CUDA(ctx, cudaMemsetAsync(m->d_tcur, 0, T * 8, ctx->stream));
    }
    KJ_LAUNCH(kj_toff_first_kernel, 1, 1, 0, ctx->stream, m->d_toff);
    if (T) {
#ifdef KJ_CPU_EMU
        for (uint64_t t = 0; t < T; ++t) m->d_toff[t + 1] = m->d_toff[t] + m->d_part[t];
#else
        size_t tmp_bytes = 0;
        void *d_tmp = nullptr;
        KJ_CUDA(ctx, cub::DeviceScan::InclusiveSum(nullptr, tmp_bytes, m->d_part, m->d_toff + 1, (int)T, ctx->stream));
        KJ_CUDA(ctx, kj_dmalloc(ctx, &d_tmp, std::max<size_t>(tmp_bytes, 16)));
        cudaError_t e = cub::DeviceScan::InclusiveSum(d_tmp, tmp_bytes, m->d_part, m->d_toff + 1, (int)T, ctx->stream);
        kj_dfree(ctx, d_tmp);
        if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
        ctx->launches += 2;
#endif
    }
    // With a bound on the matched list entries the per-template lists are filled before the host knows the hit count: the
    // one wait of the commit then has nothing queued behind it (without: wait, size the list array, fill)
    const bool fill_ahead = m->pair_bound && m->pair_bound <= (1ull << 28);
    if (fill_ahead) {
        kj_dfree(ctx, m->d_tq);
        m->d_tq = nullptr;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_tq, std::max<uint64_t>(m->pair_bound, 1) * 4));
        rc = m->distributed ? launch_walk<KJ_WALK_FILL>(m) : launch_walk<KJ_WALK_FIRST_FILL>(m);
        if (rc) return rc;
    }
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (m->from_segments) {
        if (h[5])
            return kj_fail(ctx, KJ_E_RANGE, (h[5] & 2) ? "gathered match: a segment overflowed its capacity"
                                                       : "gathered match: a rank flagged its part of the job");
        m->kmer_map_size = h[4];
        m->seg_entries = h[2]; m->seg_pairs = h[3];
    }
    m->hits0 = h[0];
    const uint64_t local_pairs = h[1];
    if (local_pairs > 0xFFFFFFF0ull * 16) return kj_fail(ctx, KJ_E_RANGE, "too many matched pairs");
    if (!fill_ahead) {
        kj_dfree(ctx, m->d_tq);
        m->d_tq = nullptr;
        KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_tq, std::max<uint64_t>(local_pairs, 1) * 4));
        // stream-ordered: whoever reads the lists comes later on the same stream.  On a single GPU first_ord is already global:
        // the list indices of the first k-mers (the tie order) are taken by the same pass
        rc = m->distributed ? launch_walk<KJ_WALK_FILL>(m) : launch_walk<KJ_WALK_FIRST_FILL>(m);
        if (rc) return rc;
    } else if (local_pairs > m->pair_bound) {
        return kj_fail(ctx, KJ_E_CUDA, "internal: matched list entries exceed their bound");
    }
    m->committed = true;
    return KJ_OK;
}

extern "C" int kj_first_match(kj_ctx *ctx, kj_counts *q, const kj_db *db, kj_match **out) {
    kj_match *m = nullptr;
    int rc = kj_first_match_local(ctx, q, db, &m);
    if (rc) return rc;
    rc = kj_match_commit(m);
    if (rc) { kj_match_free(m); return rc; }
    if (m->hits0 == 0) {
        // lib/kmerFinderServer.js:219-221; the client rejects with the same text (:159-161)
        kj_match_free(m);
        return kj_fail(ctx, KJ_E_NO_HITS, "No hits were found!");
    }
    *out = m;
    return KJ_OK;
}

extern "C" uint64_t kj_match_hits(const kj_match *m) { return m ? m->hits0 : 0; }
extern "C" uint64_t kj_match_query_size(const kj_match *m) { return m ? m->kmer_map_size : 0; }
extern "C" int kj_match_segment_sizes(const kj_match *m, uint64_t *n_entries, uint64_t *n_pairs) {
    if (!m || !n_entries || !n_pairs) return KJ_E_INVALID;
    *n_entries = m->seg_entries; *n_pairs = m->seg_pairs;
    return KJ_OK;
}
extern "C" uint32_t kj_match_n_matched(const kj_match *m) {
    if (!m || !m->committed || ensure_host_first(const_cast<kj_match *>(m))) return 0;
    return (uint32_t)m->order.size();
}

extern "C" int kj_match_scores(kj_match *m, uint64_t *uscore, uint64_t *tscore, uint32_t *order) {
    if (!m) return KJ_E_INVALID;
    if (!m->committed) return kj_fail(m->ctx, KJ_E_STATE, "kj_match_commit has not run");
    { int rc0 = ensure_host_first(m); if (rc0) return rc0; }
    if (uscore) memcpy(uscore, m->u0.data(), m->u0.size() * 8);
    if (tscore) memcpy(tscore, m->t0.data(), m->t0.size() * 8);
    if (order) memcpy(order, m->order.data(), m->order.size() * 4);
    return KJ_OK;
}

// The `kmers` Set of one template in the findFirstMatch reply (lib/kmerFinderClient.js:150-157,
// lib/kmerFinderServer.js:190-199): the query k-mers that list the template, as positions in the export
// order of the counts handle (ascending = the Set's insertion order).  Only this rank's share when the
// query is sharded.
extern "C" int kj_match_template_kmers(kj_match *m, uint32_t template_id, uint64_t *idx, uint64_t cap, uint64_t *n_out) {
    if (!m || !n_out) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!m->committed) return kj_fail(ctx, KJ_E_STATE, "kj_match_commit has not run");
    if (template_id >= m->T) return kj_fail(ctx, KJ_E_INVALID, "template id out of range");
    if (!m->q) return kj_fail(ctx, KJ_E_STATE, "a gathered match has no counts handle: ask the local match");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc0 = ensure_host_first(m); if (rc0) return rc0; }
    const uint64_t lo = m->toff_h[template_id], hi = m->toff_h[template_id + 1];
    *n_out = hi - lo;
    if (!idx || hi == lo) return KJ_OK;
    if (cap < hi - lo) return kj_fail(ctx, KJ_E_RANGE, "index buffer too small");
    std::vector<uint32_t> q(hi - lo);
    KJ_CUDA(ctx, cudaMemcpyAsync(q.data(), m->d_tq + lo, (hi - lo) * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const std::vector<uint64_t> *inv = nullptr;
    int rc = kj_counts_export_rank(m->q, &inv);       // query index -> export position
    if (rc) return rc;
    for (uint64_t i = 0; i < hi - lo; ++i) idx[i] = (*inv)[q[i]];
    std::sort(idx, idx + (hi - lo));
    return KJ_OK;
}

extern "C" int kj_match_set_max_hits(kj_match *m, uint32_t max_hits) {
    if (!m) return KJ_E_INVALID;
    m->max_hits = max_hits;
    return KJ_OK;
}

// ------------------------------------------------------------------------------------ winner takes all

// removeWinnerKmers (lib/kmerFinderClient.js:220-230) on this rank's share of K_w
static int wta_launch_remove(kj_match *m, uint32_t w) {
    kj_ctx *ctx = m->ctx;
    { int rc0 = ensure_host_first(m); if (rc0) return rc0; }
    const uint64_t range[2] = {m->toff_h[w], m->toff_h[w + 1]};
    if (range[1] > range[0]) {
        const uint64_t n = range[1] - range[0];
        const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(((n + 31) / 32 + 7) / 8, (uint64_t)ctx->sm_count * 8));
        KJ_LAUNCH(kj_remove_kernel, grid, 256, 0, ctx->stream, m->d, m->d_tq, range[0], range[1], m->d_qkmer,
                  m->qcount, m->alive, m->d_part, m->T);
        ctx->launches++;
        KJ_CUDA(ctx, cudaGetLastError());
    }
    return KJ_OK;
}

extern "C" int kj_wta_next(kj_match *m, kj_row *out) {
    if (!m || !out) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!m->committed) return kj_fail(ctx, KJ_E_STATE, "kj_match_commit has not run");
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    { int rc0 = ensure_host_first(m); if (rc0) return rc0; }
    memset(out, 0, sizeof(*out));
    // while (notFound && hitCounter < maxHits)   lib/kmerFinderClient.js:274
    if (m->ended || m->hit_counter >= m->max_hits) {
        m->ended = true;
        if (m->hit_counter == 0)
            return kj_fail(ctx, KJ_E_NO_WINNER, "No hits were found! (kmerResults.length === 0)");
        return 0;
    }
    auto launch_argmax = [&]() -> int {
        KJ_LAUNCH(kj_argmax_kernel, 1, 1024, 0, ctx->stream, m->d_glob, m->d_first_ord, m->d_first_idx, m->T, m->db->d_ulen,
                  (double)m->db->s_unique_lens, (double)m->db->s_templates, m->d_res);
        ctx->launches++;
        KJ_CUDA(ctx, cudaMemcpyAsync(m->h_res, m->d_res, sizeof(KjWtaResult), cudaMemcpyDeviceToHost, ctx->stream));
        return KJ_OK;
    };
    auto launch_remove = [&](uint32_t w) -> int { return wta_launch_remove(m, w); };
    int rc;
    if (!m->inflight) { rc = launch_argmax(); if (rc) return rc; }
    m->inflight = false;
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const KjWtaResult r = *m->h_res;
    // getMatches: nHits === 0 throws (lib/kmerFinderClient.js:264-266), also after earlier winners
    if (r.hits == 0 || r.winner == KJ_NONE32) {
        m->ended = true;
        return kj_fail(ctx, KJ_E_NO_HITS, "No hits were found! (nHits === 0)");
    }
    const uint32_t w = r.winner;
    // The device evaluated the gate in double precision (kj_argmax_kernel).  When its z is nowhere near
    // a fastp threshold the exact-decimal gate below cannot come out differently, so the removal (and the
    // argmax of the next round) start now and run while the host finishes the row in exact arithmetic.
    bool decisive = true;
    {
        static const double thr[27] = {10.7016, 10.4862, 10.2663, 10.0416, 9.81197, 9.5769, 9.33604, 9.08895, 8.83511,
                                       8.57394, 8.30479, 8.02686, 7.73926, 7.4409,  7.13051, 6.8065,  6.46695, 6.10941,
                                       5.73073, 5.32672, 4.89164, 4.41717, 3.89059, 3.29053, 2.57583, 1.95996, 1.64485};
        for (double t : thr) if (std::fabs(r.z - t) <= 1e-6) decisive = false;
        if (!(r.z == r.z)) decisive = false;
        if (std::fabs(r.p - 0.05) <= 1e-9) decisive = false;      // p * templates on the evalue itself: exact arithmetic decides
    }
    const bool dev_accept = r.u > 0 && r.p <= 0.05;
    bool removed = false;
    if (decisive && dev_accept) {
        rc = launch_remove(w);
        if (rc) return rc;
        removed = true;
        if (!m->distributed && m->hit_counter + 1 < m->max_hits) {
            rc = launch_argmax();
            if (rc) return rc;
            m->inflight = true;
        }
    }
    if (m->defer_rows && removed) {
        // the caller wants to issue its collective before the host spends time on the exact row
        m->pending = r;
        m->row_pending = true;
        m->hit_counter++;
        out->template_id = w;
        out->score = r.u; out->tscore = r.tau; out->hits = r.hits;
        out->z_device = r.z; out->probability_device = r.p;
        return 2;
    }
    int accepted = 0;
    if (!kj_exact_row(ctx->rounding_mode, r.u, r.tau, m->u0[w], m->t0[w], m->db->lengths[w], m->db->ulengths[w],
                      r.hits, m->kmer_map_size, m->db->s_templates, m->db->s_unique_lens, out, &accepted))
        return kj_fail(ctx, KJ_E_INVALID, "template with zero length / ulength or zero Summary.uniqueLens");
    out->template_id = w;
    out->z_device = r.z;
    out->probability_device = r.p;
    if (decisive && (accepted != 0) != dev_accept)
        return kj_fail(ctx, KJ_E_CUDA, "internal: device gate and exact-decimal gate disagree away from a threshold");
    if (!accepted) {
        // findWinner returned undefined: notFound = false (lib/kmerFinderClient.js:214-217)
        m->ended = true;
        if (m->hit_counter == 0)
            return kj_fail(ctx, KJ_E_NO_WINNER, "No hits were found! (kmerResults.length === 0)");
        return 0;
    }
    m->hit_counter++;
    if (!removed) { rc = launch_remove(w); if (rc) return rc; }
    return 1;
}

// Deferred rows (kj_match_defer_rows): kj_wta_next returned 2 after the device gate accepted the winner away
// from every threshold and the removal was launched; this finishes the row in exact arithmetic.
extern "C" int kj_match_defer_rows(kj_match *m, int on) {
    if (!m) return KJ_E_INVALID;
    m->defer_rows = on != 0;
    return KJ_OK;
}

extern "C" int kj_wta_row(kj_match *m, kj_row *out) {
    if (!m || !out) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!m->row_pending) return kj_fail(ctx, KJ_E_STATE, "kj_wta_row without a pending row");
    m->row_pending = false;
    const KjWtaResult r = m->pending;
    const uint32_t w = r.winner;
    memset(out, 0, sizeof(*out));
    int accepted = 0;
    if (!kj_exact_row(ctx->rounding_mode, r.u, r.tau, m->u0[w], m->t0[w], m->db->lengths[w], m->db->ulengths[w],
                      r.hits, m->kmer_map_size, m->db->s_templates, m->db->s_unique_lens, out, &accepted))
        return kj_fail(ctx, KJ_E_INVALID, "template with zero length / ulength or zero Summary.uniqueLens");
    out->template_id = w;
    out->z_device = r.z;
    out->probability_device = r.p;
    if (!accepted)
        return kj_fail(ctx, KJ_E_CUDA, "internal: device gate and exact-decimal gate disagree away from a threshold");
    return 1;
}

// kj_wta_all on one GPU (or on a gathered match): the rounds run in kj_wta_loop_kernel, the host waits once per
// launch and finishes the rows in exact decimal arithmetic; it steps in only for a round the double-precision
// gate could not decide.
static int wta_all_device(kj_match *m, kj_row *rows, uint32_t cap, uint32_t *n_rows, int *end_status) {
    kj_ctx *ctx = m->ctx;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t per_launch = 300;
    const size_t res_bytes = (size_t)(per_launch + 1) * sizeof(KjWtaResult) + 16;
    if (!ctx->h_wta) KJ_CUDA(ctx, cudaMallocHost(&ctx->h_wta, res_bytes));
    if (!m->d_loop) KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_loop, res_bytes));
    uint32_t *h_head = reinterpret_cast<uint32_t *>(ctx->h_wta);
    const KjWtaResult *h_res = reinterpret_cast<const KjWtaResult *>(reinterpret_cast<uint8_t *>(ctx->h_wta) + 16);
    uint32_t n = 0;
    int status = 0;
    auto exact = [&](const KjWtaResult &r, kj_row *out, int *accepted) -> int {
        const uint32_t w = r.winner;
        memset(out, 0, sizeof(*out));
        if (!kj_exact_row(ctx->rounding_mode, r.u, r.tau, r.u0, r.tau0, m->db->lengths[w], m->db->ulengths[w],
                          r.hits, m->kmer_map_size, m->db->s_templates, m->db->s_unique_lens, out, accepted))
            return kj_fail(ctx, KJ_E_INVALID, "template with zero length / ulength or zero Summary.uniqueLens");
        out->template_id = w;
        out->z_device = r.z;
        out->probability_device = r.p;
        return KJ_OK;
    };
    // the loop kernel's blocks wait for each other: all of them must be resident at once
    int loop_grid = 1;
#ifndef KJ_CPU_EMU
    {
        int per_sm = 0;
        KJ_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kj_wta_loop_kernel, 256, 0));
        loop_grid = std::max(1, std::min(ctx->sm_count, per_sm * ctx->sm_count));
    }
#endif
    if (!m->d_sync) KJ_CUDA(ctx, kj_dmalloc(ctx, &m->d_sync, (size_t)(per_launch + 2) * 4));
    for (;;) {
        // while (notFound && hitCounter < maxHits)   lib/kmerFinderClient.js:274
        if (m->hit_counter >= m->max_hits) { m->ended = true; break; }
        const uint32_t rounds = std::min<uint32_t>(per_launch, m->max_hits - m->hit_counter);
        KjWtaLoopArgs a{};
        a.d = m->d; a.glob = m->d_glob; a.glob0 = m->d_glob0; a.ford = m->d_first_ord; a.fidx = m->d_first_idx; a.ulen = m->db->d_ulen; a.toff = m->d_toff;
        a.sync = m->d_sync;
        a.tq = m->d_tq; a.qkmer = m->d_qkmer; a.qcount = m->qcount; a.alive = m->alive;
        a.T = m->T; a.max_rounds = rounds;
        a.unique_lens = (double)m->db->s_unique_lens; a.n_templates = (double)m->db->s_templates;
        a.head = reinterpret_cast<uint32_t *>(m->d_loop);
        a.res = reinterpret_cast<KjWtaResult *>(reinterpret_cast<uint8_t *>(m->d_loop) + 16);
        KJ_CUDA(ctx, cudaMemsetAsync(m->d_sync, 0, (size_t)(rounds + 2) * 4, ctx->stream));
#ifdef KJ_CPU_EMU
        KJ_LAUNCH(kj_wta_loop_kernel, 1, 256, 0, ctx->stream, a);
#else
        {
            void *kargs[1] = {(void *)&a};
            KJ_CUDA(ctx, cudaLaunchCooperativeKernel((const void *)kj_wta_loop_kernel, dim3(loop_grid), dim3(256), kargs, 0, ctx->stream));
        }
#endif
        ctx->launches++;
        KJ_CUDA(ctx, cudaGetLastError());
        KJ_CUDA(ctx, cudaMemcpyAsync(ctx->h_wta, m->d_loop, 16 + (size_t)(rounds + 1) * sizeof(KjWtaResult),
                                     cudaMemcpyDeviceToHost, ctx->stream));
        KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const uint32_t n_rec = h_head[0], st = h_head[1];
        const uint32_t n_acc = st == KJ_LOOP_MORE ? n_rec : n_rec - 1;
        for (uint32_t i = 0; i < n_acc; ++i) {
            int accepted = 0;
            int rc = exact(h_res[i], &rows[n], &accepted);
            if (rc) return rc;
            if (!accepted)
                return kj_fail(ctx, KJ_E_CUDA, "internal: device gate and exact-decimal gate disagree away from a threshold");
            ++n;
            m->hit_counter++;
        }
        if (st == KJ_LOOP_MORE) continue;
        const KjWtaResult &last = h_res[n_rec - 1];
        if (st == KJ_LOOP_NO_HITS) {
            // getMatches: nHits === 0 throws (lib/kmerFinderClient.js:264-266), also after earlier winners
            m->ended = true;
            status = kj_fail(ctx, KJ_E_NO_HITS, "No hits were found! (nHits === 0)");
            break;
        }
        kj_row tmp;
        int accepted = 0;
        int rc = exact(last, &tmp, &accepted);
        if (rc) return rc;
        if (st == KJ_LOOP_REJECTED && accepted)
            return kj_fail(ctx, KJ_E_CUDA, "internal: device gate and exact-decimal gate disagree away from a threshold");
        if (accepted) {                         // undecided on the device, accepted by the exact arithmetic
            rows[n++] = tmp;
            m->hit_counter++;
            rc = wta_launch_remove(m, last.winner);      // (fetches the host copy of the list offsets: a rare path)
            if (rc) return rc;
            continue;
        }
        // findWinner returned undefined: notFound = false (lib/kmerFinderClient.js:214-217)
        m->ended = true;
        if (m->hit_counter == 0) status = kj_fail(ctx, KJ_E_NO_WINNER, "No hits were found! (kmerResults.length === 0)");
        break;
    }
    (void)cap;
    *n_rows = n;
    *end_status = status;
    return KJ_OK;
}

// The whole findMatches loop in one call: kj_wta_next until it ends, without a trip through the host
// language per row.  Each round's exact-decimal row is finished while the device removes the winner's k-mers
// and takes the next argmax (kj_wta_next launches both before it starts the arithmetic).  rows[0 .. *n_rows)
// are the rows the generator would have yielded; *end_status is 0 when it would have returned normally, or
// the KJ_E_NO_HITS / KJ_E_NO_WINNER it would have thrown after them (text in kj_last_error).
// (Helper threads for the rows were tried and measured slower: starting them costs more than the ~20 us per
// round the dependent chain argmax -> copy -> host -> removal takes anyway.)
extern "C" int kj_wta_all(kj_match *m, kj_row *rows, uint32_t cap, uint32_t *n_rows, int *end_status) {
    if (!m || !rows || !n_rows || !end_status) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!m->committed) return kj_fail(ctx, KJ_E_STATE, "kj_match_commit has not run");
    if (cap < m->max_hits) return kj_fail(ctx, KJ_E_RANGE, "row buffer smaller than maxHits");
    if (!m->distributed && !m->inflight && !m->row_pending && !m->ended)
        return wta_all_device(m, rows, cap, n_rows, end_status);
    const bool was_deferring = m->defer_rows;
    m->defer_rows = false;
    uint32_t n = 0;
    int status = 0, hard = KJ_OK;
    for (;;) {
        kj_row tmp;
        const int rc = kj_wta_next(m, &tmp);
        if (rc == 1 && n < cap) { rows[n++] = tmp; continue; }
        if (rc == KJ_E_NO_HITS || rc == KJ_E_NO_WINNER) status = rc;
        else if (rc != 0) hard = rc < 0 ? rc : KJ_E_STATE;
        break;
    }
    m->defer_rows = was_deferring;
    if (hard != KJ_OK) return hard;
    *n_rows = n;
    *end_status = status;
    return KJ_OK;
}

// standardScoring (lib/kmerFinderServer.js:857-874): matchSummary of every matched template against
// the first-match result, sorted by score (descending, stable); templates the evalue gate rejects
// produce no row.
extern "C" int kj_standard_scoring(kj_match *m, kj_row *rows, uint32_t n_rows_cap, uint32_t *n_rows) {
    if (!m || !n_rows) return KJ_E_INVALID;
    kj_ctx *ctx = m->ctx;
    if (!m->committed) return kj_fail(ctx, KJ_E_STATE, "kj_match_commit has not run");
    { std::lock_guard<std::recursive_mutex> lk(ctx->mu); int rc0 = ensure_host_first(m); if (rc0) return rc0; }
    std::vector<kj_row> all;
    for (uint32_t t : m->order) {
        kj_row r;
        memset(&r, 0, sizeof(r));
        int accepted = 0;
        if (!kj_exact_row(ctx->rounding_mode, m->u0[t], m->t0[t], m->u0[t], m->t0[t], m->db->lengths[t],
                          m->db->ulengths[t], m->hits0, m->kmer_map_size, m->db->s_templates,
                          m->db->s_unique_lens, &r, &accepted))
            return kj_fail(ctx, KJ_E_INVALID, "template with zero length / ulength or zero Summary.uniqueLens");
        if (!accepted) continue;
        r.template_id = t;
        all.push_back(r);
    }
    std::stable_sort(all.begin(), all.end(), [](const kj_row &a, const kj_row &b) { return a.score > b.score; });
    *n_rows = (uint32_t)all.size();
    if (rows) {
        if (all.size() > n_rows_cap) return kj_fail(ctx, KJ_E_RANGE, "row buffer too small");
        if (!all.empty()) memcpy(rows, all.data(), all.size() * sizeof(kj_row));
    }
    return KJ_OK;
}

extern "C" int kj_stats_zscore_device(kj_ctx *ctx, uint64_t n, const uint64_t *r1, const uint64_t *n1,
                                      const uint64_t *r2, const uint64_t *n2, double *z, double *p) {
    if (!ctx || (n && (!r1 || !n1 || !r2 || !n2 || !z || !p))) return kj_fail(ctx, KJ_E_INVALID, "null argument");
    if (!n) return KJ_OK;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t *d_in = nullptr;
    double *d_out = nullptr;
    KJ_CUDA(ctx, kj_dmalloc(ctx, &d_in, 4 * n * 8));
    cudaError_t e = kj_dmalloc(ctx, &d_out, 2 * n * 8);
    const uint64_t *src[4] = {r1, n1, r2, n2};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
        e = cudaMemcpyAsync(d_in + i * n, src[i], n * 8, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        KJ_LAUNCH(kj_stats_kernel, kj_grid_for(ctx, n), 256, 0, ctx->stream, n, d_in, d_in + n, d_in + 2 * n,
                  d_in + 3 * n, d_out, d_out + n);
        ctx->launches++;
        e = cudaMemcpyAsync(z, d_out, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(p, d_out + n, n * 8, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    kj_dfree(ctx, d_in); kj_dfree(ctx, d_out);
    if (e != cudaSuccess) return kj_fail(ctx, KJ_E_CUDA, cudaGetErrorString(e));
    return KJ_OK;
}
