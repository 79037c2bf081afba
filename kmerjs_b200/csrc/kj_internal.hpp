// kj_internal.hpp -- host-side structures behind the opaque handles of include/kmerjs_b200.h
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <map>
#include <vector>
#include "../../include/kmerjs_b200.h"
#include "kj_device.cuh"

// Kernel launch.  The second definition belongs to tools/cuemu (a developer-only harness that
// compiles these sources with g++ to run them under AddressSanitizer on a box without a GPU);
// the shipped library is always built by nvcc and takes the first one.
#ifndef KJ_CPU_EMU
#define KJ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define KJ_DYN_SMEM(name) extern __shared__ __align__(16) uint8_t name[]
#endif

struct kj_ctx {
    std::map<const void *, int> occ_cache;   // kernel -> resident CTAs per SM (asked once)
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;      // launch stream
    bool own_stream = false;
    cudaStream_t copy_stream = nullptr; // H2D staging
    std::recursive_mutex mu;
    std::string err;
    uint64_t launches = 0;
    int rounding_mode = 4;              // bignumber.js ROUND_HALF_UP
    // scan-kernel timing (CUDA events on `stream`)
    bool timers_on = false;
    uint64_t stage_chunk = 0;           // host -> device staging chunk (0: KJ_STAGE_CHUNK_MB or 64 MiB at first use)
    double scan_ms = 0.0;
    uint64_t scan_launches = 0;
    uint64_t scan_bytes = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    double verify_ms = 0.0;            // part of scan_ms spent in kj_verify_kernel
    // host -> device staging (KJ_MEM_HOST buffers), double buffered, owned by the context so that
    // repeated count jobs do not re-allocate
    uint8_t *d_stage[2] = {nullptr, nullptr};
    uint8_t *h_stage[2] = {nullptr, nullptr};
    uint64_t stage_cap = 0;
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
    // small pinned blocks (device counters mirrored to the host, WTA results): cudaMallocHost costs
    // milliseconds, so handles borrow 1024-byte blocks from a slab owned by the context
    void *h_wta = nullptr;              // pinned: the records of one kj_wta_loop_kernel launch
    uint8_t *pin_slab = nullptr;
    std::vector<void *> pin_free;
};

void *kj_pinned_get(kj_ctx *ctx);              // 1024 bytes, nullptr when out of memory
void kj_pinned_put(kj_ctx *ctx, void *p);

// stream-ordered device memory from the default pool (release threshold = keep everything), so
// that per-job tables are recycled without cudaMalloc/cudaFree round trips
template <class T>
static inline cudaError_t kj_dmalloc(kj_ctx *ctx, T **p, uint64_t bytes) {
    return cudaMallocAsync((void **)p, bytes ? bytes : 1, ctx->stream);
}
static inline void kj_dfree(kj_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

extern thread_local std::string kj_tls_error;

int kj_fail(kj_ctx *ctx, int code, const std::string &msg);

#define KJ_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t e__ = (call);                                                       \
        if (e__ != cudaSuccess)                                                         \
            return kj_fail((ctx), KJ_E_CUDA,                                            \
                           std::string(#call) + ": " + cudaGetErrorString(e__));       \
    } while (0)

struct KjCompact {        // built by kj_counts_finish: dense list of the regular entries
    uint64_t *keys = nullptr;
    uint64_t *counts = nullptr;
    uint64_t *ords = nullptr;
    uint8_t *alive = nullptr;   // WTA alive mask (1 = still in the query map)
    uint64_t n = 0;             // all entries: [0,n_tab) table entries, then the special key (k = 32,
    uint64_t n_reg = 0;         // all G) if present -> n_reg, then the irregular entries (byte keys on
    uint64_t n_tab = 0;         // the host, irr_host, same order)
};

struct kj_counts {
    kj_ctx *ctx = nullptr;
    // parameters
    std::vector<uint8_t> prefix, rprefix;
    uint32_t k = 16, step = 1, flags = 0;
    bool order = true;
    bool use_filter = true;      // filter kernel (step == 1, 1 <= m <= k), else dense or line kernel
    bool use_dense = false;      // dense kernel (step == 1, empty prefix, k >= 2): every window is an emission
    uint64_t capacity_hint = 0;
    // stream position
    uint64_t voff = 0;           // virtual offset of the next byte (starts at base_col)
    uint64_t base_line = 0, base_col = 0;
    uint64_t consumed = 0;       // stream bytes consumed
    uint32_t parity = 0;
    bool finished = false;
    bool saw_final = false;
    // device state
    KjTable tab{};
    uint64_t cap = 0;
    KjIrrTable irr{};
    uint64_t irr_cap = 0;
    std::vector<uint64_t> tail_counts, tail_ords;   // host side of the special + irregular entries (kj_counts_finish)
    uint64_t irr_bound = 0;          // upper bound of the side table's entries: last pull + merged since
    KjOverflow ovf{};
    KjCounters *ctr = nullptr;        // device
    KjCounters *h_ctr = nullptr;      // pinned mirror
    uint64_t *tile_mem = nullptr;     // tile_cap u64: look-back status words (dense / line kernels), tile_excl (filter path)
    uint64_t tile_cap = 0;
    uint64_t *tile_cnt = nullptr;     // filter path: '\n' per tile
    void *scan_tmp = nullptr;         // filter path: temporary storage of the exclusive scan over the tiles
    size_t scan_tmp_bytes = 0;
    uint64_t *cand = nullptr;         // candidate entries of the piece in flight (filter path), 16 bytes each
    uint64_t cand_cap = 0;
    struct KjPiece *piece = nullptr;  // filter path: the piece in flight (kernel arguments + tensor map), kj_count.cu
    bool pending = false;             // the piece has been launched and not settled yet
    bool exchange_totals = false;     // totals come from the segment headers of a fixed-capacity exchange
    // results
    KjCompact reg{};
    // irregular entries, host side after finish: 56-byte records
    std::vector<uint8_t> irr_host;   // {u8 key[32], u64 len, u64 count, u64 ord} * n
    std::vector<uint64_t> export_perm;   // export position -> query index (built on first export)
    std::vector<uint64_t> export_rank;   // its inverse
    uint64_t bytes_read = 0;
    uint64_t lines = 0;
    uint64_t occurrences = 0;
    uint64_t bases = 0;
    uint64_t special_count = 0, special_ord = 0;
    // partition scratch
    uint64_t *part_rec = nullptr;
    uint64_t part_cap = 0;
};

// the template database in HBM (kj_score.cu) and what the loaders keep beside it (kj_dbio.cu)
#include <unordered_map>
#define KJ_NONE32 0xFFFFFFFFu
struct KjDbDev {
    const uint64_t *keys;
    const uint32_t *vals;
    uint64_t mask;
    const uint64_t *list_off;
    const uint32_t *tmpl;
};

struct kj_db {
    kj_ctx *ctx = nullptr;
    uint32_t k = 0;                 // length of the regular (ACGT-only) k-mers in the device index
    uint64_t n_kmers = 0, n_pairs = 0;
    uint32_t n_templates = 0;
    uint64_t *d_keys = nullptr;
    uint32_t *d_vals = nullptr;
    uint64_t cap = 0;
    uint64_t *d_list_off = nullptr;
    uint32_t *d_tmpl = nullptr;
    uint64_t *d_ulen = nullptr;
    std::vector<uint64_t> lengths, ulengths;
    std::unordered_map<std::string, uint32_t> other;   // k-mers that are not regular: bytes -> k-mer id
    uint32_t special_id = KJ_NONE32;                   // k-mer whose key equals KJ_EMPTY (k = 32, all G)
    uint64_t s_templates = 0, s_unique_lens = 0, s_total_len = 0;
    std::vector<std::string> names, species;           // per template, when the DB came through kj_db_load (else empty)
    KjDbDev dev() const { return KjDbDev{d_keys, d_vals, cap - 1, d_list_off, d_tmpl}; }
};


// records used by the exchange and by export: {key, count, ord}
struct KjRecord { uint64_t key, count, ord; };
struct KjIrrRecord { uint8_t key[32]; uint64_t len, count, ord; };

// kj_count.cu internals used by kj_score.cu
int kj_counts_check_finished(const kj_counts *c);
int kj_grid_for(const kj_ctx *ctx, uint64_t n, int threads = 256);
void kj_decode_key(uint64_t key, uint32_t k, uint8_t *out);
int kj_counts_export_rank(kj_counts *c, const std::vector<uint64_t> **rank);   // query index -> export position
