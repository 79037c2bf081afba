// kj_synth.cu -- deterministic Illumina-shaped synthetic FASTQ, generated on the device
// (SURVEY.md 8d).  Bench/test utility: not part of the reference path.  Every record has the same
// size, so any rank can generate any range of reads independently (counter-based hashing):
//
//   @SIM:1:FC:1:TTTT:XXXXX:YYYYY 1:N:0:CGATGT\n      42 bytes
//   <read_len bases>\n
//   +\n
//   <read_len qualities, '#'..'I', '@' and '+' included>\n
#include "kj_internal.hpp"

#define KJ_SYNTH_HDR 42

__device__ __forceinline__ uint64_t kj_rng(uint64_t seed, uint64_t read, uint64_t j) {
    return kj_mix64(kj_mix64(seed ^ (read * 0x9E3779B97F4A7C15ull)) + j * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull);
}

struct KjSynthArgs {
    uint64_t seed, n_reads, first_read;
    uint32_t read_len;
    const uint8_t *genome;
    uint64_t genome_len;
    uint32_t sub_thr, n_thr, lead_thr;   // probabilities scaled to 2^24
    uint8_t *out;
};

__global__ void kj_synth_kernel(const KjSynthArgs a) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t L = a.read_len;
    const uint64_t rec = (uint64_t)KJ_SYNTH_HDR + 2ull * L + 4ull;
    const char comp[4] = {'T', 'G', 'A', 'C'};       // complement by code (A C T G) -> T G A C
    const char base_of[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < a.n_reads; i += warps) {
        const uint64_t r = a.first_read + i;
        uint8_t *o = a.out + i * rec;
        const uint64_t h0 = kj_rng(a.seed, r, 0xFFFFFFF0ull);
        const uint64_t h1 = kj_rng(a.seed, r, 0xFFFFFFF1ull);
        const uint64_t start = h0 % (a.genome_len - L + 1);
        const uint32_t strand = (uint32_t)(h1 & 1);
        const bool lead_n = ((h1 >> 8) & 0xFFFFFF) < a.lead_thr;
        // header
        const uint32_t tile = 1101 + (uint32_t)((h1 >> 32) % 1000);
        const uint32_t x = (uint32_t)((h0 >> 20) % 100000), y = (uint32_t)((h0 >> 40) % 100000);
        for (uint32_t j = lane; j < KJ_SYNTH_HDR; j += 32) {
            const char *fix = "@SIM:1:FC:1:";
            const char *tail = " 1:N:0:CGATGT\n";
            char c;
            if (j < 12) c = fix[j];
            else if (j < 16) { uint32_t d = 15 - j, v = tile; while (d--) v /= 10; c = (char)('0' + v % 10); }
            else if (j == 16) c = ':';
            else if (j < 22) { uint32_t d = 21 - j, v = x; while (d--) v /= 10; c = (char)('0' + v % 10); }
            else if (j == 22) c = ':';
            else if (j < 28) { uint32_t d = 27 - j, v = y; while (d--) v /= 10; c = (char)('0' + v % 10); }
            else c = tail[j - 28];
            o[j] = (uint8_t)c;
        }
        // bases
        uint8_t *b = o + KJ_SYNTH_HDR;
        for (uint32_t j = lane; j < L; j += 32) {
            uint8_t g = strand ? a.genome[start + (L - 1 - j)] : a.genome[start + j];
            char c = strand ? comp[(g >> 1) & 3] : (char)g;
            const uint64_t h = kj_rng(a.seed, r, j);
            if ((h & 0xFFFFFF) < a.sub_thr) {
                // substitute by one of the three other bases
                uint32_t cur = (c == 'A') ? 0 : (c == 'C') ? 1 : (c == 'G') ? 2 : 3;
                c = base_of[(cur + 1 + (uint32_t)((h >> 24) % 3)) & 3];
            }
            if (((h >> 32) & 0xFFFFFF) < a.n_thr) c = 'N';
            if (j == 0 && lead_n) c = 'N';
            b[j] = (uint8_t)c;
        }
        if (lane == 0) { b[L] = '\n'; b[L + 1] = '+'; b[L + 2] = '\n'; }
        uint8_t *ql = b + L + 3;
        for (uint32_t j = lane; j < L; j += 32)
            ql[j] = (uint8_t)('#' + (uint32_t)(kj_rng(a.seed, r, 0x10000ull + j) % 39));
        if (lane == 0) ql[L] = '\n';
    }
}

__global__ void kj_genome_kernel(uint64_t seed, uint8_t *out, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const char base_of[4] = {'A', 'C', 'G', 'T'};
        out[i] = (uint8_t)base_of[kj_rng(seed, 0x67656E6F6D65ull, i) >> 62];
    }
}

extern "C" int kj_synth_size(kj_ctx *ctx, const kj_synth_params *p, uint64_t *n_bytes) {
    if (!p || !n_bytes) return kj_fail(ctx, KJ_E_INVALID, "kj_synth_size: null argument");
    *n_bytes = p->n_reads * ((uint64_t)KJ_SYNTH_HDR + 2ull * p->read_len + 4ull);
    return KJ_OK;
}

extern "C" int kj_synth_generate(kj_ctx *ctx, const kj_synth_params *p, uint8_t *dev_out, uint64_t n_bytes) {
    if (!ctx || !p || !dev_out) return kj_fail(ctx, KJ_E_INVALID, "kj_synth_generate: null argument");
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    uint64_t need = 0;
    kj_synth_size(ctx, p, &need);
    if (n_bytes < need) return kj_fail(ctx, KJ_E_INVALID, "kj_synth_generate: output buffer too small");
    if (!p->genome || p->genome_len < p->read_len || p->read_len == 0)
        return kj_fail(ctx, KJ_E_INVALID, "kj_synth_generate: genome shorter than a read");
    if (!p->n_reads) return KJ_OK;
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    KjSynthArgs a{};
    a.seed = p->seed; a.n_reads = p->n_reads; a.first_read = p->first_read; a.read_len = p->read_len;
    a.genome = p->genome; a.genome_len = p->genome_len;
    auto thr = [](double pr) { double v = pr * 16777216.0; return (uint32_t)(v < 0 ? 0 : v > 16777216.0 ? 16777216.0 : v); };
    a.sub_thr = thr(p->sub_rate); a.n_thr = thr(p->n_rate); a.lead_thr = thr(p->lead_n_rate);
    a.out = dev_out;
    const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((p->n_reads + 7) / 8, (uint64_t)ctx->sm_count * 16));
    KJ_LAUNCH(kj_synth_kernel, grid, 256, 0, ctx->stream, a);
    ctx->launches++;
    KJ_CUDA(ctx, cudaGetLastError());
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KJ_OK;
}

extern "C" int kj_synth_genome(kj_ctx *ctx, uint64_t seed, uint8_t *dev_out, uint64_t n) {
    if (!ctx || (n && !dev_out)) return kj_fail(ctx, KJ_E_INVALID, "kj_synth_genome: null argument");
    if (!n) return KJ_OK;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    KJ_CUDA(ctx, cudaSetDevice(ctx->device));
    KJ_LAUNCH(kj_genome_kernel, kj_grid_for(ctx, n), 256, 0, ctx->stream, seed, dev_out, n);
    ctx->launches++;
    KJ_CUDA(ctx, cudaGetLastError());
    KJ_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KJ_OK;
}
