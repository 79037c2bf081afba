// kj_ctx.cu -- context, error text, timers.
#include "kj_internal.hpp"

thread_local std::string kj_tls_error;

int kj_fail(kj_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg;
    kj_tls_error = msg;
    return code;
}

static const size_t KJ_PIN_BLOCK = 1024, KJ_PIN_BLOCKS = 256;

void *kj_pinned_get(kj_ctx *ctx) {
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    if (!ctx->pin_slab) {
        if (cudaMallocHost(&ctx->pin_slab, KJ_PIN_BLOCK * KJ_PIN_BLOCKS) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        for (size_t i = 0; i < KJ_PIN_BLOCKS; ++i) ctx->pin_free.push_back(ctx->pin_slab + i * KJ_PIN_BLOCK);
    }
    if (ctx->pin_free.empty()) {          // more live handles than blocks: a block of its own
        void *p = nullptr;
        if (cudaMallocHost(&p, KJ_PIN_BLOCK) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        return p;
    }
    void *p = ctx->pin_free.back();
    ctx->pin_free.pop_back();
    return p;
}

void kj_pinned_put(kj_ctx *ctx, void *p) {
    if (!p) return;
    std::lock_guard<std::recursive_mutex> lk(ctx->mu);
    uint8_t *q = (uint8_t *)p;
    if (ctx->pin_slab && q >= ctx->pin_slab && q < ctx->pin_slab + KJ_PIN_BLOCK * KJ_PIN_BLOCKS) ctx->pin_free.push_back(p);
    else cudaFreeHost(p);
}

extern "C" int kj_abi_version(void) { return KJ_ABI_VERSION; }

extern "C" int kj_init(int device, void *stream, kj_ctx **out) {
    if (!out) return kj_fail(nullptr, KJ_E_INVALID, "kj_init: out is null");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        return kj_fail(nullptr, KJ_E_NO_SM100,
                       std::string("no CUDA device (") + cudaGetErrorString(e) +
                           "): kmerjs_b200 has no CPU fallback");
    if (device < 0 || device >= n_dev) return kj_fail(nullptr, KJ_E_INVALID, "kj_init: bad device index");
    cudaDeviceProp prop{};
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return kj_fail(nullptr, KJ_E_CUDA, cudaGetErrorString(e));
    if (prop.major != 10)
        return kj_fail(nullptr, KJ_E_NO_SM100,
                       std::string("device ") + prop.name + " is sm_" + std::to_string(prop.major) +
                           std::to_string(prop.minor) +
                           "; this library carries sm_100a code only and has no CPU fallback");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return kj_fail(nullptr, KJ_E_CUDA, cudaGetErrorString(e));
    kj_ctx *ctx = new kj_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        ctx->own_stream = true;
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        // keep freed blocks in the stream-ordered pool: per-job tables are recycled, not returned
        cudaMemPool_t pool = nullptr;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess && pool) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev2);
    if (e != cudaSuccess) {
        int rc = kj_fail(nullptr, KJ_E_CUDA, cudaGetErrorString(e));
        kj_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return KJ_OK;
}

extern "C" void kj_destroy(kj_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 2; ++i) {
        cudaFree(ctx->d_stage[i]);
        if (ctx->h_stage[i]) cudaFreeHost(ctx->h_stage[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
    }
    if (ctx->pin_slab) cudaFreeHost(ctx->pin_slab);
    if (ctx->h_wta) cudaFreeHost(ctx->h_wta);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *kj_last_error(const kj_ctx *ctx) {
    return ctx ? ctx->err.c_str() : kj_tls_error.c_str();
}

extern "C" uint64_t kj_launch_count(const kj_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" double kj_scan_kernel_ms(const kj_ctx *ctx, uint64_t *n_launches) {
    if (!ctx) return 0.0;
    if (n_launches) *n_launches = ctx->scan_launches;
    return ctx->scan_launches ? ctx->scan_ms / (double)ctx->scan_launches : 0.0;
}
extern "C" double kj_verify_kernel_ms(const kj_ctx *ctx) {
    return ctx && ctx->scan_launches ? ctx->verify_ms / (double)ctx->scan_launches : 0.0;
}
extern "C" uint64_t kj_scan_kernel_bytes(const kj_ctx *ctx) { return ctx ? ctx->scan_bytes : 0; }
extern "C" void kj_reset_timers(kj_ctx *ctx) {
    if (ctx) { ctx->scan_ms = 0.0; ctx->verify_ms = 0.0; ctx->scan_launches = 0; ctx->scan_bytes = 0; }
}
extern "C" void kj_enable_timers(kj_ctx *ctx, int on) {
    if (ctx) ctx->timers_on = on != 0;
}

extern "C" int kj_set_rounding_mode(kj_ctx *ctx, int mode) {
    if (!ctx) return KJ_E_INVALID;
    if (mode < 0 || mode > 6) return kj_fail(ctx, KJ_E_INVALID, "rounding mode must be 0..6");
    ctx->rounding_mode = mode;
    return KJ_OK;
}
