// tools/cuemu/emu.cpp -- DEVELOPER TOOL (see cuda_runtime.h in this directory): fiber-based
// emulation of one CUDA block at a time on one OS thread, plus a malloc-backed "device".
#include "cuda_runtime.h"
#include <ucontext.h>
#include <stdio.h>
#include <chrono>
#include <map>
#include <mutex>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
extern "C" void __sanitizer_start_switch_fiber(void **fake, const void *bottom, size_t size);
extern "C" void __sanitizer_finish_switch_fiber(void *fake, const void **bottom_old, size_t *size_old);
#define EMU_ASAN 1
#else
#define EMU_ASAN 0
#endif

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace {
constexpr size_t kStack = 256 * 1024;
struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
    uint3 tid{};
    void *fake = nullptr;
};
struct Warp {
    uint64_t buf[32];
    int count = 0;
    unsigned gen = 0;
};
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
ucontext_t g_sched;
void *g_sched_fake = nullptr;
const void *g_sched_bottom = nullptr;
size_t g_sched_size = 0;
int g_cur = -1;
int g_block_count = 0;
unsigned g_block_gen = 0;
int g_nthreads = 0;
const std::function<void()> *g_body = nullptr;
long g_idle_spins = 0;
long g_switches = 0;
std::recursive_mutex g_launch_mu;

void fiber_main() {
#if EMU_ASAN
    __sanitizer_finish_switch_fiber(nullptr, &g_sched_bottom, &g_sched_size);
#endif
    (*g_body)();
    g_fibers[g_cur].done = true;
#if EMU_ASAN
    __sanitizer_start_switch_fiber(nullptr, g_sched_bottom, g_sched_size);
#endif
    swapcontext(&g_fibers[g_cur].ctx, &g_sched);
}
}  // namespace

void emu_yield() {
    Fiber &f = g_fibers[g_cur];
    int me = g_cur;
#if EMU_ASAN
    __sanitizer_start_switch_fiber(&f.fake, g_sched_bottom, g_sched_size);
#endif
    swapcontext(&f.ctx, &g_sched);
#if EMU_ASAN
    __sanitizer_finish_switch_fiber(g_fibers[me].fake, &g_sched_bottom, &g_sched_size);
#endif
    (void)me;
}

static void spin_guard() {
    static const long limit = getenv("EMU_SPIN_LIMIT") ? atol(getenv("EMU_SPIN_LIMIT")) : 200000000L;
    if (++g_idle_spins > limit) {
        fprintf(stderr, "cuemu: deadlock suspected (thread %u of block %u)\n", threadIdx.x, blockIdx.x);
        abort();
    }
}

void emu_syncthreads() {
    if (++g_block_count == g_nthreads) {
        g_block_count = 0;
        ++g_block_gen;
        g_idle_spins = 0;
        return;
    }
    unsigned gen = g_block_gen;
    while (g_block_gen == gen) { spin_guard(); emu_yield(); }
}

namespace { struct NamedBar { int count = 0; unsigned gen = 0; }; NamedBar g_named[16]; }
void emu_named_barrier(int id, int expected) {
    NamedBar &b = g_named[id & 15];
    if (++b.count == expected) {
        b.count = 0;
        ++b.gen;
        g_idle_spins = 0;
        return;
    }
    unsigned gen = b.gen;
    while (b.gen == gen) { spin_guard(); emu_yield(); }
}

void emu_named_arrive(int id, int expected) {
    NamedBar &b = g_named[id & 15];
    if (++b.count == expected) {
        b.count = 0;
        ++b.gen;
        g_idle_spins = 0;
    }
}

unsigned emu_lane() { return threadIdx.x & 31u; }

static int warp_size_of(unsigned w) {
    int lo = (int)w * 32;
    int n = g_nthreads - lo;
    return n > 32 ? 32 : n;
}

static void warp_bar(Warp &w, int expected) {
    if (++w.count == expected) {
        w.count = 0;
        ++w.gen;
        g_idle_spins = 0;
        return;
    }
    unsigned gen = w.gen;
    while (w.gen == gen) { spin_guard(); emu_yield(); }
}

uint64_t emu_warp_xchg(uint64_t v, int src) {
    unsigned wi = threadIdx.x >> 5;
    Warp &w = g_warps[wi];
    int n = warp_size_of(wi);
    w.buf[threadIdx.x & 31] = v;
    warp_bar(w, n);
    uint64_t r = (src < n) ? w.buf[src] : v;
    warp_bar(w, n);
    return r;
}
uint32_t emu_warp_ballot(int pred) {
    unsigned wi = threadIdx.x >> 5;
    Warp &w = g_warps[wi];
    int n = warp_size_of(wi);
    w.buf[threadIdx.x & 31] = pred ? 1 : 0;
    warp_bar(w, n);
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) if (w.buf[i]) r |= 1u << i;
    warp_bar(w, n);
    return r;
}
uint64_t emu_warp_reduce_add(uint64_t v) {
    unsigned wi = threadIdx.x >> 5;
    Warp &w = g_warps[wi];
    int n = warp_size_of(wi);
    w.buf[threadIdx.x & 31] = v;
    warp_bar(w, n);
    uint64_t r = 0;
    for (int i = 0; i < n; ++i) r += w.buf[i];
    warp_bar(w, n);
    return r;
}
uint64_t emu_warp_reduce_op(uint64_t v, int op) {
    unsigned wi = threadIdx.x >> 5;
    Warp &w = g_warps[wi];
    int n = warp_size_of(wi);
    w.buf[threadIdx.x & 31] = v;
    warp_bar(w, n);
    uint64_t r = w.buf[0];
    for (int i = 1; i < n; ++i) {
        uint64_t x = w.buf[i];
        switch (op) { case 0: if (x < r) r = x; break; case 1: if (x > r) r = x; break;
                      case 2: r |= x; break; default: r &= x; break; }
    }
    warp_bar(w, n);
    return r;
}
uint32_t emu_match_any(uint64_t v) {
    unsigned wi = threadIdx.x >> 5;
    Warp &w = g_warps[wi];
    int n = warp_size_of(wi);
    w.buf[threadIdx.x & 31] = v;
    warp_bar(w, n);
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) if (w.buf[i] == v) r |= 1u << i;
    warp_bar(w, n);
    return r;
}

static std::vector<uint8_t> g_dyn;
uint8_t *emu_dyn_smem() { return g_dyn.data(); }

void emu_launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body) {
    std::lock_guard<std::recursive_mutex> lk(g_launch_mu);
    g_dyn.assign(dyn_smem + 16, 0xCD);
    g_switches = 0;
    const int nthreads = (int)(block.x * block.y * block.z);
    if ((int)g_fibers.size() < nthreads) {
        size_t old = g_fibers.size();
        g_fibers.resize(nthreads);
        for (size_t i = old; i < g_fibers.size(); ++i) g_fibers[i].stack = (char *)malloc(kStack);
    }
    g_warps.assign((nthreads + 31) / 32, Warp());
    gridDim = grid;
    blockDim = block;
    g_nthreads = nthreads;
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
        g_block_count = 0;
        for (auto &w : g_warps) { w.count = 0; }
        for (auto &nb : g_named) { nb.count = 0; }
        for (int t = 0; t < nthreads; ++t) {
            Fiber &f = g_fibers[t];
            f.done = false;
            f.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, fiber_main, 0);
        }
        int live = nthreads;
        while (live) {
            for (int t = 0; t < nthreads; ++t) {
                Fiber &f = g_fibers[t];
                if (f.done) continue;
                g_cur = t;
                threadIdx = f.tid;
                blockIdx = uint3{bx, by, bz};
#if EMU_ASAN
                __sanitizer_start_switch_fiber(&g_sched_fake, f.stack, kStack);
#endif
                ++g_switches;
                swapcontext(&g_sched, &f.ctx);
#if EMU_ASAN
                __sanitizer_finish_switch_fiber(g_sched_fake, nullptr, nullptr);
#endif
                if (f.done) --live;
            }
        }
    }
    g_body = nullptr;
    if (getenv("EMU_STATS")) fprintf(stderr, "cuemu: launch grid=%u block=%u switches=%ld\n", grid.x, block.x, g_switches);
}

// ---- "device" memory -----------------------------------------------------------------------
namespace {
std::map<const void *, std::pair<size_t, int>> g_allocs;   // base -> (size, kind 1 host-pinned / 2 device)
std::mutex g_alloc_mu;
cudaError_t alloc(void **p, size_t n, int kind) {
    void *q = nullptr;
    if (posix_memalign(&q, 256, n ? n : 1)) return cudaErrorMemoryAllocation;
    if (n <= (16u << 20)) memset(q, 0xCD, n);   // poison: uninitialised reads show up as garbage deterministically
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_allocs[q] = {n, kind};
    *p = q;
    return cudaSuccess;
}
cudaError_t release(void *p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        auto it = g_allocs.find(p);
        if (it == g_allocs.end()) { fprintf(stderr, "cuemu: free of unknown pointer %p\n", p); abort(); }
        g_allocs.erase(it);
    }
    free(p);
    return cudaSuccess;
}
}  // namespace

cudaError_t cudaMalloc(void **p, size_t n) { return alloc(p, n, 2); }
cudaError_t cudaFree(void *p) { return release(p); }
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return alloc(p, n, 2); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return release(p); }
cudaError_t cudaMallocHost(void **p, size_t n) { return alloc(p, n, 1); }
cudaError_t cudaFreeHost(void *p) { return release(p); }
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = (cudaStream_t)malloc(8); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
struct EmuEvent { double t; };
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new EmuEvent{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof(*p));
    strcpy(p->name, "cuemu (CPU emulation, developer tool)");
    p->major = 10; p->minor = 0; p->multiProcessorCount = 2; p->totalGlobalMem = 8ull << 30;
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e ? "cuemu error" : "no error"; }
cudaError_t cudaMemGetInfo(size_t *f, size_t *t) { *f = 6ull << 30; *t = 8ull << 30; return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    a->type = cudaMemoryTypeUnregistered; a->device = 0; a->devicePointer = nullptr; a->hostPointer = nullptr;
    auto it = g_allocs.upper_bound(p);
    if (it != g_allocs.begin()) {
        --it;
        if ((const char *)p < (const char *)it->first + it->second.first)
            a->type = it->second.second == 1 ? cudaMemoryTypeHost : cudaMemoryTypeDevice;
    }
    return cudaSuccess;
}
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int) { *pool = nullptr; return cudaSuccess; }
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
