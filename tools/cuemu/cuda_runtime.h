// tools/cuemu/cuda_runtime.h -- DEVELOPER TOOL, NOT PRODUCT CODE, NEVER SHIPPED OR LOADED BY THE PACKAGE.
//
// A minimal single-OS-thread emulation of the CUDA execution model (fibers for the threads of a
// block, blocks run one after another) so that the kernels under kmerjs_b200/csrc can be compiled
// with g++ and run under AddressSanitizer/UBSan on a box that has no GPU.  It exists to find
// out-of-bounds accesses and logic errors BEFORE GPU time is spent; it is not a fallback: the
// shipped libkmerjs_b200.so is built by nvcc for sm_100a only (kmerjs_b200/build.py) and fails
// with KJ_E_NO_SM100 without a Blackwell device.  Nothing in tests/, bench.py or the package
// builds or loads the emulated library; tools/cuemu/run.sh does, by hand.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <unordered_map>
#include <vector>
#include <cmath>
#include <cstdio>

#define KJ_CPU_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__
#define __align__(n) __attribute__((aligned(n)))
#define __restrict__

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3() {} dim3(unsigned a, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(uint32_t a, uint32_t b) { return uint2{a, b}; }
struct __attribute__((aligned(16))) ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }

using std::min;
using std::max;
extern uint3 threadIdx, blockIdx;
extern dim3 blockDim, gridDim;

// ---- host API ------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct EmuStream *cudaStream_t;
typedef struct EmuEvent *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaEventDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes { cudaMemoryType type; int device; void *devicePointer; void *hostPointer; };
struct cudaDeviceProp { char name[256]; int major, minor, multiProcessorCount; size_t totalGlobalMem; };
typedef struct EmuPool *cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };

cudaError_t cudaMalloc(void **p, size_t n);
template <class T> static inline cudaError_t cudaMalloc(T **p, size_t n) { return cudaMalloc((void **)p, n); }
cudaError_t cudaFree(void *p);
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t);
template <class T> static inline cudaError_t cudaMallocAsync(T **p, size_t n, cudaStream_t s) { return cudaMallocAsync((void **)p, n, s); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t);
cudaError_t cudaMallocHost(void **p, size_t n);
template <class T> static inline cudaError_t cudaMallocHost(T **p, size_t n) { return cudaMallocHost((void **)p, n); }
cudaError_t cudaFreeHost(void *p);
cudaError_t cudaMemset(void *p, int v, size_t n);
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s = nullptr);
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind k);
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t st = nullptr);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags = 0);
cudaError_t cudaDeviceSynchronize();
cudaError_t cudaEventCreate(cudaEvent_t *e);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s = nullptr);
cudaError_t cudaEventSynchronize(cudaEvent_t e);
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b);
cudaError_t cudaGetDeviceCount(int *n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int dev);
cudaError_t cudaSetDevice(int dev);
cudaError_t cudaGetDevice(int *dev);
cudaError_t cudaGetLastError();
const char *cudaGetErrorString(cudaError_t e);
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p);
cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *pool, int dev);
cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t pool, cudaMemPoolAttr attr, void *value);
template <class F> static inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int *n, F, int, size_t) { *n = 2; return cudaSuccess; }
template <class F> static inline cudaError_t cudaFuncSetAttribute(F, int, int) { return cudaSuccess; }
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

// ---- launch --------------------------------------------------------------------------------
void emu_launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
uint8_t *emu_dyn_smem();
#define KJ_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch(dim3(grid), dim3(block), (smem), [=]() { kernel(__VA_ARGS__); })
#define KJ_DYN_SMEM(name) uint8_t *name = emu_dyn_smem()

// ---- device intrinsics -----------------------------------------------------------------------
void emu_yield();
void emu_syncthreads();
void emu_named_barrier(int id, int count);           // bar.sync id, count
void emu_named_arrive(int id, int count);            // bar.arrive id, count
uint64_t emu_warp_xchg(uint64_t v, int src_lane);   // every lane deposits v, returns lane src's value
uint32_t emu_warp_ballot(int pred);
uint64_t emu_warp_reduce_add(uint64_t v);
uint64_t emu_warp_reduce_op(uint64_t v, int op);     // 0 min, 1 max, 2 or, 3 and
unsigned emu_lane();

static inline void __syncthreads() { emu_syncthreads(); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { (void)emu_warp_ballot(0); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __nanosleep(unsigned) { emu_yield(); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __popcll(uint64_t x) { return __builtin_popcountll(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __ffsll(long long x) { return __builtin_ffsll(x); }
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (lo >> s) | (hi << (32 - s)) : lo; }
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) { s &= 31; return s ? (hi << s) | (lo >> (32 - s)) : hi; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t v = ((uint64_t)b << 32) | a; uint32_t r = 0;
    for (int i = 0; i < 4; ++i) { uint32_t sel = (s >> (4 * i)) & 7; r |= (uint32_t)((v >> (8 * sel)) & 0xFF) << (8 * i); }
    return r;
}
static inline uint32_t __vcmpeq4(uint32_t a, uint32_t b) {
    uint32_t r = 0; for (int i = 0; i < 4; ++i) if (((a >> (8 * i)) & 0xFF) == ((b >> (8 * i)) & 0xFF)) r |= 0xFFu << (8 * i); return r;
}
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; memcpy(&v, &d, 8); return v; }
static inline uint32_t __brev(uint32_t x) { uint32_t r = 0; for (int i = 0; i < 32; ++i) if (x >> i & 1) r |= 1u << (31 - i); return r; }

template <class T> static inline T emu_from_bits(uint64_t b) { T t; memcpy(&t, &b, sizeof(T)); return t; }
template <class T> static inline uint64_t emu_to_bits(T t) { static_assert(sizeof(T) <= 8, ""); uint64_t b = 0; memcpy(&b, &t, sizeof(T)); return b; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu_from_bits<T>(emu_warp_xchg(emu_to_bits(v), src & 31)); }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { int l = (int)emu_lane(); int s = l - (int)d; return emu_from_bits<T>(emu_warp_xchg(emu_to_bits(v), s < 0 ? l : s)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { int l = (int)emu_lane(); int s = l + (int)d; return emu_from_bits<T>(emu_warp_xchg(emu_to_bits(v), s > 31 ? l : s)); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu_from_bits<T>(emu_warp_xchg(emu_to_bits(v), (int)(emu_lane() ^ (unsigned)m) & 31)); }
static inline uint32_t __ballot_sync(unsigned, int pred) { return emu_warp_ballot(pred); }
static inline int __any_sync(unsigned, int pred) { return emu_warp_ballot(pred) != 0; }
static inline int __all_sync(unsigned, int pred) { return emu_warp_ballot(!pred) == 0; }
static inline uint32_t __activemask() { return 0xFFFFFFFFu; }
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v) { return (uint32_t)emu_warp_reduce_add(v); }
static inline uint32_t __reduce_min_sync(unsigned, uint32_t v) { return (uint32_t)emu_warp_reduce_op(v, 0); }
static inline uint32_t __reduce_max_sync(unsigned, uint32_t v) { return (uint32_t)emu_warp_reduce_op(v, 1); }
static inline uint32_t __reduce_or_sync(unsigned, uint32_t v) { return (uint32_t)emu_warp_reduce_op(v, 2); }
uint32_t emu_match_any(uint64_t v);
template <class T> static inline uint32_t __match_any_sync(unsigned, T v) { return emu_match_any(emu_to_bits(v)); }

// atomics (single OS thread: plain read-modify-write is atomic between yields)
template <class T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
template <class T> static inline T __ldg(const T *p) { return *p; }
