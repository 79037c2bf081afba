// tools/exp/scan_exp.cu -- DEVELOPER EXPERIMENT (not product code): how fast can one warp-autonomous
// pipeline stream FASTQ bytes on a B200?  Each warp owns a ring of 4 KiB tiles in shared memory, filled by
// the TMA engine through a 2-D tensor map [rows of 128 bytes] with the 128-byte swizzle, so that a lane can
// read its own contiguous 128 bytes (8 chunks of 16) without bank conflicts.  Modes: 0 sum, 1 convert
// (2-bit pack + newline mask), 2 convert + prefix search + queue push, 3 plain LDG.128 grid-stride sum.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../kmerjs_b200/csrc/kj_bits.cuh"

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)
#define TILE_BYTES 4096u

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void tma_tile(void *dst, const CUtensorMap *tm, int32_t row, uint64_t *bar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(TILE_BYTES) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(tm), "r"(0), "r"(row), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t a = smem_u32(bar);
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

struct Pat { uint32_t f[5], r[5]; uint32_t c0a, c7f; };

// newline mask with the constants held in registers (one LOP3 for (w ^ a) & b)
__device__ __forceinline__ uint32_t nl_flags_r(uint32_t w, uint32_t c0a, uint32_t c7f) {
    const uint32_t t = ((w ^ c0a) & c7f) + c7f;
    return ~(t | w) & 0x80808080u;
}
__device__ __forceinline__ uint32_t nl16_r(uint4 v, uint32_t c0a, uint32_t c7f) {
    uint32_t lo = __dp4a(nl_flags_r(v.x, c0a, c7f), 0x08040201u, 0u);
    lo = __dp4a(nl_flags_r(v.y, c0a, c7f), 0x80402010u, lo);
    uint32_t hi = __dp4a(nl_flags_r(v.z, c0a, c7f), 0x08040201u, 0u);
    hi = __dp4a(nl_flags_r(v.w, c0a, c7f), 0x80402010u, hi);
    return (lo >> 7) | ((hi << 1) & 0xFF00u);
}

template <int MODE, int NSTAGE, int REGC>
__global__ void __launch_bounds__(256) exp_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n_tiles,
                                                  const Pat pat, unsigned long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    uint8_t *my = smem + (size_t)warp * NSTAGE * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nwarps * NSTAGE * TILE_BYTES) + warp * NSTAGE;
    uint32_t *queue = reinterpret_cast<uint32_t *>(smem + (size_t)nwarps * NSTAGE * TILE_BYTES + nwarps * NSTAGE * 8) + warp * 65;
    if (lane == 0) {
        for (int s = 0; s < NSTAGE; ++s) bar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        queue[64] = 0;
    }
    __syncwarp();
    const uint64_t G = (uint64_t)gridDim.x * nwarps, g = (uint64_t)blockIdx.x * nwarps + warp;
    if (lane == 0)
        for (int s = 0; s < NSTAGE; ++s) {
            const uint64_t t = g + (uint64_t)s * G;
            if (t < n_tiles) tma_tile(my + s * TILE_BYTES, &tmap, (int32_t)(t * 32), &bars[s]);
        }
    uint32_t off[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) off[i] = smem_u32(my) + lane * 128u + (((uint32_t)i << 4) ^ ((lane & 7u) << 4));
    uint32_t c0a = pat.c0a, c7f = pat.c7f;
    unsigned long long acc = 0;
    uint32_t cnt = 0;
    uint32_t phase = 0;
    uint64_t t = g;
    while (t < n_tiles) {
#pragma unroll
        for (int s = 0; s < NSTAGE; ++s) {
            if (t < n_tiles) {
                bar_wait(&bars[s], (phase >> s) & 1u);
                phase ^= 1u << s;
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = lds128(off[i] + s * TILE_BYTES);
                if (MODE == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc += (unsigned long long)(v[i].x ^ v[i].y) + (v[i].z ^ v[i].w);
                } else {
                    uint32_t cw[10];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        cw[i] = kj_pack16(v[i].x, v[i].y, v[i].z, v[i].w);
                        const uint32_t nl = REGC ? nl16_r(v[i], c0a, c7f) : kj_nl16(v[i].x, v[i].y, v[i].z, v[i].w);
                        cnt += __popc(nl);
                    }
                    if (MODE == 1) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc ^= cw[i];
                    } else {
                        cw[8] = __shfl_down_sync(0xFFFFFFFFu, cw[0], 1);
                        cw[9] = __shfl_down_sync(0xFFFFFFFFu, cw[1], 1);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const uint32_t a0 = cw[i], a1 = cw[i + 1];
                            uint32_t accf = 0, accr = 0;
#pragma unroll
                            for (int j = 0; j < 5; ++j) {
                                accf |= __funnelshift_r(a0, a1, 2 * j) ^ pat.f[j];
                                accr |= __funnelshift_r(a0, a1, 2 * (11 + j)) ^ pat.r[j];
                            }
                            const uint32_t z = kj_zero_lanes(accf) | (kj_zero_lanes(accr) << 1);
                            if (z) {
                                const uint32_t q = atomicAdd(&queue[64], 1u);
                                queue[q & 63] = z;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) {
                    const uint64_t tn = t + (uint64_t)NSTAGE * G;
                    if (tn < n_tiles) tma_tile(my + s * TILE_BYTES, &tmap, (int32_t)(tn * 32), &bars[s]);
                }
                t += G;
            }
        }
    }
    acc += cnt;
    if (MODE == 2) acc += queue[64] + queue[lane];
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if (lane == 0) atomicAdd(out, acc);
}

// verification: copy the tile back out through the same swizzled reads
template <int NSTAGE>
__global__ void __launch_bounds__(256) copy_kernel(const __grid_constant__ CUtensorMap tmap, uint64_t n_tiles, uint4 *dst) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    uint8_t *my = smem + (size_t)warp * NSTAGE * TILE_BYTES;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)nwarps * NSTAGE * TILE_BYTES) + warp * NSTAGE;
    if (lane == 0) {
        bar_init(&bars[0], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    const uint64_t G = (uint64_t)gridDim.x * nwarps, g = (uint64_t)blockIdx.x * nwarps + warp;
    uint32_t ph = 0;
    for (uint64_t t = g; t < n_tiles; t += G) {
        if (lane == 0) tma_tile(my, &tmap, (int32_t)(t * 32), &bars[0]);
        __syncwarp();
        bar_wait(&bars[0], ph);
        ph ^= 1;
        for (int i = 0; i < 8; ++i) {
            const uint4 v = lds128(smem_u32(my) + lane * 128u + (((uint32_t)i << 4) ^ ((lane & 7u) << 4)));
            dst[t * 256 + lane * 8 + i] = v;
        }
        __syncwarp();
    }
}

__global__ void ldg_kernel(const uint4 *p, uint64_t n16, unsigned long long *out) {
    unsigned long long acc = 0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 a = __ldg(p + i), b = __ldg(p + i + stride), c = __ldg(p + i + 2 * stride), d = __ldg(p + i + 3 * stride);
        acc += (unsigned long long)(a.x ^ a.y) + (a.z ^ a.w) + (b.x ^ b.y) + (b.z ^ b.w) + (c.x ^ c.y) + (c.z ^ c.w) + (d.x ^ d.y) + (d.z ^ d.w);
    }
    for (; i < n16; i += stride) { uint4 a = __ldg(p + i); acc += (unsigned long long)(a.x ^ a.y) + (a.z ^ a.w); }
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

__global__ void gen_kernel(uint8_t *buf, uint64_t n_rec) {
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_rec; r += (uint64_t)gridDim.x * blockDim.x) {
        uint8_t *p = buf + r * 346;
        uint64_t h = kj_mix64(r + 12345);
        for (int i = 0; i < 41; ++i) p[i] = "@SIM:1:FC:1:1234:56789:12345 1:N:0:CGATGT"[i];
        p[41] = '\n';
        for (int i = 0; i < 150; ++i) { if ((i & 31) == 0) h = kj_mix64(h + i); p[42 + i] = "ACGT"[(h >> (2 * (i & 31))) & 3]; }
        p[192] = '\n'; p[193] = '+'; p[194] = '\n';
        for (int i = 0; i < 150; ++i) { if ((i & 7) == 0) h = kj_mix64(h + i + 7); p[195 + i] = (uint8_t)('#' + ((h >> (8 * (i & 7))) & 0xFF) % 39); }
        p[345] = '\n';
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <class F>
static float time_it(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

template <int MODE, int NSTAGE, int REGC>
static void run(const CUtensorMap &tm, uint64_t n_tiles, const Pat &pat, unsigned long long *d_out, int warps, int sms, double gb) {
    const size_t smem = (size_t)warps * NSTAGE * TILE_BYTES + warps * NSTAGE * 8 + warps * 65 * 4 + 64;
    CK(cudaFuncSetAttribute(exp_kernel<MODE, NSTAGE, REGC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, exp_kernel<MODE, NSTAGE, REGC>, warps * 32, smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, exp_kernel<MODE, NSTAGE, REGC>));
    const int grid = sms * occ;
    float ms = time_it([&]() { exp_kernel<MODE, NSTAGE, REGC><<<grid, warps * 32, smem>>>(tm, n_tiles, pat, d_out); }, 5);
    printf("mode %d stages %d regconst %d warps/cta %d occ %d (warps/SM %d) regs %d: %.3f ms  %.0f GB/s\n", MODE, NSTAGE, REGC, warps, occ,
           occ * warps, fa.numRegs, ms, gb / (ms * 1e-3));
}

int main(int argc, char **argv) {
    const uint64_t n_rec = argc > 1 ? strtoull(argv[1], 0, 10) : 10000000ull;
    const uint64_t n = n_rec * 346;
    const uint64_t n_tiles = n / TILE_BYTES;
    const double gb = (double)n_tiles * TILE_BYTES / 1e9;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    uint8_t *buf;
    CK(cudaMalloc(&buf, n + 4096));
    gen_kernel<<<sms * 8, 256>>>(buf, n_rec);
    CK(cudaDeviceSynchronize());
    unsigned long long *d_out;
    CK(cudaMalloc(&d_out, 8));
    CK(cudaMemset(d_out, 0, 8));

    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap tm;
    cuuint64_t gdim[2] = {128, n / 128};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {128, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, buf, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d, %llu tiles, %.3f GB, %d SMs\n", (int)r, (unsigned long long)n_tiles, gb, sms);
    if (r != CUDA_SUCCESS) return 1;

    {   // correctness of the swizzled read-back on the first 64 MiB
        const uint64_t vt = n_tiles < 16384 ? n_tiles : 16384;
        uint4 *dst;
        CK(cudaMalloc(&dst, vt * TILE_BYTES));
        const size_t smem = 8 * 1 * TILE_BYTES + 8 * 8 + 64;
        CK(cudaFuncSetAttribute(copy_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        copy_kernel<1><<<sms, 256, smem>>>(tm, vt, dst);
        CK(cudaDeviceSynchronize());
        std::vector<uint8_t> a(vt * TILE_BYTES), b(vt * TILE_BYTES);
        CK(cudaMemcpy(a.data(), buf, a.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), dst, b.size(), cudaMemcpyDeviceToHost));
        printf("swizzled read-back %s\n", memcmp(a.data(), b.data(), a.size()) == 0 ? "OK" : "MISMATCH");
        CK(cudaFree(dst));
    }
    Pat pat;
    const uint8_t pf[5] = {'A', 'T', 'G', 'A', 'C'}, pr[5] = {'G', 'T', 'C', 'A', 'T'};
    for (int i = 0; i < 5; ++i) { pat.f[i] = kj_code(pf[i]) * 0x55555555u; pat.r[i] = kj_code(pr[i]) * 0x55555555u; }
    pat.c0a = 0x0A0A0A0Au; pat.c7f = 0x7F7F7F7Fu;

    float ms = time_it([&]() { ldg_kernel<<<sms * 8, 256>>>((const uint4 *)buf, n / 16, d_out); }, 5);
    printf("ldg grid-stride sum: %.3f ms  %.0f GB/s\n", ms, (double)n / 1e9 / (ms * 1e-3));
    ms = time_it([&]() { ldg_kernel<<<sms * 16, 512>>>((const uint4 *)buf, n / 16, d_out); }, 5);
    printf("ldg grid-stride sum (16x512): %.3f ms  %.0f GB/s\n", ms, (double)n / 1e9 / (ms * 1e-3));
    for (int warps : {4, 8}) {
        run<0, 2, 0>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<0, 3, 0>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<0, 4, 0>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<1, 2, 0>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<1, 2, 1>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<1, 3, 1>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<2, 2, 0>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<2, 2, 1>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<2, 3, 1>(tm, n_tiles, pat, d_out, warps, sms, gb);
        run<2, 4, 1>(tm, n_tiles, pat, d_out, warps, sms, gb);
    }
    return 0;
}
