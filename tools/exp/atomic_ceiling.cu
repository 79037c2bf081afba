// tools/exp/atomic_ceiling.cu -- DEVELOPER EXPERIMENT for BASELINE config 5 (empty prefix, k = 31: every window is an
// emission, the hash table is the bound).  On one batch of emissions (default 240 M keys drawn from 18 M distinct 62-bit
// values: what 1 M reads of 150 bp give) it measures
//   (a) the ceiling: fire-and-forget random 64-bit atomic adds into a table far larger than L2,
//   (b) the count path's insert (kj_insert: key read, CAS for a new key, RED.ADD count, RED.MIN first-seen ordinal),
//   (c) the alternative of the north star: cub::DeviceRadixSort + cub::DeviceRunLengthEncode of the same keys.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../kmerjs_b200/csrc/kj_device.cuh"

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void gen_kernel(uint64_t *keys, uint64_t n, uint64_t distinct) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        keys[i] = kj_mix64(kj_mix64(i * 0x9E3779B97F4A7C15ull) % distinct + 1) >> 2;        // 62-bit keys (k = 31)
}
__global__ void red_kernel(const uint64_t *keys, uint64_t n, unsigned long long *table, uint64_t mask) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        atomicAdd(&table[kj_mix64(keys[i]) & mask], 1ull);
}
__global__ void insert_kernel(const uint64_t *keys, uint64_t n, KjTable t, KjCounters *ctr, int with_ord) {
    KjTable tt = t;
    if (!with_ord) tt.ords = nullptr;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        if (!kj_insert(tt, ctr, keys[i], i, 1)) atomicAdd(&ctr->n_overflow, 1ull);
}

template <class F>
static float time_it(F f, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
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

int main(int argc, char **argv) {
    const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 240000000ull;
    const uint64_t distinct = argc > 2 ? strtoull(argv[2], 0, 10) : 18000000ull;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int grid = prop.multiProcessorCount * 16;
    uint64_t *keys;
    CK(cudaMalloc(&keys, n * 8));
    gen_kernel<<<grid, 256>>>(keys, n, distinct);
    CK(cudaDeviceSynchronize());
    printf("%llu emissions, %llu distinct keys, %s\n", (unsigned long long)n, (unsigned long long)distinct, prop.name);

    uint64_t cap = 1;
    while (cap < 2 * distinct) cap <<= 1;
    {   // (a) random 64-bit RED into cap slots
        unsigned long long *table;
        CK(cudaMalloc(&table, cap * 8));
        CK(cudaMemset(table, 0, cap * 8));
        float ms = time_it([&]() { red_kernel<<<grid, 256>>>(keys, n, table, cap - 1); }, 3);
        printf("(a) random RED.ADD.64 into %llu slots (%.0f MB): %.2f ms  %.2f G atomics/s\n", (unsigned long long)cap, cap * 8 / 1e6, ms, n / ms / 1e6);
        CK(cudaFree(table));
    }
    for (int with_ord = 1; with_ord >= 0; --with_ord) {   // (b) the count path's insert
        KjTable t{};
        KjCounters *ctr;
        CK(cudaMalloc(&t.keys, cap * 8)); CK(cudaMalloc(&t.counts, cap * 8)); CK(cudaMalloc(&t.ords, cap * 8));
        CK(cudaMalloc(&ctr, sizeof(KjCounters)));
        t.mask = cap - 1;
        auto reset = [&]() {
            CK(cudaMemset(t.keys, 0xFF, cap * 8)); CK(cudaMemset(t.counts, 0, cap * 8)); CK(cudaMemset(t.ords, 0xFF, cap * 8));
            CK(cudaMemset(ctr, 0, sizeof(KjCounters)));
        };
        reset();
        cudaEvent_t a, b;
        CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a));
        insert_kernel<<<grid, 256>>>(keys, n, t, ctr, with_ord);            // cold table: every distinct key is inserted once
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float cold = 0; CK(cudaEventElapsedTime(&cold, a, b));
        float warm = time_it([&]() { insert_kernel<<<grid, 256>>>(keys, n, t, ctr, with_ord); }, 2);   // keys present: read + 2 RED
        printf("(b) kj_insert %s first-seen ordinal, table %llu slots: first pass %.2f ms (%.2f G/s), keys present %.2f ms (%.2f G/s)\n",
               with_ord ? "with" : "without", (unsigned long long)cap, cold, n / cold / 1e6, warm, n / warm / 1e6);
        CK(cudaFree(t.keys)); CK(cudaFree(t.counts)); CK(cudaFree(t.ords)); CK(cudaFree(ctr));
    }
    {   // (c) sort + run-length encode
        uint64_t *sorted, *uniq;
        unsigned long long *counts, *n_runs;
        CK(cudaMalloc(&sorted, n * 8)); CK(cudaMalloc(&uniq, n * 8)); CK(cudaMalloc(&counts, n * 8)); CK(cudaMalloc(&n_runs, 8));
        size_t t1 = 0, t2 = 0;
        CK(cub::DeviceRadixSort::SortKeys(nullptr, t1, keys, sorted, (int)n, 0, 62));
        CK(cub::DeviceRunLengthEncode::Encode(nullptr, t2, sorted, uniq, counts, n_runs, (int)n));
        void *tmp;
        CK(cudaMalloc(&tmp, t1 > t2 ? t1 : t2));
        float ms_sort = time_it([&]() { cub::DeviceRadixSort::SortKeys(tmp, t1, keys, sorted, (int)n, 0, 62); }, 2);
        float ms_rle = time_it([&]() { cub::DeviceRunLengthEncode::Encode(tmp, t2, sorted, uniq, counts, n_runs, (int)n); }, 2);
        unsigned long long runs = 0;
        CK(cudaMemcpy(&runs, n_runs, 8, cudaMemcpyDeviceToHost));
        printf("(c) cub radix sort (62 bits) %.2f ms + run-length encode %.2f ms = %.2f ms (%.2f G keys/s), %llu runs; first-seen ordinals would need a pair sort (more)\n",
               ms_sort, ms_rle, ms_sort + ms_rle, n / (ms_sort + ms_rle) / 1e6, runs);
    }
    return 0;
}
