// cub_bench.cu -- bench-only (NOT part of the product library, which links neither CUB nor Thrust):
// device-resident cub::DeviceRadixSort::SortKeys / SortPairs on 2^log2n uniform uint32 keys -- the kernel
// the reference's sortByThrust (SourceCode/Baseline1.cu:66-70) resolves to with this toolkit (CCCL Onesweep,
// Policy1000).  Same generator as the product's workload C2 (splitmix64 counter hash).  Prints ONE JSON line.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/cub_bench.cu -o tools/_bin/cub_bench
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("{\"error\": \"%s at line %d\"}\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
__global__ void fill(uint32_t *k, uint32_t *v, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        k[i] = (uint32_t)(sm64(0x5EED0001ULL + i) >> 32);
        v[i] = (uint32_t)i;
    }
}
template <typename F> float time_ms(F f, int warm, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}
int main(int argc, char **argv) {
    const int log2n = argc > 1 ? atoi(argv[1]) : 28;
    const int reps = argc > 2 ? atoi(argv[2]) : 10;
    const uint64_t n = 1ull << log2n;
    uint32_t *k, *o, *v, *vo; void *tmp = nullptr; size_t tb = 0;
    CK(cudaMalloc(&k, n * 4)); CK(cudaMalloc(&o, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&vo, n * 4));
    fill<<<148 * 8, 256>>>(k, v, n);
    CK(cudaDeviceSynchronize());
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, k, o, (int64_t)n));
    CK(cudaMalloc(&tmp, tb));
    const float keys_ms = time_ms([&] { CK(cub::DeviceRadixSort::SortKeys(tmp, tb, k, o, (int64_t)n)); }, 3, reps);
    CK(cudaFree(tmp)); tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k, o, v, vo, (int64_t)n));
    CK(cudaMalloc(&tmp, tb));
    const float pairs_ms = time_ms([&] { CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k, o, v, vo, (int64_t)n)); }, 3, reps);
    printf("{\"impl\": \"cub::DeviceRadixSort (CCCL %d.%d.%d), device-resident, CUDA events\", \"n\": %llu, "
           "\"sortkeys_ms\": %.4f, \"sortkeys_keys_per_s\": %.4e, \"sortpairs_ms\": %.4f, \"sortpairs_pairs_per_s\": %.4e}\n",
           CUB_MAJOR_VERSION, CUB_MINOR_VERSION, CUB_SUBMINOR_VERSION, (unsigned long long)n, keys_ms, n / (keys_ms * 1e-3),
           pairs_ms, n / (pairs_ms * 1e-3));
    return 0;
}
