// gpu_probe.cu -- bench-only probe (NOT part of the product library): the same-box bars the
// sort is compared with, and the instruction-throughput numbers that motivated the kernel design.
//   * device-resident cub::DeviceRadixSort::SortKeys / SortPairs (CCCL's Onesweep, the kernel the
//     reference's thrust::sort call resolves to -- SourceCode/Baseline1.cu:66-70)
//   * cudaMemcpy D2D and a plain uint4 copy kernel (the HBM roofline denominators)
//   * match.any / shared-atomic issue rates
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/gpu_probe.cu -o tools/_bin/gpu_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cub/device/device_radix_sort.cuh>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __host__ inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}
__global__ void fill(uint32_t *k, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        k[i] = (uint32_t)(sm64(0x5EED0001ULL + i) >> 32);
}
__global__ void copy4(const uint4 *__restrict__ a, uint4 *__restrict__ b, uint64_t n4) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
        b[i] = a[i];
}
__global__ void read4(const uint4 *__restrict__ a, uint64_t n4, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v = a[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void match_rate(uint32_t *sink, int iters, uint32_t seed) {
    uint32_t x = (uint32_t)sm64(seed + threadIdx.x + blockIdx.x * 1024u), acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            acc += __match_any_sync(0xffffffffu, (x >> (u * 3)) & 255u);
        }
        x = x * 1664525u + 1013904223u;
    }
    if (acc == 0x12345678u) *sink = acc;
}
__global__ void atoms_rate(uint32_t *sink, int iters, uint32_t seed, uint32_t mask) {
    __shared__ uint32_t h[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) h[i] = 0;
    __syncthreads();
    uint32_t x = (uint32_t)sm64(seed + threadIdx.x + blockIdx.x * 1024u);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) atomicAdd(&h[((x >> (u * 8)) & mask) + u * 256], 1u);
        x = x * 1664525u + 1013904223u;
    }
    __syncthreads();
    if (h[threadIdx.x] == 0x12345678u) *sink = 1;
}

// shared-memory op rates under controlled address patterns: mode 0 red.add, 1 atom.add (return used),
// 2 st.shared, 3 ld.shared.  `mask` limits the number of distinct addresses per warp instruction.
__global__ void smem_rate(uint32_t *sink, int iters, uint32_t seed, uint32_t mask, int mode) {
    __shared__ uint32_t h[32][256];
    const uint32_t warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
    __syncthreads();
    uint32_t x = (uint32_t)sm64(seed + threadIdx.x + blockIdx.x * 1024u), acc = 0;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(&h[warp][0]);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t a = base + (((x >> (u * 8)) & mask) << 2);
            if (mode == 0) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(a) : "memory");
            else if (mode == 1) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], 4;" : "=r"(r) : "r"(a) : "memory"); acc += r; }
            else if (mode == 2) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
            else { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory"); acc += r; }
        }
        x = x * 1664525u + 1013904223u;
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char **argv) {
    int log2n = argc > 1 ? atoi(argv[1]) : 28;
    uint64_t n = 1ull << log2n;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("{\"device\": \"%s\", \"sms\": %d, \"log2n\": %d", p.name, p.multiProcessorCount, log2n);
    uint32_t *k, *o, *v, *vo, *sink; void *tmp = nullptr; size_t tb = 0;
    CK(cudaMalloc(&k, n * 4)); CK(cudaMalloc(&o, n * 4)); CK(cudaMalloc(&v, n * 4)); CK(cudaMalloc(&vo, n * 4));
    CK(cudaMalloc(&sink, 4));
    fill<<<p.multiProcessorCount * 8, 256>>>(k, n); fill<<<p.multiProcessorCount * 8, 256>>>(v, n);
    CK(cudaDeviceSynchronize());

    float ms = time_ms([&] { CK(cudaMemcpyAsync(o, k, n * 4, cudaMemcpyDeviceToDevice)); }, 10);
    printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * n * 4 / ms / 1e6);
    ms = time_ms([&] { copy4<<<p.multiProcessorCount * 8, 512>>>((uint4 *)k, (uint4 *)o, n / 4); }, 10);
    printf(", \"copy_kernel_gbs\": %.1f", 2.0 * n * 4 / ms / 1e6);
    ms = time_ms([&] { read4<<<p.multiProcessorCount * 8, 512>>>((uint4 *)k, n / 4, sink); }, 10);
    printf(", \"read_kernel_gbs\": %.1f", 1.0 * n * 4 / ms / 1e6);

    CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, k, o, (int64_t)n));
    CK(cudaMalloc(&tmp, tb));
    ms = time_ms([&] { CK(cub::DeviceRadixSort::SortKeys(tmp, tb, k, o, (int64_t)n)); }, 10);
    printf(", \"cub_sortkeys_ms\": %.4f, \"cub_sortkeys_gkeys\": %.2f, \"cub_keys_temp_mb\": %.1f", ms, n / ms / 1e6, tb / 1e6);
    CK(cudaFree(tmp)); tb = 0;
    CK(cub::DeviceRadixSort::SortPairs(nullptr, tb, k, o, v, vo, (int64_t)n));
    CK(cudaMalloc(&tmp, tb));
    ms = time_ms([&] { CK(cub::DeviceRadixSort::SortPairs(tmp, tb, k, o, v, vo, (int64_t)n)); }, 10);
    printf(", \"cub_sortpairs_ms\": %.4f, \"cub_sortpairs_gpairs\": %.2f", ms, n / ms / 1e6);

    // instruction issue rates: warp-instructions per clock per SM
    const int iters = 2000, blocks = p.multiProcessorCount * 2, threads = 1024;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    ms = time_ms([&] { match_rate<<<blocks, threads>>>(sink, iters, 1); }, 3);
    double warp_instr = (double)blocks * (threads / 32) * iters * 8;
    printf(", \"match_any_warp_instr_per_us_per_sm\": %.1f", warp_instr / (ms * 1e3) / p.multiProcessorCount);
    for (uint32_t mask : {255u, 15u, 0u}) {
        ms = time_ms([&] { atoms_rate<<<blocks, threads>>>(sink, iters, 1, mask); }, 3);
        warp_instr = (double)blocks * (threads / 32) * iters * 4;
        printf(", \"atoms_mask%u_warp_instr_per_us_per_sm\": %.1f", mask, warp_instr / (ms * 1e3) / p.multiProcessorCount);
    }
    const char *names[4] = {"red_add", "atom_add_ret", "st", "ld"};
    for (int mode = 0; mode < 4; ++mode)
        for (uint32_t mask : {255u, 15u, 1u, 0u}) {
            ms = time_ms([&] { smem_rate<<<blocks, threads>>>(sink, iters, 1, mask, mode); }, 3);
            warp_instr = (double)blocks * (threads / 32) * iters * 4;
            printf(", \"smem_%s_mask%u_cycles_per_warp_instr\": %.2f", names[mode], mask,
                   (ms * 1e-3) * (clk_khz * 1e3) * p.multiProcessorCount / warp_instr);
        }
    printf(", \"clock_khz\": %d}\n", clk_khz);
    return 0;
}
