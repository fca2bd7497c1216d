// gpu_probe2.cu -- round-2 micro-benchmarks behind the redesign of the digit-pass kernel
// (bench-only, NOT part of the product library).  Everything is reported in SM cycles per warp
// instruction (or per bulk copy) per SM, with the loop bodies kept free of address arithmetic so
// that the shared-memory pipe, not the integer pipe, is what is measured.
//   * shared-memory atomics / stores on lane-private words (bank == lane) versus random banks
//   * cp.async.bulk shared -> global (UBLKCP) for the short runs a digit bin produces (48..2048 B)
//   * cp.async.bulk global -> shared of 1408-byte column chunks (the tile load of the pass kernel)
//   * named-barrier hand-off latency between the warps of a CTA, SHFL rate
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/gpu_probe2.cu -o tools/_bin/gpu_probe2
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __host__ inline uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31);
}

// mode: 0 red lane-private, 1 atom(ret) lane-private, 2 red random bank, 3 atom(ret) random bank,
//       4 st random bank, 5 st lane-linear, 6 ld random bank, 7 ld lane-linear (conflict-free),
//       8 shfl.up, 9 ld.v4 conflict-free (quad stride odd)
template <int MODE>
__global__ void __launch_bounds__(256) smem_rate(uint32_t *sink, int iters, uint32_t seed) {
    extern __shared__ __align__(1024) uint32_t sm[];  // 16 KB table (128 rows x 32 lanes) + 44 KB buffer
    const uint32_t lane = threadIdx.x & 31u;
    for (int i = threadIdx.x; i < 15 * 1024; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint32_t base = smem_u32(sm);
    uint32_t a[8];
    uint64_t x = sm64(seed + threadIdx.x + blockIdx.x * 1024ull);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const uint32_t r = (uint32_t)(x >> (u * 8)) & 0xFFu;
        if (MODE == 0 || MODE == 1) a[u] = base + ((r & 127u) << 7) + lane * 4u;           // row random, bank = lane
        else if (MODE == 2 || MODE == 3) a[u] = base + (threadIdx.x >> 5) * 1024u + r * 4u;  // per-warp 256-entry table
        else if (MODE == 4 || MODE == 6) a[u] = base + 16384u + (((uint32_t)(x >> (u * 7)) % 11264u) << 2);
        else if (MODE == 9) a[u] = base + 16384u + (lane * 89u + (threadIdx.x >> 5) * 11u + u) * 16u;
        else a[u] = base + 16384u + (threadIdx.x + u * 256u) * 4u;
    }
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (MODE == 0 || MODE == 2) asm volatile("red.shared.add.u32 [%0], 4;" ::"r"(a[u]) : "memory");
            else if (MODE == 1 || MODE == 3) { uint32_t r; asm volatile("atom.shared.add.u32 %0, [%1], 4;" : "=r"(r) : "r"(a[u]) : "memory"); acc ^= r; }
            else if (MODE == 4 || MODE == 5) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[u]), "r"(acc) : "memory");
            else if (MODE == 8) { acc += __shfl_up_sync(0xffffffffu, acc + u, 1); }
            else if (MODE == 9) { uint32_t r0, r1, r2, r3; asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a[u]) : "memory"); acc ^= r0 ^ r1 ^ r2 ^ r3; }
            else { uint32_t r; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a[u] ^ ((i & 1) << 7)) : "memory"); acc += r; }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

// cp.async.bulk shared::cta -> global in pieces of `bytes`; `issuers` threads of the CTA issue `per_thread`
// copies each per round; destination offsets are 16-byte aligned and scattered inside the CTA's slice.
__global__ void __launch_bounds__(256) bulk_s2g_rate(char *dst, size_t slice, int rounds, int issuers, int per_thread, uint32_t bytes) {
    extern __shared__ __align__(1024) uint32_t sm[];
    for (int i = threadIdx.x; i < 11 * 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    char *mine = dst + (size_t)blockIdx.x * slice;
    const uint32_t sbase = smem_u32(sm);
    const uint32_t span = 44u * 1024u - bytes;  // smem source offsets stay inside the buffer
    for (int r = 0; r < rounds; ++r) {
        if ((int)threadIdx.x < issuers) {
            for (int k = 0; k < per_thread; ++k) {
                const uint32_t idx = (uint32_t)(r * issuers * per_thread + k * issuers + threadIdx.x);
                const uint32_t so = ((idx * 2654435761u) % span) & ~15u;
                const size_t go = (((size_t)idx * (bytes + 16u)) % (slice - bytes)) & ~(size_t)15;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(mine + go), "r"(sbase + so), "r"(bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if ((int)threadIdx.x < issuers) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the same bytes written with st.global.u32 from shared memory (coalesced runs of `bytes`), for comparison
__global__ void __launch_bounds__(256) stg_rate(uint32_t *dst, size_t slice_words, int rounds) {
    extern __shared__ __align__(1024) uint32_t sm[];
    for (int i = threadIdx.x; i < 11 * 1024; i += blockDim.x) sm[i] = i;
    __syncthreads();
    uint32_t *mine = dst + (size_t)blockIdx.x * slice_words;
    for (int r = 0; r < rounds; ++r) {
#pragma unroll 4
        for (int k = 0; k < 44; ++k) {
            const uint32_t j = threadIdx.x + k * 256u;
            mine[((size_t)r * 11264u + j) % slice_words] = sm[j];
        }
    }
}

// global -> shared: 32 column chunks of 1408 B per tile (what the pass kernel loads), one lane each
__global__ void __launch_bounds__(256) bulk_g2s_rate(const char *src, size_t n_tiles, uint32_t *sink) {
    extern __shared__ __align__(1024) uint32_t sm[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t phase = 0, acc = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (threadIdx.x == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(32u * 1408u) : "memory");
        __syncthreads();
        if (threadIdx.x < 32) {
            const char *g = src + t * 45056ull + threadIdx.x * 1408u;
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm) + threadIdx.x * 1424u), "l"(g), "r"(1408u), "r"(b) : "memory");
        }
        asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(b), "r"(phase) : "memory");
        phase ^= 1u;
        acc ^= sm[threadIdx.x];
        __syncthreads();
    }
    if (acc == 0x12345678u) *sink = acc;
}

// 8 warps pass a token: warp w waits on named barrier w, works (n_atoms lane-private atomics), arrives on w+1
__global__ void __launch_bounds__(256) chain_latency(uint32_t *sink, long long *cycles, int rounds, int n_atoms) {
    extern __shared__ __align__(1024) uint32_t sm[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    const uint32_t base = smem_u32(sm) + lane * 4u;
    uint32_t acc = 0;
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        if (warp > 0) asm volatile("bar.sync %0, 64;" ::"r"(warp) : "memory");
        for (int k = 0; k < n_atoms; ++k) {
            uint32_t v;
            asm volatile("atom.shared.add.u32 %0, [%1], 4;" : "=r"(v) : "r"(base + (((k * 37u + warp * 11u) & 127u) << 7)) : "memory");
            acc ^= v;
        }
        if (warp < 7) asm volatile("bar.arrive %0, 64;" ::"r"(warp + 1u) : "memory");
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F> float time_ms(F f, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / reps;
}

template <int MODE> void run_smem(const char *name, int sms, double clk_hz, uint32_t *sink) {
    const int iters = 4000, blocks = sms * 3, threads = 256;
    const size_t smem = 64 * 1024;
    CK(cudaFuncSetAttribute(smem_rate<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const float ms = time_ms([&] { smem_rate<MODE><<<blocks, threads, smem>>>(sink, iters, 7); }, 3);
    const double wi = (double)blocks * (threads / 32) * iters * 8;
    printf(", \"%s_cyc_per_warp_instr\": %.3f", name, ms * 1e-3 * clk_hz * sms / wi);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double clk_hz = clk_khz * 1e3;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d", p.name, sms, clk_khz);
    uint32_t *sink; CK(cudaMalloc(&sink, 4));

    run_smem<0>("red_lane_private", sms, clk_hz, sink);
    run_smem<1>("atom_ret_lane_private", sms, clk_hz, sink);
    run_smem<2>("red_random_bank", sms, clk_hz, sink);
    run_smem<3>("atom_ret_random_bank", sms, clk_hz, sink);
    run_smem<4>("st_random_bank", sms, clk_hz, sink);
    run_smem<5>("st_linear", sms, clk_hz, sink);
    run_smem<6>("ld_random_bank", sms, clk_hz, sink);
    run_smem<7>("ld_linear", sms, clk_hz, sink);
    run_smem<8>("shfl_up", sms, clk_hz, sink);
    run_smem<9>("ld_v4_blocked", sms, clk_hz, sink);

    // ---- bulk shared -> global ----
    const size_t smem = 44 * 1024 + 1024;
    CK(cudaFuncSetAttribute(bulk_s2g_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(stg_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(bulk_g2s_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ctas = sms * 3;
    const size_t slice = 4u << 20;  // 4 MiB per CTA -> 1.7 GiB, far larger than L2
    char *dst; CK(cudaMalloc(&dst, (size_t)ctas * slice));
    CK(cudaMemset(dst, 0, (size_t)ctas * slice));
    for (uint32_t bytes : {48u, 112u, 176u, 256u, 512u, 2048u, 16384u}) {
        for (int issuers : {256, 32}) {
            // ~44 KB per round per CTA, like one tile
            int per_thread = (int)((44u * 1024u / bytes + issuers - 1) / issuers);
            if (per_thread < 1) per_thread = 1;
            const int rounds = 40;
            const float ms = time_ms([&] { bulk_s2g_rate<<<ctas, 256, smem>>>(dst, slice, rounds, issuers, per_thread, bytes); }, 3);
            const double ops = (double)ctas * rounds * issuers * per_thread;
            printf(", \"s2g_%uB_%dissuers\": {\"cyc_per_op_per_sm\": %.2f, \"gbs\": %.1f}", bytes, issuers,
                   ms * 1e-3 * clk_hz * sms / ops, ops * bytes / ms / 1e6);
        }
    }
    {
        const int rounds = 40;
        const float ms = time_ms([&] { stg_rate<<<ctas, 256, smem>>>((uint32_t *)dst, slice / 4, rounds); }, 3);
        printf(", \"stg_tile_gbs\": %.1f", (double)ctas * rounds * 45056.0 / ms / 1e6);
    }
    {
        const size_t n_tiles = ((size_t)ctas * slice) / 45056ull;
        const float ms = time_ms([&] { bulk_g2s_rate<<<ctas, 256, smem>>>(dst, n_tiles, sink); }, 3);
        printf(", \"g2s_tile_gbs\": %.1f", (double)n_tiles * 45056.0 / ms / 1e6);
    }
    {
        long long *d_cyc; CK(cudaMalloc(&d_cyc, 8));
        CK(cudaFuncSetAttribute(chain_latency, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
        for (int n_atoms : {0, 44}) {
            long long h = 0;
            const int rounds = 200;
            chain_latency<<<1, 256, 16384>>>(sink, d_cyc, rounds, n_atoms);
            CK(cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost));
            printf(", \"chain8_alone_%datoms_cyc_per_round\": %.1f", n_atoms, (double)h / rounds);
            chain_latency<<<sms * 3, 256, 16384>>>(sink, d_cyc, rounds, n_atoms);
            CK(cudaMemcpy(&h, d_cyc, 8, cudaMemcpyDeviceToHost));
            printf(", \"chain8_3ctas_%datoms_cyc_per_round\": %.1f", n_atoms, (double)h / rounds);
        }
    }
    printf("}\n");
    return 0;
}
