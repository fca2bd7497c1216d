// gpu_probe3.cu -- how fast can ONE warp issue shared-memory read-modify-write sequences?
// (round 2: the column-sweep kernel's ranking turn is one warp working alone on the counter table; its
// per-key cost decides the length of the warp chain.)  One CTA per SM, `nw` active warps, each warp
// issues R x 32 operations on lane-private words (bank == lane) with 32 distinct destination
// registers, so nothing but the hardware serialises them.  Prints cycles per operation per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/gpu_probe3 tools/gpu_probe3.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(512) probe(uint32_t *sink, long long *cycles, int rounds, int nw) {
    extern __shared__ uint32_t smem[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) smem[i] = 0;
    __syncthreads();
    if ((int)warp >= nw) return;
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(smem) + lane * 4u;
    uint32_t acc = 0, x = warp * 977u + 13u;
    uint32_t a[32];  // fixed random rows, own column: address generation stays out of the timed loop
#pragma unroll
    for (int k = 0; k < 32; ++k) {
        x = x * 1664525u + 1013904223u;
        a[k] = base + ((x >> 24) << 7);
    }
    long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
        uint32_t v[32];
        if (MODE == 0) {
#pragma unroll
            for (int k = 0; k < 32; ++k) asm volatile("atom.shared.add.u32 %0, [%1], 4;" : "=r"(v[k]) : "r"(a[k]) : "memory");
        } else if (MODE == 1) {
#pragma unroll
            for (int k = 0; k < 32; ++k) { asm volatile("red.shared.add.u32 [%0], 4;" ::"r"(a[k]) : "memory"); v[k] = 0; }
        } else if (MODE == 2) {
#pragma unroll
            for (int k = 0; k < 32; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(a[k]) : "memory");
        } else if (MODE == 3) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(a[k]) : "memory");
                asm volatile("red.shared.add.u32 [%0], 4;" ::"r"(a[k]) : "memory");
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("red.shared.add.u32 [%0], 4;" ::"r"(a[k]) : "memory");
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(a[k]) : "memory");
            }
        } else if (MODE == 5) {
#pragma unroll
            for (int k = 0; k < 32; ++k) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[k]), "r"(x + k) : "memory"); v[k] = 0; }
        } else if (MODE == 6) {
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(a[k]) : "memory");
                asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[k]), "r"(v[k] + 4u) : "memory");
            }
        } else if (MODE == 7) {  // all loads first, then all stores (no same-thread duplicates handled)
#pragma unroll
            for (int k = 0; k < 32; ++k) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[k]) : "r"(a[k]) : "memory");
#pragma unroll
            for (int k = 0; k < 32; ++k) asm volatile("st.shared.u32 [%0], %1;" ::"r"(a[k]), "r"(v[k] + 4u) : "memory");
        }
#pragma unroll
        for (int k = 0; k < 32; ++k) acc ^= v[k];
    }
    long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x * 32 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int MODE>
void run(const char *name, int sms, uint32_t *sink, long long *d_cyc) {
    const int rounds = 64;
    printf("\"%s\": {", name);
    bool first = true;
    for (int nw : {1, 2, 4, 8, 16}) {
        cudaMemset(d_cyc, 0, sizeof(long long) * sms * 32);
        probe<MODE><<<sms, 512, 32768>>>(sink, d_cyc, rounds, nw);
        cudaDeviceSynchronize();
        long long h[32];
        cudaMemcpy(h, d_cyc, sizeof(h), cudaMemcpyDeviceToHost);
        long long mx = 0;
        for (int w = 0; w < nw; ++w) mx = h[w] > mx ? h[w] : mx;
        printf("%s\"warps_%d_cyc_per_op\": %.2f", first ? "" : ", ", nw, (double)mx / (rounds * 32.0));
        first = false;
    }
    printf("}");
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *sink;
    long long *d_cyc;
    cudaMalloc(&sink, 4);
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 32);
    printf("{");
    run<0>("atom_ret", sms, sink, d_cyc); printf(", ");
    run<1>("red", sms, sink, d_cyc); printf(", ");
    run<2>("ld", sms, sink, d_cyc); printf(", ");
    run<3>("ld_then_red_same_word", sms, sink, d_cyc); printf(", ");
    run<4>("red_then_ld_same_word", sms, sink, d_cyc); printf(", ");
    run<5>("st", sms, sink, d_cyc); printf(", ");
    run<6>("ld_then_dependent_st", sms, sink, d_cyc); printf(", ");
    run<7>("32_ld_then_32_st", sms, sink, d_cyc);
    printf("}\n");
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "%s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
