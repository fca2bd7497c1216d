// common.cuh -- shared types of the host driver and the kernels.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "ptx.cuh"

namespace b200sort {

constexpr int kMaxPasses = 32;       // nBits = 1 -> 32 digit passes
constexpr int kMaxRadixBits = 8;     // widest digit one kernel pass handles
constexpr int kHistThreads = 1024;
constexpr int kHistUnroll = 4;       // uint4 loads in flight per thread
constexpr uint32_t kDescValueMask = 0x3FFFFFFFu;
constexpr uint32_t kDescFlagMask = 0xC0000000u;

// The digit passes of one sort: pass p ranks keys by (key >> shift[p]) & ((1 << bits[p]) - 1).
// All passes of a sort run the kernel instantiated for `width` bits (2^width bins); a pass
// whose digit is narrower simply leaves the upper bins empty.
struct PassList {
    int count;
    int width;
    uint8_t shift[kMaxPasses];
    uint8_t bits[kMaxPasses];
};

struct HistArgs {
    const uint32_t *keys;
    uint64_t n;
    uint32_t *ghist;     // [count][2^width], zeroed before launch
    uint32_t *bin_base;  // [count][2][2^width]; slot 0 of each pass receives the exclusive scan
    uint32_t *done;      // zeroed CTA counter
    uint4 *zero_ptr;     // side job: region to clear (look-back descriptors)
    uint64_t zero_vecs;
    PassList passes;
};

struct PassArgs {
    const uint32_t *keys_in;
    const uint32_t *vals_in;
    uint32_t *keys_out;
    uint32_t *vals_out;
    const uint32_t *bin_base;  // [2^width] destination index of the first key of each bin
    uint32_t *carry_out;       // [2^width] bin_base of the next launch of this pass, or null
    uint32_t *desc;            // [num_tiles][2^width] look-back descriptors
    uint32_t *ticket;          // zeroed tile counter of this launch (persistent variants only)
    const uint64_t *bin_dst;   // optional [2 * 2^width] per-bin destination addresses
    uint32_t n;                // keys in this launch (< 2^30)
    uint32_t num_tiles;
    uint32_t shift;
    uint32_t mask;
    uint32_t parity;           // launch parity: rotates the descriptor status codes
    uint32_t prefetch;         // column sweep: the tile this many tiles ahead is prefetched into L2 (0 = off)
};

// Block-wide exclusive scan of one value per thread for the first `active_warps` warps
// (values of the other threads must be 0).  Every thread of the block must call it.
// s_warp_tot: >= 32 words of shared memory.  Contains two __syncthreads().
template <int THREADS>
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp_tot) {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // protect s_warp_tot against a previous use
    if (lane == 31) s_warp_tot[warp] = incl;
    __syncthreads();
    uint32_t before = 0;
    for (uint32_t w = 0; w < warp; ++w) before += s_warp_tot[w];
    return before + incl - v;
}

}  // namespace b200sort
