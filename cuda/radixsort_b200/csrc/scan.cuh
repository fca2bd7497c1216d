// scan.cuh -- device-wide exclusive prefix sum of uint32 (mod 2^32) in ONE pass with a decoupled
// look-back: the public form of the reference's scan stage -- scan() + scanBlocks +
// addScannedBlockSumsToScannedBlocks (SourceCode/Parallel7.cu:408-528), which scans 2*blockSize
// elements per block with a Blelloch tree, copies the block totals to the HOST, scans them there
// and launches a second kernel to add them back -- and of the standalone study
// Docs/Snippets/PrefixSum-WorkEfficient.cu.  8 bytes of HBM traffic per element, no host round trip.
#pragma once
#include "common.cuh"

namespace b200sort {

// Tile geometries (threads x uint4 per thread); the host picks one with scan_tile(variant).  The smallest
// tile bounds the descriptor array.
constexpr int kScanNumVariants = 13;
constexpr int kScanGeom[kScanNumVariants][2] = {{256, 4}, {512, 4}, {512, 8}, {1024, 4}, {1024, 8}, {128, 8}, {256, 8}, {256, 16}, {128, 16},
                                                  {1024, 4}, {512, 8}, {512, 4}, {256, 8}};  // 9..12: persistent kernel
constexpr int kScanMinTile = 256 * 4 * 4;
inline int scan_tile(int variant) { return kScanGeom[variant][0] * kScanGeom[variant][1] * 4; }
constexpr int kScanFirstPersistent = 9;

// descriptor: {status in the high word | value in the low word}; 0 = not ready
constexpr uint64_t kScanAggregate = 1ull << 32;
constexpr uint64_t kScanInclusive = 2ull << 32;

__device__ __forceinline__ uint64_t ld_relaxed_gpu64(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu64(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// desc: one zeroed uint64 per tile.  Requires in/out 16-byte aligned (checked by the host); the
// last partial tile is handled with scalar accesses.
template <int kScanThreads, int kScanVecs>
__global__ void __launch_bounds__(kScanThreads) exclusive_scan_kernel(const uint32_t *in, uint32_t *out, uint64_t n,
                                                                       uint64_t *desc, uint32_t prefetch) {
    constexpr int kScanTile = kScanThreads * kScanVecs * 4;
    // a tile `prefetch` tiles further on is requested into L2 (the copy engine does it; 16-byte aligned input only)
    if (prefetch != 0u && threadIdx.x == 0) {
        const uint64_t pt = blockIdx.x + (uint64_t)prefetch;
        if ((pt + 1) * kScanTile <= n && (reinterpret_cast<uintptr_t>(in) & 15u) == 0)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(in + pt * kScanTile), "r"((uint32_t)kScanTile * 4u) : "memory");
    }
    __shared__ uint32_t s_warp_tot[32];
    __shared__ uint32_t s_prefix;
    __shared__ uint32_t s_total;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint64_t tile = blockIdx.x;
    const uint64_t base = tile * kScanTile;
    const bool full = base + kScanTile <= n;

    // Warp w owns kScanVecs*128 consecutive elements and reads them in kScanVecs rounds of one
    // fully coalesced 512-byte request: lane l of round q holds elements [(w*kScanVecs+q)*128 + 4l, +4).
    const uint32_t warp = tid >> 5;
    uint32_t v[kScanVecs * 4];
    const uint64_t wbase = base + (uint64_t)warp * (kScanVecs * 128);
    if (full) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + wbase) + lane;
#pragma unroll
        for (int q = 0; q < kScanVecs; ++q) {
            const uint4 x = ld_stream_v4(src + q * 32);
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
    } else {
#pragma unroll
        for (int q = 0; q < kScanVecs; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint64_t i = wbase + (uint64_t)q * 128 + lane * 4 + e;
                v[4 * q + e] = (i < n) ? in[i] : 0u;
            }
    }
    // exclusive scan inside the warp's slice: within the vector, across lanes, across rounds
    uint32_t sum = 0;  // becomes the warp total (uniform across the warp)
#pragma unroll
    for (int q = 0; q < kScanVecs; ++q) {
        uint32_t t = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const uint32_t x = v[4 * q + e];
            v[4 * q + e] = t;
            t += x;
        }
        uint32_t incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        const uint32_t lane_excl = sum + incl - t;
#pragma unroll
        for (int e = 0; e < 4; ++e) v[4 * q + e] += lane_excl;
        sum += __shfl_sync(0xffffffffu, incl, 31);
    }
    // block scan over the warp totals (one value per warp, carried by lane 0)
    const uint32_t warp_excl_all = block_exclusive_scan<kScanThreads>(lane == 0 ? sum : 0u, s_warp_tot);
    const uint32_t thread_excl = __shfl_sync(0xffffffffu, warp_excl_all, 0);
    // tile total = exclusive prefix of the last warp + its total
    if (tid == kScanThreads - 32) {
        const uint32_t total = thread_excl + sum;
        st_relaxed_gpu64(desc + tile, (tile == 0 ? kScanInclusive : kScanAggregate) | total);
        s_total = total;
    }
    __syncthreads();
    // look-back by warp 0: 32 predecessors per step
    if (tid < 32) {
        uint32_t excl = 0;
        if (tile != 0) {
            int64_t t = (int64_t)tile - 1;
            for (;;) {
                const int64_t mine = t - (int64_t)lane;
                uint64_t d = kScanInclusive;  // before tile 0: an inclusive prefix of zero
                // every lane reads its predecessor once per poll; only the descriptors NEARER than the nearest
                // inclusive one have to be ready, so the warp re-polls until that prefix is complete
                uint32_t incl_mask, ready_mask, upto;
                for (;;) {
                    if (mine >= 0) d = ld_relaxed_gpu64(desc + mine);
                    incl_mask = __ballot_sync(0xffffffffu, (d >> 32) == 2u);
                    ready_mask = __ballot_sync(0xffffffffu, (d >> 32) != 0u);
                    upto = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 31u;
                    const uint32_t need = (upto == 31u) ? 0xffffffffu : ((2u << upto) - 1u);
                    if ((ready_mask & need) == need) break;
                }
                const uint32_t part = (lane <= upto) ? (uint32_t)d : 0u;
                excl += __reduce_add_sync(0xffffffffu, part);
                if (incl_mask) break;
                t -= 32;
            }
            if (lane == 0) st_relaxed_gpu64(desc + tile, kScanInclusive | (uint32_t)(excl + s_total));
        }
        if (lane == 0) s_prefix = excl;
    }
    __syncthreads();
    const uint32_t offset = s_prefix + thread_excl;
    if (full) {
        uint4 *dst = reinterpret_cast<uint4 *>(out + wbase) + lane;
#pragma unroll
        for (int q = 0; q < kScanVecs; ++q)
            dst[q * 32] = make_uint4(v[4 * q] + offset, v[4 * q + 1] + offset, v[4 * q + 2] + offset, v[4 * q + 3] + offset);
    } else {
#pragma unroll
        for (int q = 0; q < kScanVecs; ++q)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint64_t i = wbase + (uint64_t)q * 128 + lane * 4 + e;
                if (i < n) out[i] = v[4 * q + e] + offset;
            }
    }
}

// ---- persistent form (round 2) ------------------------------------------------------------------------------
// One tile per CTA leaves a slot without loads in flight while its tile scans and looks back (every geometry
// of the kernel above lands at 53-59 % of the HBM peak).  Here a resident grid of CTAs takes tiles from an atomic
// ticket (monotone, so a tile's predecessors are always running or done) and loads tile k+1 into a second
// register set BEFORE it scans, publishes, looks back and stores tile k: loads are in flight all the time.
// Tiles are large (THREADS x VECS x 4 elements): the look-back resolves at most 32 tiles per L2 round trip.
// desc: one zeroed uint64 per tile, followed by one zeroed uint32 ticket counter (at desc[num_tiles]).
template <int kScanThreads, int kScanVecs>
__global__ void __launch_bounds__(kScanThreads) exclusive_scan_persistent_kernel(const uint32_t *in, uint32_t *out, uint64_t n,
                                                                                  uint64_t *desc, uint32_t num_tiles) {
    constexpr int kScanTile = kScanThreads * kScanVecs * 4;
    __shared__ uint32_t s_warp_tot[32];
    __shared__ uint32_t s_prefix;
    __shared__ uint32_t s_total;
    __shared__ uint32_t s_ticket;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t *ticket = reinterpret_cast<uint32_t *>(desc + num_tiles);

    auto load_tile = [&](uint32_t tile, uint32_t (&v)[kScanVecs * 4]) {
        const uint64_t base = (uint64_t)tile * kScanTile;
        const uint64_t wbase = base + (uint64_t)warp * (kScanVecs * 128);
        if (base + kScanTile <= n) {
            const uint4 *src = reinterpret_cast<const uint4 *>(in + wbase) + lane;
#pragma unroll
            for (int q = 0; q < kScanVecs; ++q) {
                const uint4 x = ld_stream_v4(src + q * 32);
                v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < kScanVecs; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint64_t i = wbase + (uint64_t)q * 128 + lane * 4 + e;
                    v[4 * q + e] = (i < n) ? in[i] : 0u;
                }
        }
    };

    if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    uint32_t tile = s_ticket;
    uint32_t nxt[kScanVecs * 4];
    if (tile < num_tiles) load_tile(tile, nxt);
    while (tile < num_tiles) {
        uint32_t v[kScanVecs * 4];
#pragma unroll
        for (int k = 0; k < kScanVecs * 4; ++k) v[k] = nxt[k];
        __syncthreads();  // everybody has read s_ticket / s_prefix of the previous round
        if (tid == 0) s_ticket = atomicAdd(ticket, 1u);
        // exclusive scan inside the warp's slice: within the vector, across lanes, across rounds
        uint32_t sum = 0;
#pragma unroll
        for (int q = 0; q < kScanVecs; ++q) {
            uint32_t t = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t x = v[4 * q + e];
                v[4 * q + e] = t;
                t += x;
            }
            uint32_t incl = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            const uint32_t lane_excl = sum + incl - t;
#pragma unroll
            for (int e = 0; e < 4; ++e) v[4 * q + e] += lane_excl;
            sum += __shfl_sync(0xffffffffu, incl, 31);
        }
        const uint32_t warp_excl_all = block_exclusive_scan<kScanThreads>(lane == 0 ? sum : 0u, s_warp_tot);
        const uint32_t thread_excl = __shfl_sync(0xffffffffu, warp_excl_all, 0);
        if (tid == kScanThreads - 32) {
            const uint32_t total = thread_excl + sum;
            st_relaxed_gpu64(desc + tile, (tile == 0 ? kScanInclusive : kScanAggregate) | total);
            s_total = total;
        }
        // the next tile's loads go out before anybody waits on the look-back (s_ticket was written before the
        // barriers inside block_exclusive_scan)
        const uint32_t next_tile = s_ticket;
        if (next_tile < num_tiles) load_tile(next_tile, nxt);
        __syncthreads();
        if (tid < 32) {
            uint32_t excl = 0;
            if (tile != 0) {
                int64_t t = (int64_t)tile - 1;
                for (;;) {
                    const int64_t mine = t - (int64_t)lane;
                    uint64_t d = kScanInclusive;
                    uint32_t incl_mask, ready_mask, upto;
                    for (;;) {
                        if (mine >= 0) d = ld_relaxed_gpu64(desc + mine);
                        incl_mask = __ballot_sync(0xffffffffu, (d >> 32) == 2u);
                        ready_mask = __ballot_sync(0xffffffffu, (d >> 32) != 0u);
                        upto = incl_mask ? (uint32_t)__ffs(incl_mask) - 1u : 31u;
                        const uint32_t need = (upto == 31u) ? 0xffffffffu : ((2u << upto) - 1u);
                        if ((ready_mask & need) == need) break;
                    }
                    const uint32_t part = (lane <= upto) ? (uint32_t)d : 0u;
                    excl += __reduce_add_sync(0xffffffffu, part);
                    if (incl_mask) break;
                    t -= 32;
                }
                if (lane == 0) st_relaxed_gpu64(desc + tile, kScanInclusive | (uint32_t)(excl + s_total));
            }
            if (lane == 0) s_prefix = excl;
        }
        __syncthreads();
        const uint32_t offset = s_prefix + thread_excl;
        const uint64_t base = (uint64_t)tile * kScanTile;
        const uint64_t wbase = base + (uint64_t)warp * (kScanVecs * 128);
        if (base + kScanTile <= n) {
            uint4 *dst = reinterpret_cast<uint4 *>(out + wbase) + lane;
#pragma unroll
            for (int q = 0; q < kScanVecs; ++q)
                dst[q * 32] = make_uint4(v[4 * q] + offset, v[4 * q + 1] + offset, v[4 * q + 2] + offset, v[4 * q + 3] + offset);
        } else {
#pragma unroll
            for (int q = 0; q < kScanVecs; ++q)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint64_t i = wbase + (uint64_t)q * 128 + lane * 4 + e;
                    if (i < n) out[i] = v[4 * q + e] + offset;
                }
        }
        tile = next_tile;
    }
}

}  // namespace b200sort
