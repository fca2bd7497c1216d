// onesweep.cuh -- the digit-pass kernel: ONE kernel per digit that ranks a tile of keys,
// resolves the tile's global bin offsets with a single-pass decoupled look-back, and writes
// the keys (and values) to their final place for this digit.
//
// It replaces, per digit pass of the reference's sortByDevice loop
// (SourceCode/Parallel7.cu:561-622):
//   sortLocallyDataBlocks  (:193-251; numBits x [Blelloch scan kernel + 1-bit split kernel])
//   histogram              (:345-359; per-tile table through global atomics)
//   transpose/scan/transpose of the tile x bin table (:394-406, :485-528, host round trip)
//   scatter                (:306-316; 4-byte scattered stores)
// The tile x bin table never exists in memory: each tile publishes its 2^W bin counts as
// 32-bit descriptors {2 status bits | 30 value bits} and walks its predecessors'
// descriptors until it meets an inclusive prefix.  scan[t][d] of the reference
// (SourceCode/Baseline4.cu:127-138) == bin_base[d] + exclusive look-back prefix of (t, d).
//
// Stability: a tile covers TILE consecutive keys; warp w owns a contiguous slice of it and
// loads it warp-striped (item i of lane l = slice[i*32 + l]); items are ranked in increasing
// i, lanes inside a match.any group in increasing lane, warps in increasing w, tiles in
// increasing tile id (dynamic ids from an atomic ticket, so a tile's predecessors are always
// resident or finished -> the look-back cannot deadlock).
//
// Descriptor status codes rotate with the launch parity so a descriptor array is cleared
// once per sort, not once per pass: parity e uses NOT_READY = 2e, AGGREGATE = 2e+1,
// INCLUSIVE = 2e+2 (mod 4); every descriptor ends a launch as INCLUSIVE(e) == NOT_READY(e+1).
#pragma once
#include "common.cuh"

namespace b200sort {

template <int W, int THREADS, int ITEMS, bool PAIRS, bool DST>
struct PassTraits {
    static constexpr int B = 1 << W;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    // s_keys[TILE] | s_whist[WARPS][B] | s_gbase[B or 2B] | (DST && PAIRS: s_vbase[2B]) | s_warp_tot[32]
    static constexpr int SMEM_WORDS = TILE + WARPS * B + (DST ? 2 * B : B) + ((DST && PAIRS) ? 2 * B : 0) + 32;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_WORDS * 4;
};

template <int W, int THREADS, int ITEMS, int MIN_CTAS, bool PAIRS, bool DST>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) onesweep_pass_kernel(const PassArgs a) {
    using TR = PassTraits<W, THREADS, ITEMS, PAIRS, DST>;
    constexpr int B = TR::B;
    constexpr int WARPS = TR::WARPS;
    constexpr int TILE = TR::TILE;
    constexpr int WARP_KEYS = 32 * ITEMS;
    static_assert(B <= THREADS, "one thread per bin");
    static_assert(TILE < (1 << 16), "tile positions must fit 16 bits");

    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t *s_keys = smem;
    uint32_t *s_whist = s_keys + TILE;
    uint32_t *s_gbase = s_whist + WARPS * B;
    uint32_t *s_vbase = s_gbase + (DST ? 2 * B : B);
    uint32_t *s_warp_tot = s_vbase + ((DST && PAIRS) ? 2 * B : 0);
    __shared__ uint32_t s_tile;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    if (tid == 0) s_tile = atomicAdd(a.ticket, 1u);
    for (int i = tid; i < WARPS * B; i += THREADS) s_whist[i] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t n_valid = min((uint32_t)TILE, a.n - tile_base);
    const bool full = (n_valid == (uint32_t)TILE);

    // ---- load: warp-striped, every warp load instruction is one 128-byte line ----------
    uint32_t key[ITEMS];
    const uint32_t woff = warp * WARP_KEYS + lane;
    {
        const uint32_t *src = a.keys_in + tile_base + woff;
        if (full) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = ld_stream(src + i * 32);
        } else {
            // Out-of-range items become all-ones keys: they fall in the highest occupied bin,
            // after every real key of the tile, i.e. at tile positions >= n_valid.
#pragma unroll
            for (int i = 0; i < ITEMS; ++i)
                key[i] = (woff + i * 32 < n_valid) ? ld_stream(src + i * 32) : 0xFFFFFFFFu;
        }
    }

    // ---- rank inside the warp: match.any groups + a warp-private running histogram -----
    uint32_t rank[ITEMS];  // becomes the tile position of the key
    {
        uint32_t *wh = s_whist + warp * B;
        const uint32_t lt = lanemask_lt();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (key[i] >> a.shift) & a.mask;
            const uint32_t peers = __match_any_sync(0xffffffffu, d);
            const uint32_t below = peers & lt;
            uint32_t old = 0;
            if (below == 0) {  // lowest lane of the group owns the counter update
                old = wh[d];
                wh[d] = old + (uint32_t)__popc(peers);
            }
            __syncwarp();
            old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
            rank[i] = old + (uint32_t)__popc(below);
        }
    }

    uint32_t val[PAIRS ? ITEMS : 1];
    if (PAIRS) {
        // Issue the value loads now; they land while the block scans and looks back.
        const uint32_t *vsrc = a.vals_in + tile_base + woff;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i)
            val[i] = (full || woff + i * 32 < n_valid) ? ld_stream(vsrc + i * 32) : 0u;
    }
    __syncthreads();

    // ---- tile histogram = sum over warps; publish it as early as possible --------------
    uint32_t count = 0;
    if (tid < B) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) count += s_whist[w * B + tid];
    }
    const uint32_t st_not = ((2u * a.parity) & 3u) << 30;
    const uint32_t st_agg = ((2u * a.parity + 1u) & 3u) << 30;
    const uint32_t st_inc = ((2u * a.parity + 2u) & 3u) << 30;
    uint32_t *my_desc = a.desc + (size_t)tile * B + tid;
    if (tid < B) st_relaxed_gpu(my_desc, (tile == 0 ? st_inc : st_agg) | count);

    // ---- bin starts inside the tile, then per-(warp, bin) tile positions ----------------
    const uint32_t bin_start = block_exclusive_scan<THREADS>(count, s_warp_tot);
    if (tid < B) {
        uint32_t run = bin_start;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = s_whist[w * B + tid];
            s_whist[w * B + tid] = run;
            run += c;
        }
    }

    // ---- decoupled look-back: one thread per bin ----------------------------------------
    if (tid < B) {
        uint32_t excl = 0;
        if (tile != 0) {
            const uint32_t *p = my_desc - B;
            while (true) {
                uint32_t v;
                do {
                    v = ld_relaxed_gpu(p);
                } while ((v & kDescFlagMask) == st_not);
                excl += v & kDescValueMask;
                if ((v & kDescFlagMask) == st_inc) break;
                p -= B;
            }
            st_relaxed_gpu(my_desc, st_inc | (excl + count));
        }
        const uint32_t first = a.bin_base[tid] + excl;  // destination index of this tile's first key of bin tid
        if (a.carry_out != nullptr && tile == a.num_tiles - 1u) a.carry_out[tid] = first + count;
        if (!DST) {
            s_gbase[tid] = first - bin_start;  // mod 2^32; + tile position = destination index
        } else {
            const uint64_t delta = 4ull * (uint64_t)first - 4ull * (uint64_t)bin_start;  // mod 2^64
            reinterpret_cast<uint64_t *>(s_gbase)[tid] = a.bin_dst[tid] + delta;
            if (PAIRS) reinterpret_cast<uint64_t *>(s_vbase)[tid] = a.bin_dst[B + tid] + delta;
        }
    }
    __syncthreads();

    // ---- reorder through shared memory so each bin's keys are contiguous ----------------
    {
        const uint32_t *wh = s_whist + warp * B;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d = (key[i] >> a.shift) & a.mask;
            rank[i] += wh[d];
            s_keys[rank[i]] = key[i];
        }
    }
    __syncthreads();

    // ---- write out: consecutive threads -> consecutive tile positions -> (mostly)
    //      consecutive destination addresses inside a bin run ------------------------------
    uint32_t gpos[PAIRS ? ITEMS : 1];          // destination index (or offset) per written item
    uint32_t dpack[(PAIRS && DST) ? (ITEMS + 3) / 4 : 1];
    if (PAIRS && DST) {
#pragma unroll
        for (int q = 0; q < (ITEMS + 3) / 4; ++q) dpack[q] = 0;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const uint32_t j = tid + k * THREADS;
        if (full || j < n_valid) {
            const uint32_t kk = s_keys[j];
            const uint32_t d = (kk >> a.shift) & a.mask;
            if (!DST) {
                const uint32_t g = s_gbase[d] + j;
                a.keys_out[g] = kk;
                if (PAIRS) gpos[k] = g;
            } else {
                const uint64_t addr = reinterpret_cast<const uint64_t *>(s_gbase)[d] + 4ull * j;
                *reinterpret_cast<uint32_t *>(addr) = kk;
                if (PAIRS) dpack[k >> 2] |= d << (8 * (k & 3));
            }
        }
    }

    if (PAIRS) {
        __syncthreads();  // all keys read back from s_keys
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) s_keys[rank[i]] = val[i];
        __syncthreads();
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t j = tid + k * THREADS;
            if (full || j < n_valid) {
                const uint32_t vv = s_keys[j];
                if (!DST) {
                    a.vals_out[gpos[k]] = vv;
                } else {
                    const uint32_t d = (dpack[k >> 2] >> (8 * (k & 3))) & 0xFFu;
                    const uint64_t addr = reinterpret_cast<const uint64_t *>(s_vbase)[d] + 4ull * j;
                    *reinterpret_cast<uint32_t *>(addr) = vv;
                }
            }
        }
    }
}

}  // namespace b200sort
