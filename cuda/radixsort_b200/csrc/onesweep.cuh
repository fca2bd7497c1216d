// onesweep.cuh -- the digit-pass kernel: ONE kernel per digit that counts and ranks a tile of
// keys, resolves the tile's global bin offsets with a single-pass decoupled look-back, and
// writes the keys (and values) to their final place for this digit.
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
// Tile schedule (B200 leaves ~16 SM cycles per 32 keys per pass at 70 % of HBM bandwidth, so
// the design minimises shared-memory wavefronts and issue slots):
//   1. load      warp-striped LDG: warp w owns a contiguous slice, item i of lane l =
//                slice[i*32 + l]; every warp load is one 128-byte line
//   2. count     s_cnt[warp][digit] += 1 with shared atomics (no return value)
//   3. offsets   thread d: tile count of bin d = sum over warps -> publish the AGGREGATE
//                descriptor; block scan -> bin start in the tile; s_cnt[w][d] becomes the
//                shared-memory ADDRESS of the first key of (warp w, bin d)
//   4. rank      address of every key in stable order (three interchangeable modes, below);
//                the key is stored there, so each bin's keys end up contiguous
//   5. look-back thread d walks the predecessors' descriptors of bin d, LB at a time (8), then
//                publishes the INCLUSIVE descriptor.  Placed after (4) so that predecessors
//                have had the whole rank phase to publish.
//   6. write     thread t copies tile positions t, t+THREADS, ...: inside a bin run consecutive
//                threads write consecutive destination words
//
// Stable order inside (warp, digit): item index, then lane.  Rank modes:
//   RANK_TABLE   peers of a lane = lanes of the same warp instruction with the same digit, found
//                with one shared atomicOr + one load on a small per-warp table indexed by the
//                digit's high TB bits (32 entries -> bank-conflict free) AND'ed with ballots on
//                the remaining low bits; the highest peer fetches the run's base with ONE
//                shared atomicAdd(count) (distinct addresses -> fully defined) and broadcasts it.
//                With TB = 0 the table disappears and the ballots alone find the peers: the mode
//                used for digits of <= 3 bits (2..8 bins), where almost every lane has peers
//                and same-address atomics with a return value would serialise (measured: 18
//                cycles per warp instruction at 2 distinct addresses, 32 at one).
//   RANK_ATOMIC  address = atomicAdd(&s_cnt[warp][digit], 4) by every lane.  Needs same-address
//                shared atomics of one warp instruction to be applied in ascending lane order;
//                PTX does not promise that, so the host only selects this mode after the
//                on-device self test (atomic_order_selftest) passes.
//   RANK_MATCH   match.any.sync peers (the textbook form).  Kept for the record: MATCH.ANY
//                issues at ~1 warp instruction / 61 cycles / SM on B200, 2.4 ms per pass.
//
// Tile id = blockIdx.x: CTAs are dispatched in increasing block index, so a tile's predecessors
// are always resident or finished when it spins on them (the same assumption CUB's decoupled
// look-back scan makes); this saves one global atomic round trip at the head of every tile.
// Descriptor status codes rotate with the
// launch parity so the descriptor array is cleared once per sort, not once per pass: parity e
// uses NOT_READY = 2e, AGGREGATE = 2e+1, INCLUSIVE = 2e+2 (mod 4); every descriptor ends a
// launch as INCLUSIVE(e) == NOT_READY(e+1).
#pragma once
#include "common.cuh"

namespace b200sort {

enum RankMode { RANK_TABLE = 0, RANK_ATOMIC = 1, RANK_MATCH = 2 };

// ---- shared-memory accesses by 32-bit shared address (keeps address arithmetic to one LOP3) ----
__device__ __forceinline__ void sm_inc(uint32_t addr) {
    asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr) : "memory");
}
// Lanes OR disjoint bits into a cleared word, so an add is equivalent -- and red.shared.add combines
// the lanes of one instruction that hit the same address, which the or form may serialise.
__device__ __forceinline__ void sm_or(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t sm_add_ret(uint32_t addr, uint32_t v) {
    uint32_t r;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(r) : "r"(addr), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t sm_ld(uint32_t addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}
template <int OFF>
__device__ __forceinline__ void sm_st(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%1], %2;" ::"r"(addr), "n"(OFF), "r"(v) : "memory");
}

__device__ __forceinline__ void sm_st2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint2 sm_ld2(uint32_t addr) {
    uint2 r;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "r"(addr) : "memory");
    return r;
}

// v + (a value that is zero at run time but that neither nvcc nor ptxas can prove to be zero).
// Used on the rotate amount of each phase: the phases then compute their digit offsets from
// "different" inputs, so ptxas recomputes them (2 instructions per key) instead of keeping ITEMS
// of them alive across barriers and branches -- which it does by spilling to local memory, i.e.
// by adding L1 wavefronts to a kernel that is bound by exactly those.  (A mov-based or empty inline
// asm is folded away by ptxas and does not help.)
__device__ __forceinline__ uint32_t launder(uint32_t v, uint32_t runtime_zero) { return v + runtime_zero; }

// Byte offset of a key's table entry: the digit, shifted to a 4-byte stride, with its high nibble
// XOR-folded into its low nibble.  The fold is a bijection on [0, 2^W) and spreads digit values
// that differ only in their high bits -- keys that are multiples of 16, sorted input -- over all
// 32 shared-memory banks instead of two (measured: 1.43 ms -> see profiles for a sorted pass 0).
// Uniform digits are unaffected.  Used for the per-warp counter tables (count and rank phases);
// the write-out's base table is read by runs of lanes with the same digit (broadcast), so it
// keeps the plain index.
__device__ __forceinline__ uint32_t digit_slot(uint32_t key, uint32_t rot, uint32_t mask4) {
    const uint32_t r = __funnelshift_r(key, key, rot) & mask4;
    return r ^ ((r >> 4) & 0x3Cu);
}
__device__ __forceinline__ uint32_t fold_bin(uint32_t bin) { return bin ^ (bin >> 4); }

template <int W, int THREADS, int ITEMS, int MODE, int TB, bool PAIRS, bool DST>
struct PassTraits {
    static constexpr int B = 1 << W;
    static constexpr int WARPS = THREADS / 32;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int NBALLOT = (MODE == RANK_TABLE && W > TB) ? W - TB : 0;
    static constexpr int TABLE = (MODE == RANK_TABLE) ? (1 << (W - NBALLOT)) : 0;  // entries per mask table
    // word offsets inside dynamic shared memory (base is 1024-byte aligned; every table starts
    // on a multiple of its own size so `base | offset` replaces `base + offset`)
    static constexpr int OFF_KEYS = 0;   // keys: uint32[TILE]; pairs: {key, value}[TILE] (8-byte slots)
    static constexpr int OFF_VALS = TILE;  // second half of the pair region (staging of a ragged tile only)
    static constexpr int OFF_CNT = OFF_VALS + (PAIRS ? TILE : 0);        // [WARPS][B]
    static constexpr int OFF_GBASE = OFF_CNT + WARPS * B;                 // [B] or [B] x 64 bit
    static constexpr int OFF_VBASE = OFF_GBASE + (DST ? 2 * B : B);       // [B] x 64 bit (DST pairs)
    static constexpr int OFF_MASK = OFF_VBASE + ((DST && PAIRS) ? 2 * B : 0);  // [WARPS][2][TABLE]
    static constexpr int OFF_HOT = (OFF_MASK + WARPS * 2 * TABLE + 3) / 4 * 4;  // [WARPS] 1 = clustered warp (16-byte aligned)
    static constexpr int OFF_MISC = OFF_HOT + ((WARPS + 3) / 4) * 4;      // warp totals[32]
    static constexpr int SMEM_WORDS = OFF_MISC + 36;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_WORDS * 4;
    static_assert(TILE % B == 0, "tables must stay aligned to their size");
};

constexpr int kGroup = 6;  // shared-memory operations in flight per thread in the rank / write-out loops

template <int W, int THREADS, int ITEMS, int MIN_CTAS, int MODE, int TB, int LB, bool PERSIST, bool TMA, bool PAIRS, bool DST>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) onesweep_pass_kernel(const PassArgs a) {
    constexpr int kLookbackBatch = LB;  // descriptors in flight per bin thread during the look-back
    using TR = PassTraits<W, THREADS, ITEMS, MODE, TB, PAIRS, DST>;
    constexpr int B = TR::B;
    constexpr int WARPS = TR::WARPS;
    constexpr int TILE = TR::TILE;
    constexpr int WARP_KEYS = 32 * ITEMS;
    constexpr int NBALLOT = TR::NBALLOT;
    constexpr int TABLE = TR::TABLE;
    static_assert(B <= THREADS, "one thread per bin");
    static_assert(TILE < (1 << 16), "tile positions must fit 16 bits");
    constexpr uint32_t kSlot = PAIRS ? 8u : 4u;  // bytes per reordered element in shared memory

    extern __shared__ __align__(1024) uint32_t smem[];
    uint32_t *s_keys = smem + TR::OFF_KEYS;
    uint32_t *s_vals = smem + TR::OFF_VALS;
    uint32_t *s_cnt = smem + TR::OFF_CNT;
    uint32_t *s_gbase = smem + TR::OFF_GBASE;
    uint32_t *s_vbase = smem + TR::OFF_VBASE;
    uint32_t *s_warp_tot = smem + TR::OFF_MISC;
    uint32_t *s_hot = smem + TR::OFF_HOT;

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t sa_keys = smem_u32(s_keys);
    const uint32_t sa_wcnt = smem_u32(s_cnt + warp * B);  // multiple of 4*B bytes

    // byte offset of a key's digit inside a 4-byte-entry table: (rotr(key, shift-2) & mask4).
    // Every phase uses its own laundered copy of the rotate amount (see launder()).
    const uint32_t rot = (a.shift + 30u) & 31u;
    const uint32_t rot_fast = launder(rot, a.parity >> 8);
    const uint32_t rot_clustered = launder(rot, a.parity >> 9);
    const uint32_t mask4 = a.mask << 2;

    // ---- 1. load ----------------------------------------------------------------------------
    uint32_t key[ITEMS];
    uint32_t val[PAIRS ? ITEMS : 1];
    const uint32_t woff = warp * WARP_KEYS + lane;
    // TMA: the tile is fetched by the bulk-copy engine (cp.async.bulk -> SASS UBLKCP) into the
    // reorder buffer and read from there, so the 350 key-load requests of a tile do not sit in the
    // SM's load/store queue in front of other CTAs' look-back reads and stores.  Needs 16-byte
    // aligned inputs; otherwise (and for a ragged tile) the LDG paths are used.
    __shared__ __align__(8) uint64_t s_bar;
    const bool use_tma = TMA && !PERSIST && ((reinterpret_cast<uintptr_t>(a.keys_in) & 15u) == 0) &&
                         (!PAIRS || (reinterpret_cast<uintptr_t>(a.vals_in) & 15u) == 0);
    auto load_tma_issue = [&](uint32_t t) {  // thread 0 only
        constexpr uint32_t kParts = 8, kPartBytes = TILE * 4 / kParts;
        static_assert((TILE * 4) % (kParts * 16) == 0, "bulk copies move multiples of 16 bytes");
        mbar_init(&s_bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(&s_bar, (PAIRS ? 2u : 1u) * TILE * 4u);
#pragma unroll
        for (uint32_t q = 0; q < kParts; ++q) {
            bulk_g2s(reinterpret_cast<char *>(s_keys) + q * kPartBytes,
                     reinterpret_cast<const char *>(a.keys_in + t * (uint32_t)TILE) + q * kPartBytes, kPartBytes, &s_bar);
            if (PAIRS)
                bulk_g2s(reinterpret_cast<char *>(s_vals) + q * kPartBytes,
                         reinterpret_cast<const char *>(a.vals_in + t * (uint32_t)TILE) + q * kPartBytes, kPartBytes, &s_bar);
        }
    };
    auto load_full = [&](uint32_t t) {
        const uint32_t *src = a.keys_in + t * (uint32_t)TILE + woff;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) key[i] = ld_stream(src + i * 32);
        if (PAIRS) {
            const uint32_t *vsrc = a.vals_in + t * (uint32_t)TILE + woff;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) val[i] = ld_stream(vsrc + i * 32);
        }
    };
    // Ragged last tile (once per launch): staged through shared memory so that the path above
    // stays free of per-item predicates.  Out-of-range items become all-ones keys: they fall in
    // the highest occupied bin, after every real key of the tile, i.e. at tile positions
    // >= n_valid, and are never written out.  Must be called by the whole CTA.
    auto load_staged = [&](uint32_t t, uint32_t nv) {
        const uint32_t tb = t * (uint32_t)TILE;
#pragma unroll 1
        for (uint32_t j = tid; j < (uint32_t)TILE; j += THREADS) {
            s_keys[j] = (j < nv) ? a.keys_in[tb + j] : 0xFFFFFFFFu;
            if (PAIRS) s_vals[j] = (j < nv) ? a.vals_in[tb + j] : 0u;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            key[i] = s_keys[woff + i * 32];
            if (PAIRS) val[i] = s_vals[woff + i * 32];
        }
        __syncthreads();
    };

    // PERSIST: the CTA takes tiles from an atomic ticket (dynamic order keeps the tiles in flight
    // staggered, which the look-back needs; a static stride would start a whole round of tiles in
    // lockstep) and issues the loads of its next tile as soon as the rank phase has emptied the key
    // registers, so they are in flight during the look-back and the write-out of the current tile.
    // The ticket for the next tile is requested at the top of an iteration and published to the CTA
    // just before the barrier that ends the offsets phase.
    uint32_t *s_next = smem + TR::OFF_MISC + 32;
    uint32_t tile = blockIdx.x;
    if (PERSIST) {
        if (tid == 0) *s_next = atomicAdd(a.ticket, 1u);
        __syncthreads();
        tile = *s_next;
        if (tile >= a.num_tiles) return;  // more CTAs than tiles
        __syncthreads();
    }
    bool tma_pending = false;
    {
        const uint32_t nv = min((uint32_t)TILE, a.n - tile * (uint32_t)TILE);
        if (nv != (uint32_t)TILE) {
            load_staged(tile, nv);
        } else if (use_tma) {
            if (tid == 0) load_tma_issue(tile);
            tma_pending = true;
        } else {
            load_full(tile);
        }
    }
    for (;;) {
    uint32_t ticket = 0;
    if (PERSIST && tid == 0) ticket = atomicAdd(a.ticket, 1u);  // consumed in step 3
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t n_valid = min((uint32_t)TILE, a.n - tile_base);
    const bool full = (n_valid == (uint32_t)TILE);
    {
        // s_cnt .. s_mask are contiguous and 16-byte aligned; rounding the vector count up spills
        // at most 3 words into s_warp_tot, which is written before it is read.
        uint4 *z = reinterpret_cast<uint4 *>(s_cnt);
        constexpr int ZV = (TR::OFF_MISC - TR::OFF_CNT + 3) / 4;
        for (int i = tid; i < ZV; i += THREADS) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
    if (TMA && tma_pending) {  // block-uniform
        mbar_wait(&s_bar, 0);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            key[i] = s_keys[woff + i * 32];
            if (PAIRS) val[i] = s_vals[woff + i * 32];
        }
        tma_pending = false;  // s_keys is rewritten two barriers later (rank phase)
    }

    // ---- 2. count ---------------------------------------------------------------------------
    // A warp is "clustered" when neighbouring lanes of its warp instructions mostly share digits
    // (sorted or partially sorted input, few distinct keys, data grouped by earlier passes); two
    // sample items decide.  Clustered warps rank runs of equal digits with one atomic (step 4) and
    // use plain table slots -- they touch few distinct entries per instruction, so bank spreading
    // (digit_slot) would only cost them instructions.
    bool clustered = false;
    if (MODE == RANK_ATOMIC) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t dq = __funnelshift_r(key[q * (ITEMS / 2)], key[q * (ITEMS / 2)], rot) & mask4;
            const uint32_t heads = __ballot_sync(0xffffffffu, dq != __shfl_up_sync(0xffffffffu, dq, 1));
            clustered = clustered || __popc(heads) <= 16;
        }
        if (lane == 0) s_hot[warp] = clustered ? 1u : 0u;
    }
    if (MODE == RANK_ATOMIC && clustered) {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) sm_inc(sa_wcnt | (__funnelshift_r(key[i], key[i], rot) & mask4));
    } else {
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) sm_inc(sa_wcnt | digit_slot(key[i], rot, mask4));
    }
    __syncthreads();

    // ---- 3. offsets -------------------------------------------------------------------------
    uint32_t count = 0;
    uint32_t c[WARPS];
    const uint32_t folded = fold_bin(tid);  // table entry of bin `tid` in warps that use digit_slot
    uint32_t plain_mask = 0;                // bit w: warp w is clustered and uses plain slots
    if (MODE == RANK_ATOMIC && tid < B) {
#pragma unroll
        for (int w4 = 0; w4 < (WARPS + 3) / 4; ++w4) {
            const uint4 f = reinterpret_cast<const uint4 *>(s_hot)[w4];  // broadcast loads
            plain_mask |= (f.x | (f.y << 1) | (f.z << 2) | (f.w << 3)) << (4 * w4);
        }
    }
    if (tid < B) {
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            c[w] = s_cnt[w * B + ((plain_mask >> w & 1u) ? tid : folded)];
            count += c[w];
        }
    }
    const uint32_t st_not = ((2u * a.parity) & 3u) << 30;
    const uint32_t st_agg = ((2u * a.parity + 1u) & 3u) << 30;
    const uint32_t st_inc = ((2u * a.parity + 2u) & 3u) << 30;
    uint32_t *my_desc = a.desc + (size_t)tile * B + tid;
    if (tid < B) st_relaxed_gpu(my_desc, (tile == 0 ? st_inc : st_agg) | count);

    const uint32_t bin_start = block_exclusive_scan<THREADS>(count, s_warp_tot);
    if (tid < B) {
        // RANK_ATOMIC keeps ready-to-use shared addresses; the other modes keep tile positions
        uint32_t run = (MODE == RANK_ATOMIC) ? sa_keys + kSlot * bin_start : bin_start;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            s_cnt[w * B + ((plain_mask >> w & 1u) ? tid : folded)] = run;
            run += (MODE == RANK_ATOMIC) ? kSlot * c[w] : c[w];
        }
    }
    if (PERSIST && tid == 0) *s_next = ticket;
    __syncthreads();

    // ---- 4. rank + reorder through shared memory ----------------------------------------------
    if (MODE == RANK_ATOMIC) {
        const uint32_t lt = lanemask_lt();
        if (!clustered) {
            // Software-pipelined in groups: kGroup atomics in flight before their dependent stores,
            // so a warp pays one shared-memory round trip per group instead of one per key.
#pragma unroll
            for (int i0 = 0; i0 < ITEMS; i0 += kGroup) {
                uint32_t at[kGroup];
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
                    if (i0 + g < ITEMS)
                        at[g] = sm_add_ret(sa_wcnt | digit_slot(key[i0 + g], rot_fast, mask4), kSlot);
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
                    if (i0 + g < ITEMS) {
                        if (PAIRS) sm_st2(at[g], key[i0 + g], val[i0 + g]);
                        else sm_st<0>(at[g], key[i0 + g]);
                    }
            }
        } else {
            // atom.shared.add with a return value serialises lanes that share an address (measured:
            // 32 cycles per instruction when all 32 do), so every RUN of equal digits among
            // consecutive lanes is ranked by its first lane with ONE atomic of the run length.
            const uint32_t lanebit = 1u << lane;
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const uint32_t k = key[i];
                const uint32_t d4 = __funnelshift_r(k, k, rot_clustered) & mask4;
                const uint32_t prev = __shfl_up_sync(0xffffffffu, d4, 1);
                const bool head = (lane == 0u) || (d4 != prev);
                const uint32_t hm = __ballot_sync(0xffffffffu, head);        // bit 0 is always set
                const uint32_t start = 31u - (uint32_t)__clz(hm & (lt | lanebit));  // first lane of my run
                const uint32_t above = hm & ~(lt | lanebit);                  // run heads above me
                const uint32_t end = above ? (uint32_t)__ffs(above) - 1u : 32u;
                uint32_t at = 0;
                if (head) at = sm_add_ret(sa_wcnt | d4, kSlot * (end - lane));
                at = __shfl_sync(0xffffffffu, at, start) + kSlot * (lane - start);
                if (PAIRS) sm_st2(at, k, val[i]);
                else sm_st<0>(at, k);
            }
        }
    } else {
        const uint32_t lt = lanemask_lt();
        const uint32_t lanebit = 1u << lane;
        const uint32_t sa_wmask = smem_u32(smem + TR::OFF_MASK + warp * 2 * TABLE);
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const uint32_t d4 = digit_slot(key[i], rot_fast, mask4);
            uint32_t peers;
            if (MODE == RANK_MATCH) {
                peers = __match_any_sync(0xffffffffu, d4);
            } else {
                // two tables alternate so that clearing one never races with filling the next
                const uint32_t row = sa_wmask + (uint32_t)(i & 1) * (TABLE * 4);
                const uint32_t slot = row + ((d4 >> NBALLOT) & ~3u);
                if (TABLE > 1) {
                    sm_or(slot, lanebit);
                    __syncwarp();
                    peers = sm_ld(slot);
                } else {
                    peers = 0xffffffffu;  // narrow digit: ballots alone identify the peers
                }
#pragma unroll
                for (int b = 0; b < NBALLOT; ++b) {
                    const bool bit = (d4 & (4u << b)) != 0u;
                    const uint32_t bal = __ballot_sync(0xffffffffu, bit);
                    peers &= bit ? bal : ~bal;
                }
                if (TABLE > 1) {
                    __syncwarp();
                    sm_st<0>(slot, 0u);
                }
            }
            uint32_t base = 0;
            if ((peers >> lane) == 1u)  // highest lane of the group
                base = sm_add_ret(sa_wcnt | d4, (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, 31 - __clz(peers));
            const uint32_t at = sa_keys + kSlot * (base + (uint32_t)__popc(peers & lt));
            if (PAIRS) sm_st2(at, key[i], val[i]);
            else sm_st<0>(at, key[i]);
        }
    }

    // ---- next tile's loads (PERSIST): the key registers are free from here on.  Measured on B200
    // (profiles/r01_sweep_v8_persistent.jsonl): 0.69 ms per pass against 0.64 for one tile per CTA --
    // the 350 load requests queue in front of the look-back's descriptor reads; issuing them after
    // the look-back instead (overlapping only the write-out) is worse still (1.03 ms).  The
    // persistent variants are kept as measured experiments, not as the default.
    uint32_t next_tile = tile;
    bool has_next = false, next_full = false;
    if (PERSIST) {
        next_tile = *s_next;
        has_next = next_tile < a.num_tiles;
        next_full = has_next && (a.n - next_tile * (uint32_t)TILE >= (uint32_t)TILE);
        if (next_full) load_full(next_tile);
    }

    // ---- 5. decoupled look-back: one thread per bin, kLookbackBatch descriptors in flight ------
    if (tid < B) {
        uint32_t excl = 0;
        if (tile != 0) {
            const uint32_t *col = a.desc + tid;
            int32_t t = (int32_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[kLookbackBatch];
#pragma unroll
                for (int k = 0; k < kLookbackBatch; ++k)
                    v[k] = (t - k >= 0) ? ld_relaxed_gpu(col + (size_t)(t - k) * B) : st_inc;
                bool stop = false;
#pragma unroll
                for (int k = 0; k < kLookbackBatch; ++k) {
                    const uint32_t f = v[k] & kDescFlagMask;
                    if (!stop && f == st_not) stop = true;  // not published yet: poll again from here
                    if (!stop) {
                        excl += v[k] & kDescValueMask;
                        --t;
                        if (f == st_inc) { stop = true; done = true; }
                    }
                }
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

    // ---- 6. write out -------------------------------------------------------------------------
    const uint32_t sa_gbase = smem_u32(s_gbase);
    uint32_t *const kout = a.keys_out;
    uint32_t *const vout = a.vals_out;
    if (full) {
#pragma unroll
        for (int k0 = 0; k0 < ITEMS; k0 += kGroup) {
            uint32_t kk[kGroup], vv[PAIRS ? kGroup : 1], gb[kGroup];
#pragma unroll
            for (int g = 0; g < kGroup; ++g)
                if (k0 + g < ITEMS) {
                    const uint32_t j = tid + (k0 + g) * THREADS;
                    if (PAIRS) {
                        const uint2 kv = sm_ld2(sa_keys + 8u * j);
                        kk[g] = kv.x;
                        vv[g] = kv.y;
                    } else {
                        kk[g] = sm_ld(sa_keys + 4u * j);
                    }
                }
            if (!DST) {
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
                    if (k0 + g < ITEMS) gb[g] = sm_ld(sa_gbase | (__funnelshift_r(kk[g], kk[g], rot) & mask4));
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
                    if (k0 + g < ITEMS) {
                        const uint32_t j = tid + (k0 + g) * THREADS;
                        kout[gb[g] + j] = kk[g];
                        if (PAIRS) vout[gb[g] + j] = vv[g];
                    }
            } else {
#pragma unroll
                for (int g = 0; g < kGroup; ++g)
                    if (k0 + g < ITEMS) {
                        const uint32_t j = tid + (k0 + g) * THREADS;
                        const uint32_t d = (__funnelshift_r(kk[g], kk[g], rot) & mask4) >> 2;
                        const uint64_t off = 4ull * j;
                        *reinterpret_cast<uint32_t *>(reinterpret_cast<const uint64_t *>(s_gbase)[d] + off) = kk[g];
                        if (PAIRS)
                            *reinterpret_cast<uint32_t *>(reinterpret_cast<const uint64_t *>(s_vbase)[d] + off) = vv[g];
                    }
            }
        }
    } else {
#pragma unroll 2
        for (int k = 0; k < ITEMS; ++k) {
            const uint32_t j = tid + k * THREADS;
            if (j < n_valid) {
                uint32_t kk, vv = 0;
                if (PAIRS) {
                    const uint2 kv = sm_ld2(sa_keys + 8u * j);
                    kk = kv.x;
                    vv = kv.y;
                } else {
                    kk = sm_ld(sa_keys + 4u * j);
                }
                const uint32_t d4 = __funnelshift_r(kk, kk, rot) & mask4;
                if (!DST) {
                    const uint32_t g = sm_ld(sa_gbase | d4) + j;
                    kout[g] = kk;
                    if (PAIRS) vout[g] = vv;
                } else {
                    const uint64_t off = 4ull * j;
                    *reinterpret_cast<uint32_t *>(reinterpret_cast<const uint64_t *>(s_gbase)[d4 >> 2] + off) = kk;
                    if (PAIRS)
                        *reinterpret_cast<uint32_t *>(reinterpret_cast<const uint64_t *>(s_vbase)[d4 >> 2] + off) = vv;
                }
            }
        }
    }

    if (!PERSIST || !has_next) break;
    __syncthreads();  // the write-out has finished reading s_keys / s_gbase before they are reused
    if (!next_full) load_staged(next_tile, a.n - next_tile * (uint32_t)TILE);
    tile = next_tile;
    }  // tile loop
}

// ---- self test for RANK_ATOMIC ------------------------------------------------------------------
// RANK_ATOMIC needs atom.shared.add lanes of ONE warp instruction that hit the same address to be applied
// in ascending lane order.  PTX does not promise that, so every device is tested before the mode is
// used on it, under the conditions of the real kernel: 8 warps per CTA and three CTAs per SM; the even
// warps replay `rounds` random digit patterns (1 to 256 distinct addresses per instruction) with a
// DIFFERENT addend per lane (the clustered path adds run lengths) against a private table and check
// that every lane got back the sum of the addends of all earlier instructions plus those of the LOWER
// lanes of the same instruction with the same digit; the odd warps meanwhile hammer the banks of the
// neighbouring table with reductions, stores and loads (the count, reorder and write-out traffic of
// other warps).  mismatches[0] counts violations.  ~3 ms, once per device and process.
template <int UNUSED>  // template only so the header can be included in several translation units
__global__ void __launch_bounds__(256, 3) atomic_order_selftest(uint32_t *mismatches, int rounds, uint32_t seed) {
    __shared__ __align__(1024) uint32_t table[8][256];
    __shared__ uint32_t shadow[4][256];
    __shared__ uint32_t noise[8 * 1024];  // 32 KB more per CTA: three CTAs per SM like the digit-pass kernel
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 256; i += 256) (&table[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 4 * 256; i += 256) (&shadow[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 8 * 1024; i += 256) noise[i] = 0;
    __syncthreads();
    uint32_t x = seed ^ (blockIdx.x * 2654435761u) ^ (threadIdx.x * 40503u);
    if (warp & 1u) {
        // interference: random-bank traffic of all three kinds next to the tables under test
        const uint32_t sa_noise = smem_u32(noise), sa_next = smem_u32(&table[warp][0]);
        uint32_t acc = 0;
        for (int r = 0; r < rounds; ++r) {
            x = x * 1664525u + 1013904223u;
            const uint32_t o = (x >> 9) & 0x7FFCu;
            sm_inc(sa_next | ((x >> 20) & 0x3FCu));
            sm_st<0>(sa_noise + o, x);
            acc += sm_ld(sa_noise + ((o * 7u) & 0x7FFCu));
        }
        if (acc == 0x12345678u) atomicAdd(mismatches + 1, 1u);
        return;
    }
    const uint32_t sa_table = smem_u32(&table[warp][0]);
    uint32_t *my_shadow = shadow[warp >> 1];
    uint32_t bad = 0;
    const uint32_t lt = lanemask_lt();
    for (int r = 0; r < rounds; ++r) {
        x = x * 1664525u + 1013904223u;
        const uint32_t spread = (1u << (r % 9)) - 1u;  // 1, 2, 4, ..., 256 distinct values
        const uint32_t d = (x >> 13) & spread & 255u;
        const uint32_t add = 4u * (1u + ((x >> 24) & 31u));  // 4 .. 128, different per lane
        const uint32_t got = sm_add_ret(sa_table | (d << 2), add);
        __syncwarp();
        uint32_t peers = 0xffffffffu;  // reference rank by ballots over the 8 digit bits
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
        // sum of the addends of the lower peers: walk the peer mask (<= 32 steps, warp-uniform trip count)
        uint32_t below = 0, total = 0;
#pragma unroll 1
        for (uint32_t src = 0; src < 32u; ++src) {
            const uint32_t a_src = __shfl_sync(0xffffffffu, add, src);
            if ((peers >> src) & 1u) {
                total += a_src;
                if (src < lane) below += a_src;
            }
        }
        (void)lt;
        const uint32_t want = my_shadow[d] + below;
        __syncwarp();
        if ((peers >> lane) == 1u) my_shadow[d] += total;
        __syncwarp();
        if (got != want) ++bad;
        if ((r & 63) == 63) {  // keep counters small
            for (int i = lane; i < 256; i += 32) { table[warp][i] = 0; my_shadow[i] = 0; }
            __syncwarp();
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

}  // namespace b200sort
