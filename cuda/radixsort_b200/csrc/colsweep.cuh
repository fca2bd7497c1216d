// colsweep.cuh -- the digit-pass kernel, second generation ("column sweep"): ONE kernel per digit
// that counts and ranks a tile of keys with conflict-free, lane-private shared-memory atomics,
// resolves the tile's global bin offsets with a single-pass decoupled look-back, and writes the
// keys (and values) to their final place for this digit.
//
// It replaces, per digit pass of the reference's sortByDevice loop
// (SourceCode/Parallel7.cu:561-622): sortLocallyDataBlocks (:193-251), histogram (:345-359),
// transpose/scan/transpose (:394-406, :485-528, host round trip) and scatter (:306-316).
//
// Why a second design (numbers: profiles/r01_ncu_bench_2p28.json, profiles/r02_probe2_b200.json).
// B200 leaves ~16 SM cycles per 32 keys per pass at 70 % of HBM bandwidth, and the shared-memory
// pipe moves one 128-byte wavefront per cycle.  The first kernel (onesweep.cuh) ranks with per-warp
// 2^W-entry tables shared by the 32 lanes of a warp: 32 random digits on 32 banks cost 3.7 wavefronts
// per warp instruction (measured), twice per key (count + rank) -- 6.9 of its 13.7 shared wavefronts
// per 32 keys -- and the rank needs same-address atomics of one instruction to be applied in lane
// order, which PTX does not promise.  Here the counters are LANE-PRIVATE: word (row, lane) of a
// [2^W / 2][32] table holds two 16-bit counters (bins 2*row and 2*row+1) that only lane `lane` of any
// warp ever touches, so every table access of a warp instruction hits 32 different banks: 1.02
// cycles per warp instruction measured, for any key distribution, and no two lanes of one
// instruction ever share an address.
//
// The price is the order.  Lane l owns COLUMN l of the tile -- keys [l*COL, (l+1)*COL) -- and inside
// the column warp w owns keys [w*ITEMS, (w+1)*ITEMS), so thread (w, l) holds ITEMS consecutive keys
// (blocked arrangement; the tile is brought in by the bulk-copy engine and read with conflict-free
// 128-bit loads).  The stable order of a bin is (column, warp, item) = index order.  After the
// count, a scan turns counter (d, l) into the shared-memory position of the first key of bin d in
// column l; the ranking atomics of the warps then have to reach a counter in warp order, so the
// warps take turns: warp w waits on named barrier w, issues its ITEMS atomics, and arrives on
// barrier w+1 (bar.arrive/bar.sync order the accesses: spec-safe).  The hand-off costs ~60 cycles;
// a tile's chain is ~4-5 k cycles of a ~15 k cycle tile time, overlapped by the other CTAs of the SM.
//
// Tile schedule (THREADS = 32*WARPS, TILE = THREADS*ITEMS, COL = WARPS*ITEMS):
//   0. issue the bulk copy of the tile (cp.async.bulk -> UBLKCP), zero the counter table
//   1. load     ITEMS/4 x ld.shared.v4 per thread (quad stride COL/4 is odd: conflict-free)
//   2. count    red.shared.add on (row, lane): +1 in the low or high half
//   3. scan     4 threads per row: 8 columns each in registers, prefix by shuffles inside the
//               quad of lanes; thread d: tile count of bin d -> AGGREGATE descriptor; block scan ->
//               bin start; the table then holds byte positions
//   4. rank     warp chain: position = atom.shared.add(word, 4 << 16*half); key stored there
//   5. look-back thread d walks the predecessors' descriptors of bin d, LB at a time, then
//               publishes the INCLUSIVE descriptor
//   6. write    thread t copies tile positions t, t+THREADS, ...
//
// COL % 8 == 4 (WARPS odd, ITEMS % 8 == 4) makes the column stride an odd number of quads (step 1)
// and keeps the worst regular input -- all keys equal: lane l stores to position l*COL + r -- at a
// 4-way bank conflict, the same as random positions.
//
// The default form since round 2 (template flags WIDE + DUAL + LBV4, launch.h: kDualVariant) changes three things,
// each after a measurement (tools/gpu_probe3.cu, tools/col_timeline.py, profiles/r02_probe3_b200.json):
//   * A warp's shared-memory atomics WITH a return value complete one after the other -- 16.5 cycles each, whatever
//     else the SM does -- while those of different warps overlap perfectly.  One chain of turns therefore ranks one
//     key per lane per 16.5 cycles, as fast as the whole SM may take per 32 keys at 70 % of HBM bandwidth.  DUAL
//     splits the tile into two halves with their own column sets, chains and 16-bit counter halves of one 32-bit
//     word per (bin, lane) (row = digit: no per-key half select, constant addend); keys stay rotated from load to
//     write-out.  8 warps as 5 + 3, 60 keys per thread, 128 registers, two CTAs per SM.
//   * A warp issues one strong (.relaxed.gpu) load per ~55 cycles, so a look-back round of 2 bins x 4 rows per
//     thread was mostly issue time.  LBV4: 64 threads, four neighbouring bins each, 16-byte strong loads.
//   * The tile this SM slot runs next (PassArgs::prefetch tiles ahead) is requested into L2 when a tile starts.
// 0.636 -> 0.545 ms per pass for 2^28 keys (0.60 of measured HBM peak), 0.475 ms with 4-bit digits (0.69).
//
// Descriptor protocol, ragged last tile, per-bin destinations (DST) and the carry between launches
// of one pass are those of onesweep.cuh.
#pragma once
#include "common.cuh"
#include "onesweep.cuh"  // shared-memory access helpers

namespace b200sort {

__device__ __forceinline__ uint4 sm_ld4(uint32_t addr) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sm_st4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sm_red(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

#ifdef B200_COL_DEBUG
__device__ long long g_col_dbg[16 * 20];
#define COL_STAMP(k) do { if (blockIdx.x == a.num_tiles / 2 && (threadIdx.x & 31u) == 0) g_col_dbg[(threadIdx.x >> 5) * 20 + (k)] = clock64(); } while (0)
#else
#define COL_STAMP(k) do { } while (0)
#endif

template <int W, int WARPS, int ITEMS, bool PAIRS, bool DST, bool WIDE = false, bool DUAL = false>
struct ColTraits {
    static constexpr int B = 1 << W;
    static constexpr int ROWS = WIDE ? B : B / 2;  // packed: two bins per 32-bit counter word; WIDE: one
    static constexpr int THREADS = 32 * WARPS;
    static constexpr int TILE = THREADS * ITEMS;
    static constexpr int COL = WARPS * ITEMS;     // keys per column (= per lane)
    static constexpr int SCAN_ITEMS = ROWS * 4;   // (row, octet of columns) work items of the scan
    static constexpr int SCAN_ITERS = (SCAN_ITEMS + THREADS - 1) / THREADS;
    // word offsets inside dynamic shared memory
    // keys-only kernels with per-bin destinations lay every bin out at the 16-byte phase of its destination
    // (bulk-copy write-out): up to 6 words of padding per bin
    static constexpr bool BULK = DST && !PAIRS && !WIDE;
    static constexpr int PAD_WORDS = BULK ? (6 * B + 31) / 32 * 32 : 0;
    static constexpr int OFF_BUF = 0;                                   // keys [TILE] (+ values [TILE]); later the reorder buffer
    static constexpr int OFF_TABLE = (PAIRS ? 2 : 1) * TILE + PAD_WORDS;  // [ROWS][32], 128-byte aligned (TILE % 32 == 0)
    static constexpr int OFF_BINSTART = OFF_TABLE + ROWS * 32;          // [B] first tile position of each bin
    static constexpr int OFF_GBASE = OFF_BINSTART + B;                  // [B] or [B] x 64 bit
    static constexpr int OFF_VBASE = OFF_GBASE + (DST ? 2 * B : B);     // [B] x 64 bit (DST pairs)
    static constexpr int OFF_ROWTOT = OFF_VBASE + ((DST && PAIRS) ? 2 * B : 0);  // [ROWS]
    static constexpr int OFF_MISC = OFF_ROWTOT + ROWS;                  // warp totals [32] + pad
    static constexpr int OFF_BAR = (OFF_MISC + 36 + 1) / 2 * 2;          // mbarrier (8 bytes, 8-byte aligned)
    static constexpr int SMEM_WORDS = OFF_BAR + 2;
    static constexpr size_t SMEM_BYTES = (size_t)SMEM_WORDS * 4;
    static_assert(W >= 1 && W <= 8, "digit width");
    // DUAL: the warps form two groups (A = warps [0, WA), B = the rest), each with its own half of the tile
    // and its own 16-bit half of every counter word; both group sizes odd
    static constexpr int WA = !DUAL ? WARPS : ((WARPS / 2) % 2 == 1 ? WARPS / 2 : WARPS / 2 + 1);
    static constexpr int COL_A = WA * ITEMS;
    static constexpr int COL_B = (WARPS - WA) * ITEMS;
    static_assert(DUAL ? (COL_A % 8 == 4 && COL_B % 8 == 4) : (COL % 8 == 4),
                  "column stride must be an odd number of quads (odd warp count per group, ITEMS % 8 == 4)");
    static_assert(WIDE || (TILE + PAD_WORDS) * 4 < 65536, "byte positions must fit the 16-bit counters");  // the DUAL kernel checks its own
    static_assert(WARPS <= 15, "one named barrier per hand-off");
    static_assert((OFF_BAR % 2) == 0, "mbarrier alignment");
};

// shared -> global bulk copy (UBLKCP): bytes % 16 == 0, both addresses 16-byte aligned; completion is
// tracked by the issuing thread's bulk async-group.
__device__ __forceinline__ void bulk_s2g(uint64_t gmem_dst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_src), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_read() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// Exclusive prefix of bin `bin` over the tiles before `tile` (decoupled look-back, LB descriptors in flight).
template <int B, int LB>
__device__ __forceinline__ uint32_t look_back_one_bin(const uint32_t *desc, uint32_t tile, uint32_t bin, uint32_t st_not,
                                                      uint32_t st_inc) {
    uint32_t excl = 0;
    const uint32_t *col = desc + bin;
    int32_t t = (int32_t)tile - 1;
    bool done = (tile == 0);
    while (!done) {
        uint32_t v[LB];
#pragma unroll
        for (int k = 0; k < LB; ++k) v[k] = (t - k >= 0) ? ld_relaxed_gpu(col + (size_t)(t - k) * B) : st_inc;
        bool stop = false;
#pragma unroll
        for (int k = 0; k < LB; ++k) {
            const uint32_t f = v[k] & kDescFlagMask;
            if (!stop && f == st_not) stop = true;  // not published yet: poll again from here
            if (!stop) {
                excl += v[k] & kDescValueMask;
                --t;
                if (f == st_inc) { stop = true; done = true; }
            }
        }
    }
    return excl;
}

template <int W, int WARPS, int ITEMS, int MIN_CTAS, int LB, int GROUP, bool PAIRS, bool DST, bool WIDE = false, bool DUAL = false, bool LBV4 = false>
__global__ void __launch_bounds__(32 * WARPS, MIN_CTAS) colsweep_pass_kernel(const PassArgs a) {
    using TR = ColTraits<W, WARPS, ITEMS, PAIRS, DST, WIDE, DUAL>;
    constexpr int B = TR::B;
    constexpr int ROWS = TR::ROWS;
    constexpr int THREADS = TR::THREADS;
    constexpr int TILE = TR::TILE;
    constexpr bool BULK = TR::BULK;

    extern __shared__ __align__(1024) uint32_t smem[];
    uint32_t *s_buf = smem + TR::OFF_BUF;
    uint32_t *s_binstart = smem + TR::OFF_BINSTART;
    uint32_t *s_gbase = smem + TR::OFF_GBASE;
    uint32_t *s_vbase = smem + TR::OFF_VBASE;
    uint32_t *s_rowtot = smem + TR::OFF_ROWTOT;
        uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem + TR::OFF_BAR);

    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t sa_buf = smem_u32(s_buf);
    const uint32_t sa_table = smem_u32(smem + TR::OFF_TABLE);
    const uint32_t sa_tlane = sa_table + lane * 4u;  // this lane's column of the counter table

    const uint32_t tile = blockIdx.x;
    const uint32_t tile_base = tile * (uint32_t)TILE;
    const uint32_t n_valid = min((uint32_t)TILE, a.n - tile_base);
    const bool full = (n_valid == (uint32_t)TILE);

    // digit d = (key >> shift) & mask.  r = rotr(key, shift - 2) puts d at bits 2..9: bit 2 = d & 1
    // selects the half ((r & 4) * 0xFFFF + 4 = 4 or 4 << 16), (r & m3) = (d >> 1) * 8, times 16 = byte
    // offset of the row.  Both multiply-adds run on the FMA pipe, next to the integer pipe that
    // carries the rotate and the two ANDs.  Counters count in units of 4 (bytes of a key slot).
    const uint32_t rot = (a.shift + 30u) & 31u;
    const uint32_t rot_rank = launder(rot, a.parity >> 8);  // see launder(): keeps ptxas from caching ITEMS digits
    const uint32_t m3 = (a.mask >> 1) << 3;
    const uint32_t rot_w = (a.shift + 30u) & 31u;           // write-out: d * 4 at bits 2..9
    const uint32_t mask4 = a.mask << 2;
    // WIDE: one 32-bit counter per (bin, lane) -- row = digit, constant addend, no half to select.  The keys
    // stay ROTATED (digit * 4 at bits 2..9) in the registers and in the reorder buffer from the load to the
    // write-out, so a table address is two instructions ((r & mask4) << 5) + column) and the counters hold
    // absolute shared-memory addresses: the turn of a warp in the chain is 3 instructions per key.
    const uint32_t mask4_rank = launder(mask4, a.parity >> 8);
    // DUAL (with WIDE): a warp's ranking atomics complete one after the other (16.5 cycles each on B200,
    // tools/gpu_probe3.cu) while those of different warps overlap, so the tile is split into two halves that are
    // ranked at the same time.  Group A (warps [0, WA)) owns tile positions [0, 32 * COL_A) as 32 columns, group
    // B the rest; the word of (bin, lane) holds A's 16-bit counter in its low half and B's in its high half, so
    // the scan is the packed scan of two independent column sets, a bin's keys are A's (column order) followed
    // by B's = index order, and each group runs its own chain of turns.  The addend is a per-warp constant, the
    // half of the returned word a per-warp byte selector.
    static_assert(!DUAL || (WIDE && TR::TILE < 65536), "DUAL: 16-bit positions (in keys)");
    constexpr uint32_t WA = TR::WA;
    const bool group_b = DUAL && warp >= WA;
    const uint32_t gadd = DUAL ? (group_b ? (1u << 16) : 1u) : 4u;  // DUAL counts keys, the other forms bytes
    const uint32_t gsel = group_b ? 0x4432u : 0x4410u;
    auto tile_bytes = [&](uint32_t bin) -> uint32_t {  // bytes of bin `bin` in this tile (4 per key)
        return DUAL ? ((s_rowtot[bin] & 0xFFFFu) + (s_rowtot[bin] >> 16)) << 2
                    : WIDE ? s_rowtot[bin] : ((s_rowtot[bin >> 1] >> ((bin & 1u) * 16u)) & 0xFFFFu);
    };

    COL_STAMP(0);
    // ---- 0. tile -> shared memory --------------------------------------------------------------
#ifdef B200_COL_NOTMA
    const bool use_tma = false &&
#else
    const bool use_tma = full &&
#endif
                         ((reinterpret_cast<uintptr_t>(a.keys_in) & 15u) == 0) &&
                         (!PAIRS || (reinterpret_cast<uintptr_t>(a.vals_in) & 15u) == 0);
    // The SM slot that runs this tile runs tile + (resident CTAs) next: its keys are requested into L2 now, so
    // that the bulk load of that tile is an L2 hit (measured: 0.570 -> 0.548 ms per pass).
    if (WIDE && tid == 32 && a.prefetch != 0u && tile + a.prefetch < a.num_tiles - 1u &&
        ((reinterpret_cast<uintptr_t>(a.keys_in) & 15u) == 0) && (!PAIRS || (reinterpret_cast<uintptr_t>(a.vals_in) & 15u) == 0)) {
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.keys_in + (size_t)(tile + a.prefetch) * TILE), "r"((uint32_t)TILE * 4u) : "memory");
        if (PAIRS)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a.vals_in + (size_t)(tile + a.prefetch) * TILE), "r"((uint32_t)TILE * 4u) : "memory");
    }
    if (use_tma && tid == 0) {
        mbar_init(s_bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(s_bar, (PAIRS ? 2u : 1u) * TILE * 4u);
        bulk_g2s(s_buf, a.keys_in + tile_base, TILE * 4u, s_bar);
        if (PAIRS) bulk_g2s(s_buf + TILE, a.vals_in + tile_base, TILE * 4u, s_bar);
    }
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (int i = tid; i < ROWS * 8; i += THREADS) sm_st4(sa_table + 16u * i, z);
    }
    if (!use_tma) {
        // Ragged last tile (once per launch) or unaligned input: plain loads.  Out-of-range items
        // become all-ones keys: they fall in the highest occupied bin, after every real key of the
        // tile, i.e. at tile positions >= n_valid, and are never written out.
#pragma unroll 4
        for (uint32_t j = tid; j < (uint32_t)TILE; j += THREADS) {
            s_buf[j] = (j < n_valid) ? ld_stream(a.keys_in + tile_base + j) : 0xFFFFFFFFu;
            if (PAIRS) s_buf[TILE + j] = (j < n_valid) ? ld_stream(a.vals_in + tile_base + j) : 0u;
        }
    }
    __syncthreads();  // table zeroed; mbarrier initialised; staged tile complete
    COL_STAMP(1);
    if (use_tma) mbar_wait(s_bar, 0);
    COL_STAMP(2);

    // ---- 1. load: thread (warp, lane) owns keys [lane*COL + warp*ITEMS, +ITEMS) ----------------
    uint32_t key[ITEMS];
    uint32_t val[PAIRS ? ITEMS : 1];
    {
        const uint32_t first = !group_b ? lane * (uint32_t)TR::COL_A + warp * (uint32_t)ITEMS
                                        : 32u * TR::COL_A + lane * (uint32_t)TR::COL_B + (warp - WA) * (uint32_t)ITEMS;
        const uint32_t src = sa_buf + first * 4u;
#pragma unroll
        for (int q = 0; q < ITEMS / 4; ++q) {
            const uint4 v = sm_ld4(src + 16u * q);
            key[4 * q] = v.x; key[4 * q + 1] = v.y; key[4 * q + 2] = v.z; key[4 * q + 3] = v.w;
        }
        if (PAIRS) {
#pragma unroll
            for (int q = 0; q < ITEMS / 4; ++q) {
                const uint4 v = sm_ld4(src + (uint32_t)TILE * 4u + 16u * q);
                val[4 * q] = v.x; val[4 * q + 1] = v.y; val[4 * q + 2] = v.z; val[4 * q + 3] = v.w;
            }
        }
        if (WIDE) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) key[i] = __funnelshift_r(key[i], key[i], rot);
        }
    }

    COL_STAMP(3);
    // ---- 2. count --------------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) {
        if (WIDE) {
            sm_red(((key[i] & mask4) << 5) + sa_tlane, gadd);
        } else {
            const uint32_t r = __funnelshift_r(key[i], key[i], rot);
            sm_red((r & m3) * 16u + sa_tlane, (r & 4u) * 0xFFFFu + 4u);
        }
    }
    COL_STAMP(4);
    __syncthreads();  // counts complete; every thread holds its keys, so the buffer may be overwritten
    COL_STAMP(5);

    // ---- 3. scan ---------------------------------------------------------------------------------
    // Work item it = (row, octet g): 8 columns of one row.  The 4 lanes of a row are neighbours.
    // Row parity swaps the order of the two 16-byte loads so that the 8 lanes of a quarter warp
    // (2 rows x 4 octets) touch 8 different 16-byte bank groups.
    // REREAD (more than 8 warps: 96 registers or fewer per thread): the second half of the scan loads the counters
    // again instead of keeping 8 prefixes per work item in registers across two barriers.
    constexpr bool REREAD = WIDE && WARPS > 8;
    uint32_t ex[REREAD ? 1 : TR::SCAN_ITERS][8];   // exclusive prefix of the item's 8 columns, packed
    uint32_t oct[REREAD ? 1 : TR::SCAN_ITERS];     // + exclusive prefix of the octets before it in the row
    auto load_prefix = [&](uint32_t it, uint32_t (&e)[8]) -> uint32_t {  // returns the total of the 8 columns
        const bool active = it < (uint32_t)TR::SCAN_ITEMS;
        const uint32_t row = it >> 2, g = it & 3u, par = row & 1u;
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        if (active) {
            const uint32_t rb = sa_table + row * 128u + g * 32u;
            const uint4 qa = sm_ld4(rb + par * 16u);
            const uint4 qb = sm_ld4(rb + (par ^ 1u) * 16u);
            lo = par ? qb : qa;
            hi = par ? qa : qb;
        }
        e[0] = 0;
        e[1] = lo.x;
        e[2] = e[1] + lo.y;
        e[3] = e[2] + lo.z;
        e[4] = e[3] + lo.w;
        e[5] = e[4] + hi.x;
        e[6] = e[5] + hi.y;
        e[7] = e[6] + hi.z;
        return e[7] + hi.w;
    };
    auto quad_inclusive = [&](uint32_t tot, uint32_t g) -> uint32_t {  // inclusive prefix over the 4 lanes of a row
        uint32_t incl = tot;
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, 1, 4);
        if (g >= 1u) incl += t;
        t = __shfl_up_sync(0xffffffffu, incl, 2, 4);
        if (g >= 2u) incl += t;
        return incl;
    };
#pragma unroll
    for (int k = 0; k < TR::SCAN_ITERS; ++k) {
        const uint32_t it = tid + k * THREADS;
        const uint32_t row = it >> 2, g = it & 3u;
        uint32_t e[8];
        const uint32_t tot = load_prefix(it, e);
        const uint32_t incl = quad_inclusive(tot, g);
        if (!REREAD) {
#pragma unroll
            for (int j = 0; j < 8; ++j) ex[k][j] = e[j];
            oct[k] = incl - tot;
        }
        if (it < (uint32_t)TR::SCAN_ITEMS && g == 3u) s_rowtot[row] = incl;
    }
    COL_STAMP(6);
    __syncthreads();
    COL_STAMP(7);

    // Bin threads publish the tile's counts; warp 0 turns the row totals into bin starts (byte units).
    const uint32_t st_not = ((2u * a.parity) & 3u) << 30;
    const uint32_t st_agg = ((2u * a.parity + 1u) & 3u) << 30;
    const uint32_t st_inc = ((2u * a.parity + 2u) & 3u) << 30;
    for (uint32_t bin = tid; bin < (uint32_t)B; bin += THREADS) {
        const uint32_t count = tile_bytes(bin) >> 2;
        st_relaxed_gpu(a.desc + (size_t)tile * B + bin, (tile == 0 ? st_inc : st_agg) | count);
        if (BULK) {
            // The destination of this tile's run of every bin must be known BEFORE the keys are ranked: the
            // run is laid out in shared memory at the 16-byte phase of its destination, so that its body is
            // one shared->global bulk copy (to local or peer memory).  Predecessors publish their inclusive
            // prefixes at this same early point, so the walk is short.
            const uint32_t excl = look_back_one_bin<B, LB>(a.desc, tile, bin, st_not, st_inc);
            if (tile != 0) st_relaxed_gpu(a.desc + (size_t)tile * B + bin, st_inc | (excl + count));
            const uint32_t first = a.bin_base[bin] + excl;
            if (a.carry_out != nullptr && tile == a.num_tiles - 1u) a.carry_out[bin] = first + count;
            reinterpret_cast<uint64_t *>(s_gbase)[bin] = a.bin_dst[bin] + 4ull * (uint64_t)first;  // address of the run
        }
    }
    if (BULK) __syncthreads();
    if (warp == 0) {
        uint32_t c[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (WIDE) {
                c[2 * j] = (8u * lane + 2 * j < (uint32_t)B) ? tile_bytes(8u * lane + 2 * j) : 0u;
                c[2 * j + 1] = (8u * lane + 2 * j + 1 < (uint32_t)B) ? tile_bytes(8u * lane + 2 * j + 1) : 0u;
            } else {
                const uint32_t row = 4u * lane + j;
                const uint32_t w = (row < (uint32_t)ROWS) ? s_rowtot[row] : 0u;
                c[2 * j] = w & 0xFFFFu;
                c[2 * j + 1] = w >> 16;
            }
        }
        uint32_t phase[8];  // BULK: byte offset of the run's destination inside its 16-byte line
        uint32_t len[8];    // bytes the bin occupies in the reorder buffer (BULK: padded to whole 16-byte lines)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            phase[j] = 0;
            len[j] = c[j];
            if (BULK) {
                const uint32_t bin = 8u * lane + j;
                if (bin < (uint32_t)B && c[j] != 0u) {
                    phase[j] = (uint32_t)reinterpret_cast<const uint64_t *>(s_gbase)[bin] & 15u;
                    len[j] = (c[j] + phase[j] + 15u) & ~15u;
                }
            }
        }
        uint32_t e[8];
        e[0] = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) e[j] = e[j - 1] + len[j - 1];
        const uint32_t tot = e[7] + len[7];
        uint32_t incl = tot;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += t;
        }
        const uint32_t base = incl - tot;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (8u * lane + j < (uint32_t)B) s_binstart[8u * lane + j] = base + e[j] + phase[j];
    }
    COL_STAMP(8);
    __syncthreads();
    COL_STAMP(9);

#pragma unroll
    for (int k = 0; k < TR::SCAN_ITERS; ++k) {
        const uint32_t it = tid + k * THREADS;
        if (it < (uint32_t)TR::SCAN_ITEMS) {   // whole warps (SCAN_ITEMS and THREADS are multiples of 32)
            const uint32_t row = it >> 2, g = it & 3u, par = row & 1u;
            uint32_t e[8], o;
            if (REREAD) {
                const uint32_t tot = load_prefix(it, e);
                o = quad_inclusive(tot, g) - tot;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) e[j] = ex[k][j];
                o = oct[k];
            }
            uint32_t base;
            if (WIDE) {
                // absolute shared-memory address of the slot (pairs: half of it; the slot is 8 bytes)
                // DUAL: {B's first position : A's first position} of the bin, B's keys after all of A's
                if (DUAL) base = (s_binstart[row] >> 2) * 0x10001u + (s_rowtot[row] << 16) + o;
                else base = s_binstart[row] + o + (PAIRS ? sa_buf / 2u : sa_buf);
            } else {
                const uint2 bs = *reinterpret_cast<const uint2 *>(s_binstart + 2 * row);
                base = (bs.x | (bs.y << 16)) + o;  // byte positions < 2^16: no carry between the halves
            }
            uint4 lo, hi;
            {
            lo.x = base + e[0]; lo.y = base + e[1]; lo.z = base + e[2]; lo.w = base + e[3];
            hi.x = base + e[4]; hi.y = base + e[5]; hi.z = base + e[6]; hi.w = base + e[7];
            }
            const uint32_t rb = sa_table + row * 128u + g * 32u;
            sm_st4(rb + par * 16u, par ? hi : lo);
            sm_st4(rb + (par ^ 1u) * 16u, par ? lo : hi);
        }
    }
    COL_STAMP(10);
    __syncthreads();
    COL_STAMP(11);

    // ---- 4. rank + reorder: the warps take turns on the counter table -----------------------------
    if (warp != 0 && warp != WA) named_bar_sync(warp, 64);
    COL_STAMP(13);
#pragma unroll
    for (int i0 = 0; i0 < ITEMS; i0 += GROUP) {
        uint32_t old[GROUP];
#pragma unroll
        for (int g = 0; g < GROUP; ++g)
            if (i0 + g < ITEMS) {
                if (WIDE) {
                    const uint32_t slot = ((key[i0 + g] & mask4_rank) << 5) + sa_tlane;
                    old[g] = sm_add_ret(slot, gadd);
                } else {
                    const uint32_t r = __funnelshift_r(key[i0 + g], key[i0 + g], rot_rank);
                    old[g] = sm_add_ret((r & m3) * 16u + sa_tlane, (r & 4u) * 0xFFFFu + 4u);
                }
            }
        if (i0 + GROUP >= ITEMS && warp + 1 < (uint32_t)WARPS && warp + 1 != WA) named_bar_arrive(warp + 1u, 64);
        if (i0 + GROUP >= ITEMS) COL_STAMP(14);
#pragma unroll
        for (int g = 0; g < GROUP; ++g)
            if (i0 + g < ITEMS) {
                if (DUAL) {
                    const uint32_t pos = __byte_perm(old[g], 0u, gsel);  // in keys
                    if (PAIRS) sm_st2(sa_buf + 8u * pos, key[i0 + g], val[i0 + g]);
                    else sm_st<0>(sa_buf + 4u * pos, key[i0 + g]);
                } else if (WIDE) {
                    if (PAIRS) sm_st2(2u * old[g], key[i0 + g], val[i0 + g]);
                    else sm_st<0>(old[g], key[i0 + g]);
                } else {
                    const uint32_t r = __funnelshift_r(key[i0 + g], key[i0 + g], rot_rank);
                    const uint32_t pos = (r & 4u) ? (old[g] >> 16) : (old[g] & 0xFFFFu);
                    if (PAIRS) sm_st2(sa_buf + 2u * pos, key[i0 + g], val[i0 + g]);
                    else sm_st<0>(sa_buf + pos, key[i0 + g]);
                }
            }
    }
    COL_STAMP(15);
    // ---- 5. decoupled look-back, done by the warps whose turn in the chain came first (they would only
    // wait for the others): LBT threads, NB bins per thread walked TOGETHER (LB descriptors in flight
    // per bin).  Late in the life of the tile on purpose: the predecessors' aggregates were published
    // long ago, and most of them have finished their own look-back (measured: the same walk placed
    // right after the aggregate publish takes 14 k cycles instead of 2 k).
    constexpr int LB_WARPS = (WARPS - 1 < 4) ? WARPS - 1 : 4;
    constexpr int LBT = 32 * LB_WARPS;
    constexpr int NB = (B + LBT - 1) / LBT;
    // (DUAL: the first warps of both groups, in the order in which their turns end)
    const uint32_t lb_slot = !DUAL ? warp : group_b ? 2u * (warp - WA) + 1u : (warp < (uint32_t)WARPS - WA ? 2u * warp : warp + ((uint32_t)WARPS - WA));
    if constexpr (LBV4 && !BULK && (B % 4 == 0)) {
        // Vector look-back: B / 4 threads, four neighbouring bins each, one 16-byte strong load per tile row (a warp
        // issues one strong load per ~55 cycles, measured: the scalar walk spends most of its round on issuing them).
        // The four bins of a thread consume the rows in order, each at its own position `pos`.
        const uint32_t u = lb_slot * 32u + lane;
        if (u < (uint32_t)B / 4u) {
            uint32_t excl[4] = {0, 0, 0, 0};
            int32_t pos[4];
            bool done[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                pos[c] = (int32_t)tile - 1;
                done[c] = (tile == 0);
            }
            bool all_done = (tile == 0);
#ifdef B200_COL_DEBUG
            long long dbg_rounds = 0;
#endif
            while (!all_done) {
#ifdef B200_COL_DEBUG
                ++dbg_rounds;
#endif
                int32_t t0 = -1;
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (!done[c]) t0 = max(t0, pos[c]);
                uint4 v[LB];
#pragma unroll
                for (int k = 0; k < LB; ++k)
                    v[k] = (t0 - k >= 0) ? ld_relaxed_gpu_v4(a.desc + (size_t)(t0 - k) * B + 4u * u) : make_uint4(st_not, st_not, st_not, st_not);
                all_done = true;
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    const uint32_t w4[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const uint32_t f = w4[c] & kDescFlagMask;
                        if (!done[c] && pos[c] == t0 - k && f != st_not) {
                            excl[c] += w4[c] & kDescValueMask;
                            --pos[c];
                            if (f == st_inc) done[c] = true;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) all_done = all_done && done[c];
            }
#ifdef B200_COL_DEBUG
            if (blockIdx.x == a.num_tiles / 2 && (threadIdx.x & 31u) == 0) {
                g_col_dbg[(threadIdx.x >> 5) * 20 + 18] = dbg_rounds;
                g_col_dbg[(threadIdx.x >> 5) * 20 + 19] = (long long)tile - 1 - pos[0];
            }
#endif
            uint32_t inc4[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint32_t bin = 4u * u + c;
                const uint32_t count = tile_bytes(bin) >> 2;
                const uint32_t bin_start = s_binstart[bin] >> 2;
                inc4[c] = st_inc | (excl[c] + count);
                const uint32_t first = a.bin_base[bin] + excl[c];
                if (a.carry_out != nullptr && tile == a.num_tiles - 1u) a.carry_out[bin] = first + count;
                if (!DST) {
                    s_gbase[bin] = first - bin_start;
                } else {
                    const uint64_t delta = 4ull * (uint64_t)first - 4ull * (uint64_t)bin_start;
                    reinterpret_cast<uint64_t *>(s_gbase)[bin] = a.bin_dst[bin] + delta;
                    if (PAIRS) reinterpret_cast<uint64_t *>(s_vbase)[bin] = a.bin_dst[B + bin] + delta;
                }
            }
            if (tile != 0) st_relaxed_gpu_v4(a.desc + (size_t)tile * B + 4u * u, make_uint4(inc4[0], inc4[1], inc4[2], inc4[3]));
        }
    } else
    if (!BULK && lb_slot < (uint32_t)LB_WARPS) {
        const uint32_t u = lb_slot * 32u + lane;
        uint32_t excl[NB];
        int32_t t[NB];
        bool done[NB];
        bool all_done = true;
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            excl[j] = 0;
            t[j] = (int32_t)tile - 1;
            done[j] = (tile == 0) || (u + j * LBT >= (uint32_t)B);
            all_done = all_done && done[j];
        }
#ifdef B200_COL_DEBUG
        long long dbg_rounds = 0;
#endif
        while (!all_done) {
#ifdef B200_COL_DEBUG
            ++dbg_rounds;
#endif
            uint32_t v[NB][LB];
#pragma unroll
            for (int j = 0; j < NB; ++j)
#pragma unroll
                for (int k = 0; k < LB; ++k)
                    v[j][k] = (!done[j] && t[j] - k >= 0) ? ld_relaxed_gpu(a.desc + (size_t)(t[j] - k) * B + (u + j * LBT)) : st_inc;
            all_done = true;
#pragma unroll
            for (int j = 0; j < NB; ++j) {
                bool stop = done[j];
#pragma unroll
                for (int k = 0; k < LB; ++k) {
                    const uint32_t f = v[j][k] & kDescFlagMask;
                    if (!stop && f == st_not) stop = true;  // not published yet: poll again from here
                    if (!stop) {
                        excl[j] += v[j][k] & kDescValueMask;
                        --t[j];
                        if (f == st_inc) { stop = true; done[j] = true; }
                    }
                }
                all_done = all_done && done[j];
            }
        }
#ifdef B200_COL_DEBUG
        if (blockIdx.x == a.num_tiles / 2 && (threadIdx.x & 31u) == 0) {
            g_col_dbg[(threadIdx.x >> 5) * 20 + 18] = dbg_rounds;
            g_col_dbg[(threadIdx.x >> 5) * 20 + 19] = (long long)tile - 1 - t[0];
        }
#endif
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const uint32_t bin = u + j * LBT;
            if (bin < (uint32_t)B) {
                const uint32_t count = tile_bytes(bin) >> 2;
                const uint32_t bin_start = s_binstart[bin] >> 2;
                if (tile != 0) st_relaxed_gpu(a.desc + (size_t)tile * B + bin, st_inc | (excl[j] + count));
                const uint32_t first = a.bin_base[bin] + excl[j];  // destination index of this tile's first key of the bin
                if (a.carry_out != nullptr && tile == a.num_tiles - 1u) a.carry_out[bin] = first + count;
                if (!DST) {
                    s_gbase[bin] = first - bin_start;  // mod 2^32; + tile position = destination index
                } else {
                    const uint64_t delta = 4ull * (uint64_t)first - 4ull * (uint64_t)bin_start;  // mod 2^64
                    reinterpret_cast<uint64_t *>(s_gbase)[bin] = a.bin_dst[bin] + delta;
                    if (PAIRS) reinterpret_cast<uint64_t *>(s_vbase)[bin] = a.bin_dst[B + bin] + delta;
                }
            }
        }
    }
    COL_STAMP(12);
    if (BULK) fence_proxy_async();  // generic-proxy scatter stores -> visible to the bulk-copy engine
    __syncthreads();
    COL_STAMP(16);

    // ---- 6. write out ---------------------------------------------------------------------------------
    const uint32_t sa_gbase = smem_u32(s_gbase);
    uint32_t *const kout = a.keys_out;
    uint32_t *const vout = a.vals_out;
    constexpr int kWG = 6;  // loads in flight per thread
    auto digit4 = [&](uint32_t k) -> uint32_t { return WIDE ? (k & mask4) : (__funnelshift_r(k, k, rot_w) & mask4); };
    auto original = [&](uint32_t k) -> uint32_t { return WIDE ? __funnelshift_l(k, k, rot_w) : k; };
    if (BULK) {
        // One thread per bin: the run [S, S + bytes) of the reorder buffer goes to its destination as
        // <= 3 head words, one bulk copy of whole 16-byte lines, <= 3 tail words.  The scatter stores were
        // made visible to the async proxy by the fence every thread executed before the barrier above.
        for (uint32_t bin = tid; bin < (uint32_t)B; bin += THREADS) {
            uint32_t bytes = tile_bytes(bin);
            if (!full && bin == a.mask) bytes -= 4u * ((uint32_t)TILE - n_valid);  // the padding keys sit at the end of the top bin
            if (bytes != 0u) {
                const uint32_t src = sa_buf + s_binstart[bin];
                const uint64_t dst = reinterpret_cast<const uint64_t *>(s_gbase)[bin];
                const uint32_t head = min(bytes, (16u - ((uint32_t)dst & 15u)) & 15u);
                const uint32_t body = (bytes - head) & ~15u;
                const uint32_t tail = bytes - head - body;
                if (body != 0u) bulk_s2g(dst + head, src + head, body);
                for (uint32_t o = 0; o < head; o += 4u) *reinterpret_cast<uint32_t *>(dst + o) = sm_ld(src + o);
                for (uint32_t o = head + body; o < head + body + tail; o += 4u)
                    *reinterpret_cast<uint32_t *>(dst + o) = sm_ld(src + o);
                if (body != 0u) bulk_commit_and_wait_read();  // the buffer must outlive the copy engine's reads
            }
        }
    } else if (full) {
#pragma unroll
        for (int k0 = 0; k0 < ITEMS; k0 += kWG) {
            uint32_t kk[kWG], vv[PAIRS ? kWG : 1], gb[kWG];
#pragma unroll
            for (int g = 0; g < kWG; ++g)
                if (k0 + g < ITEMS) {
                    const uint32_t j = tid + (k0 + g) * THREADS;
                    if (PAIRS) {
                        const uint2 kv = sm_ld2(sa_buf + 8u * j);
                        kk[g] = kv.x;
                        vv[g] = kv.y;
                    } else {
                        kk[g] = sm_ld(sa_buf + 4u * j);
                    }
                }
            if (!DST) {
#pragma unroll
                for (int g = 0; g < kWG; ++g)
                    if (k0 + g < ITEMS) gb[g] = sm_ld(sa_gbase + digit4(kk[g]));
#pragma unroll
                for (int g = 0; g < kWG; ++g)
                    if (k0 + g < ITEMS) {
                        const uint32_t j = tid + (k0 + g) * THREADS;
                        kout[gb[g] + j] = original(kk[g]);
                        if (PAIRS) vout[gb[g] + j] = vv[g];
                    }
            } else {
#pragma unroll
                for (int g = 0; g < kWG; ++g)
                    if (k0 + g < ITEMS) {
                        const uint32_t j = tid + (k0 + g) * THREADS;
                        const uint32_t d = digit4(kk[g]) >> 2;
                        const uint64_t off = 4ull * j;
                        *reinterpret_cast<uint32_t *>(reinterpret_cast<const uint64_t *>(s_gbase)[d] + off) = original(kk[g]);
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
                    const uint2 kv = sm_ld2(sa_buf + 8u * j);
                    kk = kv.x;
                    vv = kv.y;
                } else {
                    kk = sm_ld(sa_buf + 4u * j);
                }
                const uint32_t d4 = digit4(kk);
                kk = original(kk);
                if (!DST) {
                    const uint32_t g = sm_ld(sa_gbase + d4) + j;
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
    COL_STAMP(17);
}

}  // namespace b200sort
