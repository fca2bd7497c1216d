// launch.h -- interface between the host driver (b200sort.cu) and the per-width kernel
// translation units (kernels_w.cu compiled once per digit width with -DB200_W=1..8).
#pragma once
#include "common.cuh"

namespace b200sort {

// Digit-pass kernel variants: tile geometry x rank mode (see onesweep.cuh).  Variant 0 is the
// default and exists for every width; the others are tuning alternatives instantiated for the
// 8-bit digit only.  mode: 0 = RANK_TABLE, 1 = RANK_ATOMIC, 2 = RANK_MATCH.
struct PassVariant {
    int threads;
    int items_keys;
    int items_pairs;
    int min_ctas;
    int mode;        // 0 table rank, 1 atomic rank, 2 match.any (onesweep.cuh); 3 = column sweep (colsweep.cuh), 4 = with 32-bit counters
    int table_bits;  // mode 0: bits of the peer table; mode 3: ranking atomics in flight per thread
    int lb_batch;  // look-back descriptors in flight per bin thread
    int persist;   // 1: persistent CTAs that prefetch their next tile; 2: one tile per CTA, loaded by TMA; mode 4: concurrent ranking chains (2 = two warp groups)
};
constexpr int kNumVariants = 126;
constexpr PassVariant kVariants[kNumVariants] = {
    {256, 30, 20, 4, 0, 5, 8, 0},   //  0 default: table(5 bits) + 3 ballots
    {256, 30, 20, 4, 1, 0, 8, 0},   //  1 atomic rank (selected only after the self test passes)
    {256, 36, 24, 4, 0, 5, 8, 0},   //  2
    {256, 36, 20, 4, 1, 0, 8, 0},   //  3
    {384, 20, 14, 3, 0, 5, 8, 0},   //  4
    {384, 20, 14, 3, 1, 0, 8, 0},   //  5
    {256, 30, 20, 4, 0, 8, 8, 0},   //  6 full-digit atomicOr table, no ballots
    {256, 30, 20, 4, 0, 6, 8, 0},   //  7 table(6 bits) + 2 ballots
    {384, 20, 14, 2, 2, 0, 8, 0},   //  8 match.any (for the record)
    {256, 40, 24, 3, 1, 0, 8, 0},   //  9
    {256, 44, 22, 3, 1, 0, 8, 0},   // 10
    {512, 36, 24, 2, 1, 0, 8, 0},   // 11
    {512, 24, 16, 3, 1, 0, 8, 0},   // 12
    {384, 32, 20, 3, 1, 0, 8, 0},   // 13
    {384, 24, 16, 3, 1, 0, 8, 0},   // 14
    {512, 30, 20, 2, 1, 0, 8, 0},   // 15
    {256, 30, 20, 4, 0, 0, 8, 0},   // 16 ballots only (narrow digits); instantiated for every width
    {256, 44, 22, 3, 1, 0, 4, 0},   // 17 = 10 with a 4-deep look-back
    {256, 44, 22, 3, 1, 0, 16, 0},  // 18 = 10 with a 16-deep look-back
    {384, 44, 22, 2, 1, 0, 8, 0},   // 19
    {512, 22, 12, 3, 1, 0, 8, 0},   // 20
    {384, 30, 16, 3, 1, 0, 8, 0},   // 21
    {256, 44, 22, 4, 1, 0, 8, 0},   // 22 = 10 at four CTAs per SM
    {256, 40, 20, 4, 1, 0, 8, 0},   // 23
    {256, 52, 26, 3, 1, 0, 8, 0},   // 24
    {256, 60, 30, 3, 1, 0, 8, 0},   // 25
    {256, 64, 32, 2, 1, 0, 8, 0},   // 26
    {256, 48, 24, 3, 1, 0, 8, 0},   // 27
    {256, 16, 12, 6, 0, 0, 8, 0},   // 28 ballots only, small tile / high occupancy (fused exchange experiments)
    {256, 44, 22, 3, 1, 0, 8, 1},   // 29 = 10, persistent
    {256, 44, 22, 2, 1, 0, 8, 1},   // 30 persistent, two CTAs per SM
    {256, 36, 20, 4, 1, 0, 8, 1},   // 31 persistent, four CTAs per SM
    {256, 44, 22, 3, 1, 0, 8, 2},   // 32 = 10 with TMA bulk loads of the tile
    {256, 36, 20, 4, 1, 0, 8, 2},   // 33 TMA, four CTAs per SM
    {256, 64, 44, 2, 1, 0, 8, 0},   // 34 two CTAs per SM: 64 keys or 44 pairs per thread
    {256, 64, 36, 2, 1, 0, 8, 0},   // 35
    // column sweep (colsweep.cuh): threads = 32 x (odd warp count), keys per thread % 8 == 4
    {288, 36, 20, 3, 3, 12, 4, 0},  // 36 nine warps x 36 keys, three CTAs per SM, 12 ranking atomics and 4 look-back descriptors in flight
    {288, 36, 20, 3, 3, 12, 8, 0},  // 37 = 36, 8-deep look-back
    {288, 36, 20, 3, 3, 12, 2, 0},  // 38 = 36, 2-deep look-back
    {288, 36, 20, 2, 3, 36, 4, 0},  // 39 two CTAs per SM, a whole turn in flight
    {288, 52, 28, 2, 3, 26, 4, 0},  // 40
    {288, 28, 20, 3, 3, 28, 4, 0},  // 41 28 keys, a whole turn in flight
    {288, 36, 20, 3, 3, 36, 4, 0},  // 42 = 36, a whole turn in flight
    {288, 28, 12, 4, 3, 14, 4, 0},  // 43 four CTAs per SM
    {160, 52, 28, 4, 3, 13, 4, 0},  // 44 five warps (shorter chain), four CTAs per SM
    {160, 36, 20, 5, 3, 12, 4, 0},  // 45 five warps, five CTAs per SM
    {224, 44, 20, 3, 3, 11, 4, 0},  // 46 seven warps, three CTAs per SM
    {224, 36, 20, 4, 3, 12, 4, 0},  // 47 seven warps, four CTAs per SM
    {160, 68, 36, 3, 3, 17, 4, 0},  // 48 five warps x 68 keys, three CTAs per SM
    {96, 84, 44, 5, 3, 14, 4, 0},   // 49 three warps x 84 keys, five CTAs per SM
    // column sweep with 32-bit counters (mode 4): 32 KB table, larger tiles, two CTAs per SM
    {288, 52, 28, 2, 4, 26, 4, 0},  // 50
    {288, 52, 28, 2, 4, 52, 4, 0},  // 51 = 50, a whole turn in flight
    {416, 36, 20, 2, 4, 36, 4, 0},  // 52 thirteen warps x 36 keys
    {288, 60, 28, 2, 4, 20, 4, 0},  // 53
    {352, 44, 20, 2, 4, 22, 4, 0},  // 54 eleven warps x 44 keys
    {288, 36, 20, 2, 4, 36, 4, 0},  // 55
    // mode 4 with persist = 2: two 16-bit counters per word, the two halves of the warps rank concurrently
    {320, 44, 20, 2, 4, 22, 4, 2},  // 56 ten warps (5 + 5) x 44 keys
    {320, 44, 20, 2, 4, 11, 4, 2},  // 57
    {320, 44, 20, 2, 4, 44, 4, 2},  // 58 a whole turn in flight
    {256, 60, 28, 2, 4, 20, 4, 2},  // 59 eight warps (5 + 3) x 60 keys, 128 registers
    {192, 76, 36, 2, 4, 19, 4, 2},  // 60 six warps (3 + 3) x 76 keys
    {448, 28, 12, 2, 4, 14, 4, 2},  // 61 fourteen warps (7 + 7) x 28 keys
    {320, 36, 20, 2, 4, 18, 4, 2},  // 62 ten warps x 36 keys
    {384, 36, 20, 2, 4, 18, 4, 2},  // 63 twelve warps (7 + 5) x 36 keys
    {256, 60, 28, 2, 4, 20, 8, 2},  // 64 = 59, 8-deep look-back
    {256, 60, 28, 2, 4, 20, 16, 2}, // 65 = 59, 16-deep look-back
    {256, 60, 28, 2, 4, 30, 8, 2},  // 66
    {320, 44, 20, 2, 4, 22, 8, 2},  // 67 = 56, 8-deep look-back
    {320, 44, 20, 2, 4, 22, 16, 2}, // 68
    {256, 52, 28, 2, 4, 26, 8, 2},  // 69
    {192, 76, 36, 2, 4, 19, 8, 2},  // 70
    {256, 60, 28, 2, 4, 20, 12, 2}, // 71
    {256, 36, 20, 3, 4, 18, 4, 2},  // 72 three CTAs per SM
    {256, 36, 20, 3, 4, 18, 8, 2},  // 73
    {192, 44, 20, 3, 4, 22, 4, 2},  // 74 six warps (3 + 3) x 44 keys, three CTAs per SM
    {320, 28, 12, 3, 4, 14, 4, 2},  // 75 ten warps x 28 keys, three CTAs per SM
    // 76-83: were the scan-agent experiment (CTA 0 turns the tiles' aggregates into exclusive prefixes, the tiles poll
    // their own row: bit-exact, 0.72 ms per pass -- one warp issues a strong load every ~55 cycles, the agent cannot
    // keep up with a row every 55 cycles; profiles/r02_experiment_scan_agent.patch).  Now plain 4-deep look-back.
    {256, 60, 28, 2, 4, 20, 4, 2},  // 76
    {320, 44, 20, 2, 4, 22, 4, 2},  // 77
    {256, 60, 28, 2, 4, 30, 4, 2},  // 78
    {192, 76, 36, 2, 4, 19, 4, 2},  // 79
    {288, 52, 28, 2, 4, 26, 4, 0},  // 80
    {352, 44, 20, 2, 4, 22, 4, 0},  // 81
    {256, 36, 20, 3, 4, 18, 4, 2},  // 82
    {256, 52, 28, 2, 4, 26, 4, 2},  // 83
    // mode 4 with lb_batch = 32 + d: look-back with 16-byte strong loads (four bins per thread), d rows in flight
    {256, 60, 28, 2, 4, 20, 36, 2}, // 84 = 59, d = 4
    {256, 60, 28, 2, 4, 20, 40, 2}, // 85 d = 8
    {256, 60, 28, 2, 4, 20, 44, 2}, // 86 d = 12
    {256, 60, 28, 2, 4, 20, 48, 2}, // 87 d = 16
    {320, 44, 20, 2, 4, 22, 40, 2}, // 88 = 56, d = 8
    {320, 44, 20, 2, 4, 22, 48, 2}, // 89 d = 16
    {352, 44, 20, 2, 4, 22, 40, 0}, // 90 = 54 (one chain), d = 8
    {192, 76, 36, 2, 4, 19, 40, 2}, // 91 = 60, d = 8
    {256, 60, 28, 2, 4, 20, 34, 2}, // 92 = 84, d = 2
    {256, 60, 28, 2, 4, 20, 35, 2}, // 93 d = 3
    {256, 60, 28, 2, 4, 20, 38, 2}, // 94 d = 6
    {256, 76, 36, 2, 4, 13, 36, 2}, // 95 THE DEFAULT (kDualVariant): 76 keys / 36 pairs per thread, 13 atomics in flight
    {256, 60, 28, 2, 4, 15, 36, 2}, // 96
    {256, 60, 28, 2, 4, 30, 36, 2}, // 97
    {256, 60, 28, 2, 4, 60, 36, 2}, // 98 a whole turn in flight
    {256, 52, 28, 2, 4, 26, 36, 2}, // 99
    {256, 44, 20, 2, 4, 22, 36, 2}, // 100
    {256, 44, 20, 3, 4, 22, 36, 2}, // 101 44 keys, three CTAs per SM
    {256, 36, 20, 3, 4, 18, 36, 2}, // 102
    {192, 76, 36, 2, 4, 19, 36, 2}, // 103 six warps (3 + 3)
    {320, 44, 20, 2, 4, 22, 36, 2}, // 104 ten warps (5 + 5)
    {384, 36, 20, 2, 4, 18, 36, 2}, // 105 twelve warps (7 + 5)
    {352, 44, 20, 2, 4, 22, 36, 0}, // 106 one chain, eleven warps
    {256, 60, 36, 2, 4, 10, 36, 2}, // 107 the default before positions counted keys (60 keys per thread: 16-bit byte positions)
    {256, 60, 36, 2, 4, 10, 36, 2}, // 108 = 95 with 36 pairs per thread
    {256, 60, 36, 2, 4, 10, 34, 2}, // 109
    {256, 68, 36, 2, 4, 10, 36, 2}, // 110 68 keys per thread (positions count keys, not bytes: tiles of up to 65,535 keys)
    {256, 76, 36, 2, 4, 10, 36, 2}, // 111
    {256, 68, 44, 2, 4, 17, 36, 2}, // 112
    {256, 84, 36, 1, 4, 12, 36, 2}, // 113 one CTA per SM
    {256, 76, 36, 2, 4, 19, 36, 2}, // 114
    {256, 76, 36, 2, 4, 13, 36, 2}, // 115
    {192, 100, 44, 2, 4, 10, 36, 2}, // 116 six warps (3 + 3) x 100 keys
    {192, 92, 44, 2, 4, 23, 36, 2}, // 117
    {256, 76, 36, 2, 4, 10, 40, 2}, // 118 8 rows of look-back in flight
    {256, 76, 36, 2, 4, 10, 35, 2}, // 119 3 rows
    {320, 60, 28, 2, 4, 12, 36, 2}, // 120 ten warps (5 + 5: balanced chains) x 60 keys; the scan re-reads the counters
    {320, 60, 28, 2, 4, 20, 36, 2}, // 121
    {320, 52, 28, 2, 4, 13, 36, 2}, // 122
    {384, 44, 20, 2, 4, 11, 36, 2}, // 123 twelve warps (7 + 5) x 44 keys
    {448, 44, 20, 2, 4, 11, 36, 2}, // 124 fourteen warps (7 + 7) x 44 keys
    {320, 60, 28, 2, 4, 10, 36, 2}, // 125
};
inline int tile_keys(int variant, bool pairs) {
    const PassVariant &g = kVariants[variant];
    return g.threads * (pairs ? g.items_pairs : g.items_keys);
}
inline int variant_mode(int variant) { return kVariants[variant].mode; }
constexpr int kColVariant = 36;  // column-sweep default geometry
// Smallest tile over all variants: bounds the descriptor array when sizing temp storage.
constexpr int kMinTileKeys = 2048;

// true if (width, variant) is instantiated.  The product library carries only the kernels its automatic choice
// can select -- ballot rank (kBallotVariant) for digits of <= 3 bits, the column sweep with two ranking chains
// (kDualVariant) for wider ones, and the first column sweep (kColVariant: keys with per-bin destinations and bulk-copy
// write-out, every width).  -DB200_TUNING (B200_TUNING=1 python -m cuda.radixsort_b200.build --force) adds every
// geometry of the table above -- the atomic-rank kernels of round 1 among them -- for tools/sweep.py and the variant
// tests.
constexpr int kBallotVariant = 16;
constexpr int kBallotSmallVariant = 28;
#ifdef B200_TUNING
constexpr bool kTuningBuild = true;
#else
constexpr bool kTuningBuild = false;
#endif
constexpr int kDualVariant = 95;  // column sweep, 32-bit counter words, two ranking chains: keys, digits of >= 4 bits
constexpr bool variant_compiled(int width, int variant) {
    if (variant == kColVariant) return true;
    if (variant == kDualVariant && width >= 4) return true;
    if (kTuningBuild)
        return variant == 0 || variant == 1 || variant == kBallotVariant || variant == kBallotSmallVariant ||
               (width == 8 && variant > 0 && variant < kNumVariants);
    return width <= 3 && variant == kBallotVariant;
}
inline bool variant_available(int width, int variant) { return variant_compiled(width, variant); }
// what a request for an unavailable variant falls back to
inline int fallback_variant(int width) { return width <= 3 ? kBallotVariant : kDualVariant; }
// variants that are instantiated with per-bin destinations (b200sort_digit_pass with d_bin_dst)
constexpr bool variant_has_dst(int width, int variant) {
    return variant_compiled(width, variant) &&
           (variant <= 1 || variant == kBallotVariant || variant == kBallotSmallVariant || variant == kColVariant ||
            variant == kDualVariant);
}

#define B200_DECLARE_W(w)                                                                        \
    cudaError_t launch_hist_w##w(bool uniform, const HistArgs &a, int grid, cudaStream_t s);     \
    cudaError_t launch_pass_w##w(int variant, bool pairs, bool dst, const PassArgs &a,           \
                                 cudaStream_t s);
B200_DECLARE_W(1)
B200_DECLARE_W(2)
B200_DECLARE_W(3)
B200_DECLARE_W(4)
B200_DECLARE_W(5)
B200_DECLARE_W(6)
B200_DECLARE_W(7)
B200_DECLARE_W(8)
#undef B200_DECLARE_W

// Runs the RANK_ATOMIC self test on the current device; *mismatches = number of violations.
cudaError_t run_atomic_order_selftest(uint32_t *d_counter, int blocks, int rounds, cudaStream_t s);
// Dynamic shared memory of the histogram kernel for `passes` passes of 2^width bins.
inline size_t hist_smem_bytes(int passes, int width) { return (size_t)passes * (1u << width) * 32 * 4; }

}  // namespace b200sort
