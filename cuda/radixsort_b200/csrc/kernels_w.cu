// kernels_w.cu -- instantiates the histogram and digit-pass kernels for ONE digit width.
// Compiled eight times (-DB200_W=1 .. 8) so the widths build in parallel.
#include <algorithm>
#include <atomic>

#include "hist.cuh"
#include "launch.h"
#include "onesweep.cuh"
#include "colsweep.cuh"

#ifndef B200_W
#error "compile with -DB200_W=<1..8>"
#endif

namespace b200sort {

namespace {
constexpr int W = B200_W;
constexpr int P_UNIFORM = (32 + W - 1) / W;

template <int V, bool PAIRS, bool DST>
cudaError_t launch_col_variant(const PassArgs &a, cudaStream_t s) {
    constexpr PassVariant g = kVariants[V];
    constexpr int ITEMS = PAIRS ? g.items_pairs : g.items_keys;
    constexpr int WARPS = g.threads / 32;
    constexpr int GROUP = g.table_bits < ITEMS ? g.table_bits : ITEMS;
    constexpr bool WIDE = (g.mode == 4);
    constexpr bool DUAL = WIDE && g.persist == 2;  // mode 4: persist = number of concurrent ranking chains (1 or 2)
    using TR = ColTraits<W, WARPS, ITEMS, PAIRS, DST, WIDE, DUAL>;
    constexpr bool LBV4 = WIDE && g.lb_batch >= 32;  // mode 4, lb_batch 32 + d: 16-byte look-back loads, d rows in flight
    constexpr int LBD = LBV4 ? g.lb_batch - 32 : g.lb_batch;
    auto kernel = colsweep_pass_kernel<W, WARPS, ITEMS, g.min_ctas, LBD, GROUP, PAIRS, DST, WIDE, DUAL, LBV4>;
    static std::atomic<uint64_t> configured{0};  // one bit per device: the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TR::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    kernel<<<a.num_tiles, g.threads, TR::SMEM_BYTES, s>>>(a);
    return cudaGetLastError();
}

template <int V, bool PAIRS, bool DST>
cudaError_t launch_one_variant(const PassArgs &a, cudaStream_t s) {
    constexpr PassVariant g = kVariants[V];
    constexpr int ITEMS = PAIRS ? g.items_pairs : g.items_keys;
    constexpr int TB = g.table_bits;
    using TR = PassTraits<W, g.threads, ITEMS, g.mode, TB, PAIRS, DST>;
    auto kernel = onesweep_pass_kernel<W, g.threads, ITEMS, g.min_ctas, g.mode, TB, g.lb_batch, (g.persist == 1), (g.persist == 2), PAIRS, DST>;
    static std::atomic<uint64_t> configured{0};  // one bit per device: the attribute is per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(configured.load(std::memory_order_acquire) >> (dev & 63) & 1u)) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TR::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured.fetch_or(1ull << (dev & 63), std::memory_order_release);
    }
    unsigned grid = a.num_tiles;
    if (g.persist == 1) {
        // every CTA of a persistent launch must be resident: the look-back spins on tiles that
        // belong to other CTAs of the same launch
        static std::atomic<int> resident[64] = {};
        if (resident[dev & 63] == 0) {
            int per_sm = 0, sms = 0;
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, g.threads, TR::SMEM_BYTES);
            if (e != cudaSuccess) return e;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            resident[dev & 63] = std::max(1, std::min(per_sm, g.min_ctas) * sms);
        }
        grid = std::min<unsigned>(grid, (unsigned)resident[dev & 63]);
    }
    kernel<<<grid, g.threads, TR::SMEM_BYTES, s>>>(a);
    return cudaGetLastError();
}

template <int V, bool PAIRS, bool DST>
cudaError_t launch_variant(const PassArgs &a, cudaStream_t s) {
    if constexpr (kVariants[V].mode >= 3) return launch_col_variant<V, PAIRS, DST>(a, s);
    else return launch_one_variant<V, PAIRS, DST>(a, s);
}

template <int V>
cudaError_t launch_modes(bool pairs, bool dst, const PassArgs &a, cudaStream_t s) {
    if constexpr (!variant_compiled(W, V)) {
        return cudaErrorInvalidValue;
    } else {
        if (!pairs && !dst) return launch_variant<V, false, false>(a, s);
        if (pairs && !dst) return launch_variant<V, true, false>(a, s);
        if constexpr (variant_has_dst(W, V)) {
            if (!pairs && dst) return launch_variant<V, false, true>(a, s);
            return launch_variant<V, true, true>(a, s);
        }
        return cudaErrorInvalidValue;
    }
}

template <int P_CT>
cudaError_t launch_hist_impl(const HistArgs &a, int passes, int grid, cudaStream_t s) {
    auto kernel = hist_kernel<W, P_CT>;
    const size_t smem = hist_smem_bytes(passes, W);
    static std::atomic<size_t> configured[64] = {};  // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = smem;
    }
    kernel<<<grid, kHistThreads, smem, s>>>(a);
    return cudaGetLastError();
}
}  // namespace

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)

cudaError_t B200_CAT(launch_hist_w, B200_W)(bool uniform, const HistArgs &a, int grid, cudaStream_t s) {
    if (uniform) return launch_hist_impl<P_UNIFORM>(a, P_UNIFORM, grid, s);
    if (a.passes.count == 1) return launch_hist_impl<-1>(a, 1, grid, s);
    return launch_hist_impl<0>(a, a.passes.count, grid, s);
}

cudaError_t B200_CAT(launch_pass_w, B200_W)(int variant, bool pairs, bool dst, const PassArgs &a,
                                            cudaStream_t s) {
    if (!variant_compiled(W, variant)) variant = fallback_variant(W);
    if (dst && !variant_has_dst(W, variant)) variant = fallback_variant(W);
    switch (variant) {
    case 0: return launch_modes<0>(pairs, dst, a, s);
    case 1: return launch_modes<1>(pairs, dst, a, s);
    case kBallotVariant: return launch_modes<kBallotVariant>(pairs, dst, a, s);
    case kBallotSmallVariant: return launch_modes<kBallotSmallVariant>(pairs, dst, a, s);
    case kColVariant: return launch_modes<kColVariant>(pairs, dst, a, s);
    case kDualVariant: return launch_modes<kDualVariant>(pairs, dst, a, s);
#if B200_W == 8
    case 2: return launch_modes<2>(pairs, dst, a, s);
    case 3: return launch_modes<3>(pairs, dst, a, s);
    case 4: return launch_modes<4>(pairs, dst, a, s);
    case 5: return launch_modes<5>(pairs, dst, a, s);
    case 6: return launch_modes<6>(pairs, dst, a, s);
    case 7: return launch_modes<7>(pairs, dst, a, s);
    case 8: return launch_modes<8>(pairs, dst, a, s);
    case 9: return launch_modes<9>(pairs, dst, a, s);
    case 10: return launch_modes<10>(pairs, dst, a, s);
    case 11: return launch_modes<11>(pairs, dst, a, s);
    case 12: return launch_modes<12>(pairs, dst, a, s);
    case 13: return launch_modes<13>(pairs, dst, a, s);
    case 14: return launch_modes<14>(pairs, dst, a, s);
    case 15: return launch_modes<15>(pairs, dst, a, s);
    case 17: return launch_modes<17>(pairs, dst, a, s);
    case 18: return launch_modes<18>(pairs, dst, a, s);
    case 19: return launch_modes<19>(pairs, dst, a, s);
    case 20: return launch_modes<20>(pairs, dst, a, s);
    case 21: return launch_modes<21>(pairs, dst, a, s);
    case 22: return launch_modes<22>(pairs, dst, a, s);
    case 23: return launch_modes<23>(pairs, dst, a, s);
    case 24: return launch_modes<24>(pairs, dst, a, s);
    case 25: return launch_modes<25>(pairs, dst, a, s);
    case 26: return launch_modes<26>(pairs, dst, a, s);
    case 27: return launch_modes<27>(pairs, dst, a, s);
    case 29: return launch_modes<29>(pairs, dst, a, s);
    case 30: return launch_modes<30>(pairs, dst, a, s);
    case 31: return launch_modes<31>(pairs, dst, a, s);
    case 32: return launch_modes<32>(pairs, dst, a, s);
    case 33: return launch_modes<33>(pairs, dst, a, s);
    case 34: return launch_modes<34>(pairs, dst, a, s);
    case 35: return launch_modes<35>(pairs, dst, a, s);
    case 37: return launch_modes<37>(pairs, dst, a, s);
    case 38: return launch_modes<38>(pairs, dst, a, s);
    case 39: return launch_modes<39>(pairs, dst, a, s);
    case 40: return launch_modes<40>(pairs, dst, a, s);
    case 41: return launch_modes<41>(pairs, dst, a, s);
    case 42: return launch_modes<42>(pairs, dst, a, s);
    case 43: return launch_modes<43>(pairs, dst, a, s);
    case 44: return launch_modes<44>(pairs, dst, a, s);
    case 45: return launch_modes<45>(pairs, dst, a, s);
    case 46: return launch_modes<46>(pairs, dst, a, s);
    case 47: return launch_modes<47>(pairs, dst, a, s);
    case 48: return launch_modes<48>(pairs, dst, a, s);
    case 49: return launch_modes<49>(pairs, dst, a, s);
    case 50: return launch_modes<50>(pairs, dst, a, s);
    case 51: return launch_modes<51>(pairs, dst, a, s);
    case 52: return launch_modes<52>(pairs, dst, a, s);
    case 53: return launch_modes<53>(pairs, dst, a, s);
    case 54: return launch_modes<54>(pairs, dst, a, s);
    case 55: return launch_modes<55>(pairs, dst, a, s);
    case 56: return launch_modes<56>(pairs, dst, a, s);
    case 57: return launch_modes<57>(pairs, dst, a, s);
    case 58: return launch_modes<58>(pairs, dst, a, s);
    case 59: return launch_modes<59>(pairs, dst, a, s);
    case 60: return launch_modes<60>(pairs, dst, a, s);
    case 61: return launch_modes<61>(pairs, dst, a, s);
    case 62: return launch_modes<62>(pairs, dst, a, s);
    case 63: return launch_modes<63>(pairs, dst, a, s);
    case 64: return launch_modes<64>(pairs, dst, a, s);
    case 65: return launch_modes<65>(pairs, dst, a, s);
    case 66: return launch_modes<66>(pairs, dst, a, s);
    case 67: return launch_modes<67>(pairs, dst, a, s);
    case 68: return launch_modes<68>(pairs, dst, a, s);
    case 69: return launch_modes<69>(pairs, dst, a, s);
    case 70: return launch_modes<70>(pairs, dst, a, s);
    case 71: return launch_modes<71>(pairs, dst, a, s);
    case 72: return launch_modes<72>(pairs, dst, a, s);
    case 73: return launch_modes<73>(pairs, dst, a, s);
    case 74: return launch_modes<74>(pairs, dst, a, s);
    case 75: return launch_modes<75>(pairs, dst, a, s);
    case 76: return launch_modes<76>(pairs, dst, a, s);
    case 77: return launch_modes<77>(pairs, dst, a, s);
    case 78: return launch_modes<78>(pairs, dst, a, s);
    case 79: return launch_modes<79>(pairs, dst, a, s);
    case 80: return launch_modes<80>(pairs, dst, a, s);
    case 81: return launch_modes<81>(pairs, dst, a, s);
    case 82: return launch_modes<82>(pairs, dst, a, s);
    case 83: return launch_modes<83>(pairs, dst, a, s);
    case 84: return launch_modes<84>(pairs, dst, a, s);
    case 85: return launch_modes<85>(pairs, dst, a, s);
    case 86: return launch_modes<86>(pairs, dst, a, s);
    case 87: return launch_modes<87>(pairs, dst, a, s);
    case 88: return launch_modes<88>(pairs, dst, a, s);
    case 89: return launch_modes<89>(pairs, dst, a, s);
    case 90: return launch_modes<90>(pairs, dst, a, s);
    case 91: return launch_modes<91>(pairs, dst, a, s);
    case 92: return launch_modes<92>(pairs, dst, a, s);
    case 93: return launch_modes<93>(pairs, dst, a, s);
    case 94: return launch_modes<94>(pairs, dst, a, s);
    case 96: return launch_modes<96>(pairs, dst, a, s);
    case 97: return launch_modes<97>(pairs, dst, a, s);
    case 98: return launch_modes<98>(pairs, dst, a, s);
    case 99: return launch_modes<99>(pairs, dst, a, s);
    case 100: return launch_modes<100>(pairs, dst, a, s);
    case 101: return launch_modes<101>(pairs, dst, a, s);
    case 102: return launch_modes<102>(pairs, dst, a, s);
    case 103: return launch_modes<103>(pairs, dst, a, s);
    case 104: return launch_modes<104>(pairs, dst, a, s);
    case 105: return launch_modes<105>(pairs, dst, a, s);
    case 106: return launch_modes<106>(pairs, dst, a, s);
    case 107: return launch_modes<107>(pairs, dst, a, s);
    case 108: return launch_modes<108>(pairs, dst, a, s);
    case 109: return launch_modes<109>(pairs, dst, a, s);
    case 110: return launch_modes<110>(pairs, dst, a, s);
    case 111: return launch_modes<111>(pairs, dst, a, s);
    case 112: return launch_modes<112>(pairs, dst, a, s);
    case 113: return launch_modes<113>(pairs, dst, a, s);
    case 114: return launch_modes<114>(pairs, dst, a, s);
    case 115: return launch_modes<115>(pairs, dst, a, s);
    case 116: return launch_modes<116>(pairs, dst, a, s);
    case 117: return launch_modes<117>(pairs, dst, a, s);
    case 118: return launch_modes<118>(pairs, dst, a, s);
    case 119: return launch_modes<119>(pairs, dst, a, s);
    case 120: return launch_modes<120>(pairs, dst, a, s);
    case 121: return launch_modes<121>(pairs, dst, a, s);
    case 122: return launch_modes<122>(pairs, dst, a, s);
    case 123: return launch_modes<123>(pairs, dst, a, s);
    case 124: return launch_modes<124>(pairs, dst, a, s);
    case 125: return launch_modes<125>(pairs, dst, a, s);
#endif
    default: return cudaErrorInvalidValue;
    }
}

#if B200_W == 8 && defined(B200_COL_DEBUG)
extern "C" int b200sort_debug_read(long long *out) {
    return (int)cudaMemcpyFromSymbol(out, g_col_dbg, sizeof(long long) * 16 * 20);
}
#endif

#if B200_W == 8
cudaError_t run_atomic_order_selftest(uint32_t *d_counter, int blocks, int rounds, cudaStream_t s) {
    atomic_order_selftest<0><<<blocks, 256, 0, s>>>(d_counter, rounds, 0x5EED1234u);
    return cudaGetLastError();
}
#endif

}  // namespace b200sort
