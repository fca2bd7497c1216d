// kernels_w.cu -- instantiates the histogram and digit-pass kernels for ONE digit width.
// Compiled eight times (-DB200_W=1 .. 8) so the widths build in parallel.
#include "hist.cuh"
#include "launch.h"
#include "onesweep.cuh"

#ifndef B200_W
#error "compile with -DB200_W=<1..8>"
#endif

namespace b200sort {

namespace {
constexpr int W = B200_W;
constexpr int P_UNIFORM = (32 + W - 1) / W;

template <int V, bool PAIRS, bool DST>
cudaError_t launch_variant(const PassArgs &a, cudaStream_t s) {
    constexpr PassGeometry g = kGeometry[V];
    constexpr int ITEMS = PAIRS ? g.items_pairs : g.items_keys;
    using TR = PassTraits<W, g.threads, ITEMS, PAIRS, DST>;
    auto kernel = onesweep_pass_kernel<W, g.threads, ITEMS, g.min_ctas, PAIRS, DST>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TR::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    kernel<<<a.num_tiles, g.threads, TR::SMEM_BYTES, s>>>(a);
    return cudaGetLastError();
}

template <int V>
cudaError_t launch_modes(bool pairs, bool dst, const PassArgs &a, cudaStream_t s) {
    if (!pairs && !dst) return launch_variant<V, false, false>(a, s);
    if (pairs && !dst) return launch_variant<V, true, false>(a, s);
    if constexpr (V == 0) {
        if (!pairs && dst) return launch_variant<0, false, true>(a, s);
        return launch_variant<0, true, true>(a, s);
    }
    return cudaErrorInvalidValue;
}
}  // namespace

#define B200_CAT2(a, b) a##b
#define B200_CAT(a, b) B200_CAT2(a, b)

cudaError_t B200_CAT(launch_hist_w, B200_W)(bool uniform, const HistArgs &a, int grid, cudaStream_t s) {
    const size_t smem = (size_t)(uniform ? P_UNIFORM : a.passes.count) * (1u << W) * sizeof(uint32_t);
    if (uniform)
        hist_kernel<W, P_UNIFORM><<<grid, kHistThreads, smem, s>>>(a);
    else
        hist_kernel<W, 0><<<grid, kHistThreads, smem, s>>>(a);
    return cudaGetLastError();
}

cudaError_t B200_CAT(launch_pass_w, B200_W)(int variant, bool pairs, bool dst, const PassArgs &a,
                                            cudaStream_t s) {
    if (dst) variant = 0;
    switch (variant) {
    case 0: return launch_modes<0>(pairs, dst, a, s);
#if B200_W == 8
    case 1: return launch_modes<1>(pairs, dst, a, s);
    case 2: return launch_modes<2>(pairs, dst, a, s);
    case 3: return launch_modes<3>(pairs, dst, a, s);
    case 4: return launch_modes<4>(pairs, dst, a, s);
#endif
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace b200sort
