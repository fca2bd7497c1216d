// launch.h -- interface between the host driver (b200sort.cu) and the per-width kernel
// translation units (kernels_w.cu compiled once per digit width with -DB200_W=1..8).
#pragma once
#include "common.cuh"

namespace b200sort {

// Digit-pass kernel geometries.  Variant 0 is the default and exists for every width; the
// others are tuning alternatives instantiated for the 8-bit digit only.
struct PassGeometry {
    int threads;
    int items_keys;
    int items_pairs;
    int min_ctas;
};
constexpr int kNumVariants = 5;
constexpr PassGeometry kGeometry[kNumVariants] = {
    {384, 20, 14, 2},
    {512, 15, 10, 2},
    {256, 24, 16, 3},
    {384, 16, 12, 2},
    {512, 12, 8, 2},
};
inline int tile_keys(int variant, bool pairs) {
    const PassGeometry &g = kGeometry[variant];
    return g.threads * (pairs ? g.items_pairs : g.items_keys);
}
// Smallest tile over all variants: bounds the descriptor array when sizing temp storage.
constexpr int kMinTileKeys = 2048;

// true if (width, variant) is instantiated; callers fall back to variant 0 otherwise.
inline bool variant_available(int width, int variant) {
    return variant == 0 || (width == 8 && variant > 0 && variant < kNumVariants);
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

}  // namespace b200sort
