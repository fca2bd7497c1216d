// hist.cuh -- K1: one read of all keys -> the digit histogram of EVERY pass of the sort, plus
// (last CTA) the exclusive scan that turns each histogram into per-bin output bases.
//
// Replaces histogram()/histogramKernel (SourceCode/Parallel7.cu:318-359), which the reference
// launches once per digit pass with one global atomic per tile x bin and a memset before it,
// and the part of scan() (:485-528) that produced the per-bin bases.
//
// Shared-memory layout: s_hist[pass][bin][lane] -- every lane of a warp owns its own column,
// so the 32 shared atomics of one warp instruction always hit 32 different banks, whatever
// the key distribution (uniform, all-equal, 16 values, Zipf: same speed).  Measured on B200
// (profiles/r01_probe_b200.json): shared atomics on random bins of ONE 256-entry table run
// at ~3.2 cycles per warp instruction (bank conflicts), conflict-free ones at ~1.4; with 4
// digits per key the first layout is atomics-bound at ~0.37 ms for 2^28 keys, the lane-private
// one sits at the DRAM time (~0.17 ms).  match.any aggregation is not used: MATCH.ANY issues
// at ~1 warp instruction per 61 cycles per SM on this part.
#pragma once
#include "common.cuh"

namespace b200sort {

// W: log2 of bins per pass.  P_CT > 0: uniform pass list (pass p = bits [p*W, p*W+W)) known at
// compile time; P_CT == 0: runtime list from args.passes; P_CT == -1: exactly one runtime digit
// (args.passes entry 0) -- the multi-GPU top-digit histogram and b200sort_digit_pass.
template <int W, int P_CT>
__global__ void __launch_bounds__(kHistThreads, 1) hist_kernel(const HistArgs a) {
    constexpr int B = 1 << W;
    constexpr int T = kHistThreads;
    constexpr int U = kHistUnroll;
    extern __shared__ __align__(16) uint32_t s_hist[];  // [P][B][32]
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_warp_tot[32];

    const int P = P_CT > 0 ? P_CT : (P_CT < 0 ? 1 : a.passes.count);
    const uint32_t shift0 = a.passes.shift[0], mask0 = (1u << a.passes.bits[0]) - 1u;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;

    {
        uint4 *z = reinterpret_cast<uint4 *>(s_hist);
        for (int i = tid; i < P * B * 8; i += T) z[i] = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();

    // Side job: clear the look-back descriptors the digit passes will use.
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (uint64_t i = (uint64_t)blockIdx.x * T + tid; i < a.zero_vecs; i += (uint64_t)gridDim.x * T)
            a.zero_ptr[i] = z;
    }

    uint32_t *col = s_hist + lane;  // this lane's column
    // P_CT > 0: digits of W bits at shifts 0, W, 2W, ...; the last one may be narrower (32 % W != 0, or keys that
    // are known to agree above some bit: b200sort_keys_low_bits)
    const uint32_t last_mask = P_CT > 0 ? (1u << a.passes.bits[P_CT > 0 ? P_CT - 1 : 0]) - 1u : 0u;
    auto tally = [&](uint32_t key) {
        if constexpr (P_CT < 0) {
            atomicAdd(col + (((key >> shift0) & mask0) << 5), 1u);
        } else {
#pragma unroll
            for (int p = 0; p < (P_CT > 0 ? P_CT : kMaxPasses); ++p) {
                if (P_CT == 0 && p >= P) break;
                const uint32_t d = P_CT > 0 ? ((key >> (p * W)) & (p == P_CT - 1 ? last_mask : (uint32_t)(B - 1)))
                                            : ((key >> a.passes.shift[p]) & ((1u << a.passes.bits[p]) - 1u));
                atomicAdd(col + ((p * B + d) << 5), 1u);
            }
        }
    };

    // 16-byte aligned body; the <= 3 keys before it and <= 3 after it are tallied by CTA 0.
    const uint64_t addr = reinterpret_cast<uint64_t>(a.keys);
    uint64_t head = ((16u - (uint32_t)(addr & 15u)) & 15u) >> 2;
    if (head > a.n) head = a.n;
    const uint4 *keys4 = reinterpret_cast<const uint4 *>(a.keys + head);
    const uint64_t n_vec = (a.n - head) >> 2;
    const uint64_t tail_start = head + (n_vec << 2);

    for (uint64_t base = (uint64_t)blockIdx.x * (T * U); base < n_vec; base += (uint64_t)gridDim.x * (T * U)) {
        uint4 v[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t idx = base + (uint64_t)u * T + tid;
            ok[u] = idx < n_vec;
            if (ok[u]) v[u] = ld_stream_v4(keys4 + idx);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (ok[u]) {
                tally(v[u].x);
                tally(v[u].y);
                tally(v[u].z);
                tally(v[u].w);
            }
        }
    }
    if (blockIdx.x == 0) {
        if (tid < head) tally(a.keys[tid]);
        if (tail_start + tid < a.n) tally(a.keys[tail_start + tid]);
    }
    __syncthreads();

    // Reduce the 32 lane columns of every (pass, bin) row and add the row to the global histogram.
    for (int row = warp; row < P * B; row += T / 32) {
        const uint32_t sum = __reduce_add_sync(0xffffffffu, s_hist[(row << 5) + lane]);
        if (lane == 0 && sum) atomicAdd(&a.ghist[row], sum);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.done, 1u) == gridDim.x - 1u) ? 1u : 0u;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (int p = 0; p < P; ++p) {
            const uint32_t c = (tid < B) ? ld_relaxed_gpu(&a.ghist[p * B + tid]) : 0u;
            const uint32_t excl = block_exclusive_scan<T>(c, s_warp_tot);
            if (tid < B) a.bin_base[(size_t)(2 * p) * B + tid] = excl;
        }
    }
}

}  // namespace b200sort
