// hist.cuh -- K1: one read of all keys -> the digit histogram of EVERY pass of the sort, plus
// (last CTA) the exclusive scan that turns each histogram into per-bin output bases.
//
// Replaces histogram()/histogramKernel (SourceCode/Parallel7.cu:318-359), which the reference
// launches once per digit pass with one global atomic per tile x bin and a memset before it,
// and the part of scan() (:485-528) that produced the per-bin bases.
//
// Counting is done in shared memory.  Each pass starts in "plain" mode (one shared atomic per
// key) and switches, per pass and per CTA, to warp-aggregated mode (match.any + one atomic per
// distinct digit in the warp) once a bin of that pass holds more than 1/8 of the keys seen so
// far -- the point where same-address conflicts cost more than the match.  Narrow digits
// (<= 5 bits) start aggregated.
#pragma once
#include "common.cuh"

namespace b200sort {

// W: log2 of bins per pass.  P_CT > 0: uniform pass list (pass p = bits [p*W, p*W+W)) known at
// compile time; P_CT == 0: runtime list from args.passes.
template <int W, int P_CT>
__global__ void __launch_bounds__(kHistThreads) hist_kernel(const HistArgs a) {
    constexpr int B = 1 << W;
    constexpr int T = kHistThreads;
    constexpr int U = kHistUnroll;
    extern __shared__ uint32_t s_hist[];  // [P][B]
    __shared__ uint32_t s_flags;
    __shared__ uint32_t s_last;
    __shared__ uint32_t s_warp_tot[32];

    const int P = P_CT ? P_CT : a.passes.count;
    const uint32_t tid = threadIdx.x;
    const uint32_t lt = lanemask_lt();

    for (int i = tid; i < P * B; i += T) s_hist[i] = 0;
    if (tid == 0) s_flags = a.agg_init;
    __syncthreads();

    // Side job: clear the look-back descriptors the digit passes will use.
    {
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (uint64_t i = (uint64_t)blockIdx.x * T + tid; i < a.zero_vecs; i += (uint64_t)gridDim.x * T)
            a.zero_ptr[i] = z;
    }

    uint32_t flags = a.agg_init;

    auto tally = [&](uint32_t key, uint32_t warp_mask) {
#pragma unroll
        for (int p = 0; p < (P_CT ? P_CT : kMaxPasses); ++p) {
            if (!P_CT && p >= P) break;
            const uint32_t d = P_CT ? ((key >> (p * W)) & (B - 1))
                                    : ((key >> a.passes.shift[p]) & ((1u << a.passes.bits[p]) - 1u));
            if (flags & (1u << p)) {  // warp-uniform
                const uint32_t peers = __match_any_sync(warp_mask, d);
                if ((peers & lt) == 0) atomicAdd(&s_hist[p * B + d], (uint32_t)__popc(peers));
            } else {
                atomicAdd(&s_hist[p * B + d], 1u);
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

    uint32_t iter = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * (T * U); base < n_vec;
         base += (uint64_t)gridDim.x * (T * U), ++iter) {
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
            const uint32_t m = __ballot_sync(0xffffffffu, ok[u]);
            if (ok[u]) {
                tally(v[u].x, m);
                tally(v[u].y, m);
                tally(v[u].z, m);
                tally(v[u].w, m);
            }
        }
        if ((iter & 3u) == 0u) {
            // Re-evaluate the counting mode on the cumulative CTA histogram.
            __syncthreads();
            const uint32_t seen = (iter + 1u) * (uint32_t)(T * U * 4);
            const uint32_t thr = seen >> 3;
            for (int i = tid; i < P * B; i += T)
                if (s_hist[i] > thr) atomicOr(&s_flags, 1u << (i / B));
            __syncthreads();
            flags = s_flags;
        }
    }
    if (blockIdx.x == 0) {
        flags = 0;  // ragged ends: plain atomics, no warp-wide participation needed
        if (tid < head) tally(a.keys[tid], 0);
        if (tail_start + tid < a.n) tally(a.keys[tail_start + tid], 0);
    }
    __syncthreads();

    for (int i = tid; i < P * B; i += T) {
        const uint32_t c = s_hist[i];
        if (c) atomicAdd(&a.ghist[i], c);
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
