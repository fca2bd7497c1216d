// util_kernels.cuh -- device side of the synthetic workloads and the
// size-independent output checks (sortedness + multiset fingerprint).  Bench/test utilities;
// they generate the same bytes as oracle/radix_oracle.c:oracle_generate.
#pragma once
#include "common.cuh"

namespace b200sort {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

enum { GEN_UNIFORM = 0, GEN_ZIPF = 1, GEN_UNIQUE16 = 2, GEN_ALL_EQUAL = 3, GEN_SORTED = 4,
       GEN_REVERSED = 5, GEN_IOTA = 6 };

__global__ void __launch_bounds__(256) generate_kernel(uint32_t *out, uint64_t first, uint64_t count,
                                                       int kind, uint64_t total,
                                                       const uint32_t *__restrict__ cdf) {
    for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < count;
         j += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first + j;
        uint32_t k = 0;
        switch (kind) {
        case GEN_UNIFORM: k = (uint32_t)(splitmix64(0x5EED0001ULL + i) >> 32); break;
        case GEN_ZIPF: {
            const uint32_t u = (uint32_t)(splitmix64(0x5EED0004ULL + i) >> 32);
            int lo = 0, hi = 65535;  // first r with cdf[r] >= u
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cdf[mid] >= u) hi = mid; else lo = mid + 1;
            }
            k = (uint32_t)(splitmix64((uint64_t)lo) >> 32);
            break;
        }
        case GEN_UNIQUE16: k = (uint32_t)(splitmix64(splitmix64(0x5EED0005ULL + i) & 15ULL) >> 32); break;
        case GEN_ALL_EQUAL: k = 0xDEADBEEFu; break;
        case GEN_SORTED:
        case GEN_REVERSED:
            k = (uint32_t)((i << 32) / (total ? total : 1ULL));  // i < 2^32
            if (kind == GEN_REVERSED) k = ~k;
            break;
        case GEN_IOTA: k = (uint32_t)i; break;
        }
        out[j] = k;
    }
}

// result[0] += #{i >= 1 : keys[i-1] > keys[i]}, result[1] += sum key, result[2] += sum sm64(key),
// result[3] ^= xor sm64(key).
__global__ void __launch_bounds__(256) verify_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                     unsigned long long *result) {
    unsigned long long bad = 0, s = 0, h = 0, x = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t k = keys[i];
        if (i > 0 && keys[i - 1] > k) ++bad;
        const uint64_t m = splitmix64((uint64_t)k);
        s += k;
        h += m;
        x ^= m;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bad += __shfl_xor_sync(0xffffffffu, bad, o);
        s += __shfl_xor_sync(0xffffffffu, s, o);
        h += __shfl_xor_sync(0xffffffffu, h, o);
        x ^= __shfl_xor_sync(0xffffffffu, x, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&result[0], bad);
        atomicAdd(&result[1], s);
        atomicAdd(&result[2], h);
        atomicXor(&result[3], x);
    }
}

// route[i] = #{j < count : (value[j], tie[j]) <= (key[i], i)} in lexicographic order, for `count`
// cuts given as thresholds[0 .. count) = values in [0, 2^32] and thresholds[count .. 2*count) =
// tie indices: a key EQUAL to value[j] counts as >= cut j from local index tie[j] on.  This is the
// destination shard of a key under value splitters (multi-GPU, skewed keys); the tie index lets a
// run of equal keys be cut at a position, in input order.  Cuts must be non-decreasing.
// The result is then used as the KEY of a digit pass that carries the real keys as values.
constexpr int kMaxRouteThresholds = 255;
constexpr int kRouteRegCuts = 8;  // up to this many cuts live in registers (2..9 shards)

// counts (may be null): counts[d] += number of keys routed to d, d in [0, count].
template <bool SMALL>
__global__ void __launch_bounds__(256) route_kernel(const uint32_t *__restrict__ keys, uint64_t n,
                                                    const uint64_t *__restrict__ thresholds, int count,
                                                    uint32_t *__restrict__ route, uint32_t *__restrict__ counts) {
    __shared__ uint64_t s_v[SMALL ? 1 : kMaxRouteThresholds + 1];
    __shared__ uint64_t s_c[SMALL ? 1 : kMaxRouteThresholds + 1];
    __shared__ uint32_t s_hist[SMALL ? kRouteRegCuts + 1 : kMaxRouteThresholds + 1];
    // SMALL: cuts in registers as (32-bit value, tie index); a value of 2^32 ("nothing is at or
    // above") becomes (0xFFFFFFFF, never).  above[j] = keys of this lane at or above cut j.
    uint32_t v[SMALL ? kRouteRegCuts : 1], above[SMALL ? kRouteRegCuts : 1], seen = 0;
    uint64_t c[SMALL ? kRouteRegCuts : 1];
    if constexpr (SMALL) {
#pragma unroll
        for (int j = 0; j < kRouteRegCuts; ++j) {
            const uint64_t value = j < count ? thresholds[j] : ~0ull;
            v[j] = value > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)value;
            c[j] = value > 0xFFFFFFFFull ? ~0ull : thresholds[count + j];
            above[j] = 0;
        }
        if (threadIdx.x <= kRouteRegCuts) s_hist[threadIdx.x] = 0;
    } else {
        for (int j = threadIdx.x; j < count; j += blockDim.x) {
            s_v[j] = thresholds[j];
            s_c[j] = thresholds[count + j];
        }
        for (int j = threadIdx.x; j <= count; j += blockDim.x) s_hist[j] = 0;
    }
    __syncthreads();
    auto dest = [&](uint32_t key, uint64_t i) -> uint32_t {
        uint32_t r = 0;
        if constexpr (SMALL) {
            ++seen;
#pragma unroll
            for (int j = 0; j < kRouteRegCuts; ++j) {
                if (j >= count) break;  // uniform
                uint32_t ge = key > v[j] ? 1u : 0u;
                if (key == v[j]) ge = c[j] <= i ? 1u : 0u;
                r += ge;
                above[j] += ge;
            }
        } else {  // upper bound by bisection
            const uint64_t k = key;
            int lo = 0, hi = count;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (s_v[mid] < k || (s_v[mid] == k && s_c[mid] <= i)) lo = mid + 1; else hi = mid;
            }
            r = (uint32_t)lo;
            if (counts) atomicAdd(&s_hist[r], 1u);
        }
        return r;
    };
    // 16 bytes per lane and two loads in flight: a 4-byte grid-stride loop leaves this kernel
    // latency bound at a fifth of the DRAM bandwidth
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gsize = (uint64_t)gridDim.x * blockDim.x;
    const bool aligned = ((reinterpret_cast<uintptr_t>(keys) | reinterpret_cast<uintptr_t>(route)) & 15u) == 0;
    const uint64_t n4 = aligned ? n >> 2 : 0;
    const uint4 *keys4 = reinterpret_cast<const uint4 *>(keys);
    uint4 *route4 = reinterpret_cast<uint4 *>(route);
    for (uint64_t vi = gtid; vi < n4; vi += 2 * gsize) {
        const uint64_t vj = vi + gsize;
        const uint4 a = ld_stream_v4(keys4 + vi);
        uint4 b = make_uint4(0, 0, 0, 0);
        if (vj < n4) b = ld_stream_v4(keys4 + vj);
        route4[vi] = make_uint4(dest(a.x, 4 * vi), dest(a.y, 4 * vi + 1), dest(a.z, 4 * vi + 2), dest(a.w, 4 * vi + 3));
        if (vj < n4)
            route4[vj] = make_uint4(dest(b.x, 4 * vj), dest(b.y, 4 * vj + 1), dest(b.z, 4 * vj + 2), dest(b.w, 4 * vj + 3));
    }
    for (uint64_t i = 4 * n4 + gtid; i < n; i += gsize) route[i] = dest(keys[i], i);
    if (!counts) return;
    if constexpr (SMALL) {
        // cuts are non-decreasing, so keys routed to d = (keys at or above cut d-1) - (at or above cut d)
        uint32_t prev = __reduce_add_sync(0xffffffffu, seen);
#pragma unroll
        for (int d = 0; d <= kRouteRegCuts; ++d) {
            if (d > count) break;
            const uint32_t next = d < count ? __reduce_add_sync(0xffffffffu, above[d < kRouteRegCuts ? d : 0]) : 0u;
            if ((threadIdx.x & 31) == 0 && prev != next) atomicAdd(&s_hist[d], prev - next);
            prev = next;
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d <= count; d += blockDim.x)
        if (s_hist[d]) atomicAdd(&counts[d], s_hist[d]);
}

// Probe for the fused exchange: copies n uint32 from src to dst (dst may be peer memory) with
// 4-byte (vec = 1) or 16-byte (vec = 4) stores per lane, `chunk` consecutive keys per warp visit
// (chunk = 32 imitates the digit-pass write-out: one 128-byte run per warp store).
__global__ void __launch_bounds__(256) store_probe_kernel(uint32_t *dst, const uint32_t *__restrict__ src, uint64_t n,
                                                          int vec) {
    if (vec == 4) {
        const uint64_t n4 = n >> 2;
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (uint64_t)gridDim.x * blockDim.x)
            reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
    } else {
        for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
            dst[i] = src[i];
    }
}

}  // namespace b200sort
