// b200sort.cu -- host driver and C ABI of libb200sort.so (see include/b200sort.h).
//
// One sort =   memset(header)  ->  K1 hist_kernel (all digit histograms + bases, clears the
// look-back descriptors)  ->  one onesweep_pass_kernel launch per digit pass (per <2^30-key
// portion), ping-ponging between the output and one alternate buffer in temp storage.
// Host side of the reference's sortByDevice (SourceCode/Parallel7.cu:530-639) without its
// 88 launches, 88 device synchronisations and 4 host round trips per sort.
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../../include/b200sort.h"
#include "copy_pool.h"
#include "internal.h"
#include "launch.h"
#include "scan.cuh"
#include "util_kernels.cuh"

namespace b200sort {
namespace {

thread_local std::string g_last_error = "";
std::atomic<uint64_t> g_launches{0};

// Tuning parameters: process-wide, read once per call into locals (b200sort_set_param is meant for
// benches and tests; changing a parameter while another thread is inside a sort affects only later calls).
struct Params {
    std::atomic<int> variant{-1};        // -1 = automatic choice (see effective_variant)
    std::atomic<int> portion_tiles{0};   // 0 = as many as fit the 30-bit descriptor value
    std::atomic<int> hist_ctas_per_sm{2};
    std::atomic<int> narrow_variant{-1}; // digit passes of <= 3 bits: -1 = kBallotVariant
    std::atomic<int> safe_rank{0};       // 1 = only kernels whose ranking follows from the PTX memory model
    std::atomic<int> dst_bulk{1};        // digit pass with per-bin destinations, keys only: bulk-copy write-out
    std::atomic<int> host_overlap{1};    // host-pointer path: chunked upload + MSD split + per-bucket download
    std::atomic<int> scan_variant{8};    // tile geometry of b200sort_exclusive_scan (scan.cuh: kScanGeom)
    std::atomic<int> scan_prefetch_tiles{-1}; // exclusive scan: L2 prefetch distance in tiles (-1 = 8 MiB ahead, 0 = off)
    std::atomic<int> prefetch_tiles{-1}; // column sweep: L2 prefetch distance in tiles; -1 = one tile per SM ahead, 0 = off
} g_params;

// Per CUDA ordinal: the sm_100 check, the SM count and the verdict of the RANK_ATOMIC self test
// (-1 = not run yet, 0 = failed -> spec-safe column-sweep kernel, 1 = passed).
constexpr int kMaxDevices = 64;
struct DeviceState {
    std::atomic<int> checked{-1000};
    std::atomic<int> sms{0};
    std::atomic<int> atomic_rank_ok{-1};
};
DeviceState g_dev[kMaxDevices];

DeviceState &dev_state() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return g_dev[dev & (kMaxDevices - 1)];
}
int num_sms() { return dev_state().sms.load(); }
uint32_t prefetch_distance() {
    const int p = g_params.prefetch_tiles;
    return (uint32_t)(p >= 0 ? p : num_sms());
}

int fail(int code, const char *what) {
    g_last_error = std::string(what) + ": " + b200sort_error_string(code);
    return code;
}
int fail_cuda(cudaError_t e, const char *what) {
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return (int)e;
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);        \
    } while (0)

int check_device() {
    DeviceState &st = dev_state();
    if (st.checked != -1000) return st.checked;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        g_last_error = std::string("cudaGetDevice: ") + cudaGetErrorString(e);
        return B200SORT_ENODEVICE;  // not cached: a later call may find a device
    }
    int major = 0, minor = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10) {
        g_last_error = "libb200sort is built for sm_100a only; device is sm_" + std::to_string(major) +
                       std::to_string(minor);
        return B200SORT_ENODEVICE;
    }
    st.sms = sms;
    st.checked = 0;
    return 0;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- digit-pass list ----------------------------------------------------------------------
// nBits 1..8: one kernel pass per reference digit (the last digit is narrower when
// 32 % nBits != 0, exactly like the loop at SourceCode/Baseline1.cu:30).
// nBits 9..16: each reference digit is split into a low sub-digit of ceil(nBits/2) bits and a
// high sub-digit of the remaining bits, both stable -> same order as one wide pass.
// key_bits < 32: the keys are known to agree in their bits >= key_bits (a bucket of an MSD partition), so
// the digits above are skipped.
bool build_pass_list(int nbits, PassList &pl, int key_bits = 32) {
    if (nbits < 1 || nbits > 16 || key_bits < 1 || key_bits > 32) return false;
    const int lo = nbits <= kMaxRadixBits ? nbits : (nbits + 1) / 2;
    pl.width = lo;
    pl.count = 0;
    for (int s = 0; s < key_bits; s += nbits) {
        const int w = std::min(nbits, key_bits - s);
        pl.shift[pl.count] = (uint8_t)s;
        pl.bits[pl.count] = (uint8_t)std::min(w, lo);
        ++pl.count;
        if (w > lo) {
            pl.shift[pl.count] = (uint8_t)(s + lo);
            pl.bits[pl.count] = (uint8_t)(w - lo);
            ++pl.count;
        }
    }
    return true;
}

// ---- temp storage layout ---------------------------------------------------------------------
struct Layout {
    uint64_t n;
    int passes, bins;
    int tile;                // keys per tile of the selected digit-pass kernel
    uint64_t total_tiles;
    uint64_t portion_tiles;  // tiles per launch
    uint64_t portions;
    // zeroed header: tickets[passes * portions] | done | zeros[bins] | ghist[passes * bins]
    size_t off_tickets, off_done, off_zeros, off_ghist, header_bytes;
    size_t off_bin_base;  // [passes][2][bins]
    size_t off_desc;      // [total_tiles][bins]
    size_t desc_bytes;
    size_t off_alt_keys, off_alt_vals;
    size_t total;
};

Layout make_layout(uint64_t n, int passes, int width, bool pairs, bool need_alt, int tile,
                   int portion_tiles_param) {
    Layout L{};
    L.n = n;
    L.passes = passes;
    L.bins = 1 << width;
    L.tile = tile;
    L.total_tiles = (n + tile - 1) / tile;
    uint64_t cap = ((1ull << 30) - 1) / (uint64_t)tile;
    if (portion_tiles_param > 0) cap = std::min<uint64_t>(cap, (uint64_t)portion_tiles_param);
    L.portion_tiles = std::max<uint64_t>(1, std::min<uint64_t>(cap, std::max<uint64_t>(L.total_tiles, 1)));
    L.portions = std::max<uint64_t>(1, (L.total_tiles + L.portion_tiles - 1) / L.portion_tiles);
    size_t off = 0;
    L.off_tickets = off;  off += align_up((size_t)passes * L.portions * 4, 16);
    L.off_done = off;     off += 16;
    L.off_zeros = off;    off += (size_t)L.bins * 4;
    L.off_ghist = off;    off += (size_t)passes * L.bins * 4;
    L.header_bytes = align_up(off, 256);
    off = L.header_bytes;
    L.off_bin_base = off; off += (size_t)passes * 2 * L.bins * 4;
    off = align_up(off, 256);
    L.off_desc = off;
    L.desc_bytes = align_up((size_t)L.total_tiles * L.bins * 4, 256);
    off += L.desc_bytes;
    L.off_alt_keys = off;
    if (need_alt) off += align_up((size_t)n * 4, 256);
    L.off_alt_vals = off;
    if (need_alt && pairs) off += align_up((size_t)n * 4, 256);
    L.total = off;
    return L;
}

int check_device();

// The self test of the current device (once per device and process): three CTAs per SM, half of the
// warps replay random digit patterns with per-lane addends against the ranking atomics while the other
// half hammers the same banks with stores, loads and reductions (see atomic_order_selftest).
int run_selftest() {
    DeviceState &st = dev_state();
    if (st.atomic_rank_ok >= 0) return st.atomic_rank_ok;
    if (check_device() != 0) return 0;  // not cached: no device yet
    uint32_t *d_counter = nullptr;
    uint32_t h = 1;
    if (cudaMalloc(&d_counter, 2 * sizeof(uint32_t)) != cudaSuccess) { cudaGetLastError(); return 0; }
    bool ok = cudaMemset(d_counter, 0, 2 * sizeof(uint32_t)) == cudaSuccess &&
              run_atomic_order_selftest(d_counter, st.sms * 3, 2048, nullptr) == cudaSuccess &&
              cudaMemcpy(&h, d_counter, sizeof(uint32_t), cudaMemcpyDeviceToHost) == cudaSuccess;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaFree(d_counter);
    st.atomic_rank_ok = (ok && h == 0) ? 1 : 0;
    return st.atomic_rank_ok;
}

// Automatic choice: the fastest measured kernel (profiles/r02_sweep_*.jsonl).  Digits of >= 4 bits: the column
// sweep with two ranking chains (kDualVariant; keys and pairs; spec-safe, no self test involved); narrower
// digits: ballot rank.  The atomic-rank kernels of round 1 (variants 1, 10, 35) stay selectable by number.
int effective_variant(int width, bool pairs = false) {
    int v = g_params.variant;
    const int narrow = g_params.narrow_variant;
    (void)pairs;
    if (v < 0) v = (width >= 4) ? kDualVariant : (narrow >= 0 ? narrow : kBallotVariant);
    if (!variant_available(width, v)) v = fallback_variant(width);
    // Atomic-rank kernels need same-address shared atomics of one warp instruction to be applied in lane
    // order (not promised by PTX): they run only on a device that passed the self test, and never when
    // the caller asked for spec-safe ranking -- the column-sweep kernel (lane-private counters, ordered
    // by named barriers) takes over; ballot-only variants (mode 0, no table) are spec-safe as they are.
    if (variant_mode(v) == 1 && (g_params.safe_rank || !run_selftest())) v = kColVariant;
    return v;
}

// Upper bound of the temp storage a sort / digit pass can need, independent of the tuning
// parameters in effect (descriptors sized for the smallest tile).
size_t temp_upper_bound(uint64_t n, int nbits, bool pairs) {
    PassList pl;
    if (!build_pass_list(nbits, pl)) return 0;
    Layout L = make_layout(n, pl.count, pl.width, pairs, true, kMinTileKeys, 0);
    size_t tickets = 0;  // a forced small portion size (tests) multiplies the launches per pass
    if (g_params.portion_tiles.load() > 0)
        tickets = (size_t)pl.count * (((n + kMinTileKeys - 1) / kMinTileKeys) / g_params.portion_tiles + 2) * 4;
    return L.total + align_up(tickets, 256) + 4096;
}

// ---- profiling events --------------------------------------------------------------------------
// Marks accumulate over sorts until b200sort_profile_read() drains them.  tag -1 opens a
// sort; tag 0 closes the histogram kernel; tag p+1 closes digit pass p.
bool g_profile = false;
std::vector<cudaEvent_t> g_events;
std::vector<int> g_event_tags;
int g_events_used = 0;
constexpr int kMaxProfileMarks = 1 << 16;

cudaError_t profile_mark(cudaStream_t s, int tag) {
    if (!g_profile || g_events_used >= kMaxProfileMarks) return cudaSuccess;
    if (g_events_used == (int)g_events.size()) {
        cudaEvent_t ev;
        cudaError_t e = cudaEventCreate(&ev);
        if (e != cudaSuccess) return e;
        g_events.push_back(ev);
        g_event_tags.push_back(0);
    }
    g_event_tags[g_events_used] = tag;
    return cudaEventRecord(g_events[g_events_used++], s);
}

// ---- kernel dispatch by width --------------------------------------------------------------------
cudaError_t launch_hist(int width, bool uniform, const HistArgs &a, int grid, cudaStream_t s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (width) {
    case 1: return launch_hist_w1(uniform, a, grid, s);
    case 2: return launch_hist_w2(uniform, a, grid, s);
    case 3: return launch_hist_w3(uniform, a, grid, s);
    case 4: return launch_hist_w4(uniform, a, grid, s);
    case 5: return launch_hist_w5(uniform, a, grid, s);
    case 6: return launch_hist_w6(uniform, a, grid, s);
    case 7: return launch_hist_w7(uniform, a, grid, s);
    case 8: return launch_hist_w8(uniform, a, grid, s);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_pass(int width, int variant, bool pairs, bool dst, const PassArgs &a, cudaStream_t s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    switch (width) {
    case 1: return launch_pass_w1(variant, pairs, dst, a, s);
    case 2: return launch_pass_w2(variant, pairs, dst, a, s);
    case 3: return launch_pass_w3(variant, pairs, dst, a, s);
    case 4: return launch_pass_w4(variant, pairs, dst, a, s);
    case 5: return launch_pass_w5(variant, pairs, dst, a, s);
    case 6: return launch_pass_w6(variant, pairs, dst, a, s);
    case 7: return launch_pass_w7(variant, pairs, dst, a, s);
    case 8: return launch_pass_w8(variant, pairs, dst, a, s);
    }
    return cudaErrorInvalidValue;
}

int hist_grid(uint64_t n, int passes, int width) {
    const uint64_t per_cta = (uint64_t)kHistThreads * kHistUnroll * 4;
    const uint64_t want = std::max<uint64_t>(1, (n + per_cta - 1) / per_cta);
    const size_t smem = hist_smem_bytes(passes, width) + 1024;
    int ctas = (int)std::min<size_t>((size_t)g_params.hist_ctas_per_sm, (227u * 1024u) / smem);
    ctas = std::max(1, std::min(ctas, 2048 / kHistThreads));
    return (int)std::min<uint64_t>(want, (uint64_t)num_sms() * ctas);
}

bool ranges_overlap(const void *a, const void *b, uint64_t bytes) {
    const uintptr_t x = (uintptr_t)a, y = (uintptr_t)b;
    return x < y + bytes && y < x + bytes;
}

// ---- the sort ------------------------------------------------------------------------------------
int run_sort(const uint32_t *kin, const uint32_t *vin, uint64_t n, uint32_t *kout, uint32_t *vout,
             void *temp, size_t temp_bytes, int nbits, cudaStream_t stream, int key_bits = 32) {
    const bool pairs = (vin != nullptr) || (vout != nullptr);
    PassList pl;
    if (!build_pass_list(nbits, pl, key_bits)) return fail(B200SORT_EINVAL, "nBits must be in 1..16");
    if (n > 0xFFFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    if (n == 0) return B200SORT_OK;
    if (!kin || !kout || (pairs && (!vin || !vout))) return fail(B200SORT_EINVAL, "null buffer");
    if (ranges_overlap(kin, kout, n * 4) || (pairs && ranges_overlap(vin, vout, n * 4)))
        return fail(B200SORT_EALIAS, "output overlaps input");
    int rc = check_device();
    if (rc) return rc;

    const int variant = effective_variant(pl.width, pairs);
    const int tile = tile_keys(variant, pairs);
    const Layout L = make_layout(n, pl.count, pl.width, pairs, true, tile, g_params.portion_tiles);
    if (!temp || ((uintptr_t)temp & 255u)) return fail(B200SORT_ETEMP, "temp storage must be 256-byte aligned");
    if (temp_bytes < L.total) return fail(B200SORT_ETEMP, "temp storage too small");

    char *base = static_cast<char *>(temp);
    uint32_t *done = reinterpret_cast<uint32_t *>(base + L.off_done);
    uint32_t *ghist = reinterpret_cast<uint32_t *>(base + L.off_ghist);
    uint32_t *bin_base = reinterpret_cast<uint32_t *>(base + L.off_bin_base);
    uint32_t *desc = reinterpret_cast<uint32_t *>(base + L.off_desc);
    uint32_t *alt_keys = reinterpret_cast<uint32_t *>(base + L.off_alt_keys);
    uint32_t *alt_vals = reinterpret_cast<uint32_t *>(base + L.off_alt_vals);

    CU(cudaMemsetAsync(base, 0, L.header_bytes, stream));
    CU(profile_mark(stream, -1));

    HistArgs h{};
    h.keys = kin;
    h.n = n;
    h.ghist = ghist;
    h.bin_base = bin_base;
    h.done = done;
    h.zero_ptr = reinterpret_cast<uint4 *>(desc);
    h.zero_vecs = L.desc_bytes / 16;
    h.passes = pl;
    // the compile-time form of K1 covers ceil(32 / nBits) digits at shifts p * nBits, the last one of any width
    const bool k1_fixed = nbits <= kMaxRadixBits && pl.count == (32 + nbits - 1) / nbits;
    CU(launch_hist(pl.width, k1_fixed, h, hist_grid(n, pl.count, pl.width), stream));
    CU(profile_mark(stream, 0));

    const uint64_t portion_keys = L.portion_tiles * (uint64_t)tile;
    for (int p = 0; p < pl.count; ++p) {
        const bool to_out = ((pl.count - 1 - p) & 1) == 0;
        const uint32_t *src_k = (p == 0) ? kin : (to_out ? alt_keys : kout);
        const uint32_t *src_v = (p == 0) ? vin : (to_out ? alt_vals : vout);
        uint32_t *dst_k = to_out ? kout : alt_keys;
        uint32_t *dst_v = to_out ? vout : alt_vals;
        for (uint64_t q = 0; q < L.portions; ++q) {
            const uint64_t first = q * portion_keys;
            const uint64_t count = std::min<uint64_t>(portion_keys, n - first);
            PassArgs a{};
            a.keys_in = src_k + first;
            a.vals_in = pairs ? src_v + first : nullptr;
            a.keys_out = dst_k;
            a.vals_out = pairs ? dst_v : nullptr;
            a.bin_base = bin_base + ((size_t)(2 * p) + (q & 1)) * L.bins;
            a.carry_out = (q + 1 < L.portions) ? bin_base + ((size_t)(2 * p) + ((q + 1) & 1)) * L.bins : nullptr;
            a.desc = desc + (size_t)q * L.portion_tiles * L.bins;
            a.ticket = reinterpret_cast<uint32_t *>(base + L.off_tickets) + (size_t)p * L.portions + q;
            a.bin_dst = nullptr;
            a.n = (uint32_t)count;
            a.num_tiles = (uint32_t)((count + tile - 1) / tile);
            a.shift = pl.shift[p];
            a.mask = (1u << pl.bits[p]) - 1u;
            a.parity = (uint32_t)(p & 1);
            a.prefetch = prefetch_distance();
            // A digit narrower than the others (the last one when nBits does not divide the key width) runs the kernel
            // of its own width: a smaller counter table.  Bases, carries and descriptors are indexed by bin and keep
            // the stride of the widest digit, which they fit.
            const int w = pl.width >= 4 ? std::max<int>(pl.bits[p], 4) : pl.width;
            CU(launch_pass(w, variant, pairs, false, a, stream));
        }
        CU(profile_mark(stream, p + 1));
    }
    return B200SORT_OK;
}

// ---- host-pointer wrapper state --------------------------------------------------------------------
// Pageable host arrays (what the reference's main() passes: plain malloc, SourceCode/Parallel7.cu:
// 712-715) are moved through a ring of pinned staging buffers: a small pool of host threads copies
// chunk k+1 into/out of pinned memory while the DMA engine moves chunk k.  Pinned (or registered)
// arrays are copied directly.
constexpr size_t kStageChunk = 32u << 20;  // bytes per staging buffer
constexpr int kStageSlots = 3;

struct HostCtx {
    std::mutex mu;
    cudaStream_t stream = nullptr;
    void *d_buf = nullptr;  // [keys_in | keys_out | vals_in | vals_out | temp]
    size_t d_bytes = 0;
    void *stage[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr};
    CopyPool *pool = nullptr;
    // overlapped path (sort_host_overlapped): a copy stream, one event per upload chunk / bucket, 16 host counters
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> events;
    uint32_t *h_small = nullptr;  // pinned
} g_host;

int ensure_host_ctx(size_t bytes) {
    if (!g_host.stream) CU(cudaStreamCreateWithFlags(&g_host.stream, cudaStreamNonBlocking));
    if (g_host.d_bytes < bytes) {
        if (g_host.d_buf) CU(cudaFree(g_host.d_buf));
        g_host.d_buf = nullptr;
        g_host.d_bytes = 0;
        cudaError_t e = cudaMalloc(&g_host.d_buf, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(B200SORT_ENOMEM, "cudaMalloc of device staging buffers");
        }
        g_host.d_bytes = bytes;
    }
    return 0;
}

int ensure_staging() {
    for (int i = 0; i < kStageSlots; ++i) {
        if (!g_host.stage[i]) {
            cudaError_t e = cudaMallocHost(&g_host.stage[i], kStageChunk);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return fail(B200SORT_ENOMEM, "cudaMallocHost of pinned staging buffers");
            }
        }
        if (!g_host.stage_ev[i]) CU(cudaEventCreateWithFlags(&g_host.stage_ev[i], cudaEventDisableTiming));
    }
    if (!g_host.pool) g_host.pool = new CopyPool();
    return 0;
}

int ensure_overlap_ctx(size_t num_events) {
    if (!g_host.copy_stream) CU(cudaStreamCreateWithFlags(&g_host.copy_stream, cudaStreamNonBlocking));
    while (g_host.events.size() < num_events) {
        cudaEvent_t ev;
        CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        g_host.events.push_back(ev);
    }
    if (!g_host.h_small) {
        if (cudaMallocHost(&g_host.h_small, 4096) != cudaSuccess) {
            cudaGetLastError();
            return fail(B200SORT_ENOMEM, "cudaMallocHost");
        }
    }
    return 0;
}

bool is_pageable(const void *p) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}

// host -> device on stream `s`, asynchronous as far as the source allows, in chunks of kStageChunk bytes;
// after_chunk(offset, length) is called once the copy of a chunk has been enqueued (the overlapped path
// records an event there and queues the chunk's histogram on the compute stream).
template <typename F>
int upload(void *d_dst, const void *h_src, size_t bytes, cudaStream_t s, F after_chunk) {
    const bool staged = bytes >= (8u << 20) && is_pageable(h_src);
    if (staged) {
        int rc = ensure_staging();
        if (rc) return rc;
    }
    size_t off = 0;
    for (int c = 0; off < bytes; ++c, off += kStageChunk) {
        const int slot = c % kStageSlots;
        const size_t len = std::min(kStageChunk, bytes - off);
        const void *src = static_cast<const char *>(h_src) + off;
        if (staged) {
            CU(cudaEventSynchronize(g_host.stage_ev[slot]));  // the DMA that last used this slot is done
            g_host.pool->copy(g_host.stage[slot], src, len);
            src = g_host.stage[slot];
        }
        CU(cudaMemcpyAsync(static_cast<char *>(d_dst) + off, src, len, cudaMemcpyHostToDevice, s));
        if (staged) CU(cudaEventRecord(g_host.stage_ev[slot], s));
        int rc = after_chunk(off, len);
        if (rc) return rc;
    }
    return 0;
}
int upload(void *d_dst, const void *h_src, size_t bytes) {
    return upload(d_dst, h_src, bytes, g_host.stream, [](size_t, size_t) { return 0; });
}

// device -> host on stream `s`; returns after the data is in h_dst when staging is used
int download(void *h_dst, const void *d_src, size_t bytes, cudaStream_t s = nullptr) {
    if (!s) s = g_host.stream;
    if (bytes == 0) return 0;
    if (bytes < (8u << 20) || !is_pageable(h_dst)) {
        CU(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, s));
        return 0;
    }
    int rc = ensure_staging();
    if (rc) return rc;
    const int chunks = (int)((bytes + kStageChunk - 1) / kStageChunk);
    auto issue = [&](int c) -> cudaError_t {
        const size_t off = (size_t)c * kStageChunk;
        const size_t len = std::min(kStageChunk, bytes - off);
        cudaError_t e = cudaMemcpyAsync(g_host.stage[c % kStageSlots], static_cast<const char *>(d_src) + off, len,
                                        cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return e;
        return cudaEventRecord(g_host.stage_ev[c % kStageSlots], s);
    };
    for (int c = 0; c < std::min(chunks, kStageSlots - 1); ++c) CU(issue(c));
    for (int c = 0; c < chunks; ++c) {
        const int ahead = c + kStageSlots - 1;
        if (ahead < chunks) CU(issue(ahead));  // its slot was drained by the host copy of chunk c-1
        const size_t off = (size_t)c * kStageChunk;
        const size_t len = std::min(kStageChunk, bytes - off);
        CU(cudaEventSynchronize(g_host.stage_ev[c % kStageSlots]));
        g_host.pool->copy(static_cast<char *>(h_dst) + off, g_host.stage[c % kStageSlots], len);
    }
    return 0;
}

// ---- host-pointer path with the copies overlapped ----------------------------------
// sortByDevice brackets its digit loop with one H2D and one D2H of the whole array (Parallel7.cu:549, :624).
// Here, for large arrays:
//   * the upload runs in chunks on a copy stream; the histogram of the top kMsdBits bits of every chunk is
//     accumulated on the compute stream as soon as the chunk has landed (K1 on a key range);
//   * one stable digit pass on those top bits splits the keys into 2^kMsdBits buckets (MSD step);
//   * every bucket is then sorted on its remaining 32 - kMsdBits bits (LSD), back into the input buffer, and
//     downloaded on the copy stream while the next bucket is being sorted.
// The order is the same (a stable MSD split followed by stable LSD sorts of the buckets is a stable sort);
// the extra digit pass hides behind the download, which no longer waits for the whole sort.
constexpr int kMsdBits = 4;
constexpr uint64_t kOverlapMinKeys = 1ull << 24;

int sort_host_overlapped(const uint32_t *hk_in, const uint32_t *hv_in, uint64_t n, uint32_t *hk_out, uint32_t *hv_out,
                         bool pairs, uint32_t *dk_in, uint32_t *dk_out, uint32_t *dv_in, uint32_t *dv_out, void *temp,
                         size_t temp_bytes, int nbits) {
    constexpr int B = 1 << kMsdBits;
    const size_t bytes = (size_t)n * 4;
    const size_t chunks = (bytes + kStageChunk - 1) / kStageChunk;
    int rc = ensure_overlap_ctx(chunks + B + 2);
    if (rc) return rc;
    cudaStream_t s = g_host.stream, sc = g_host.copy_stream;

    const int variant = effective_variant(kMsdBits, pairs);
    const int tile = tile_keys(variant, pairs);
    const Layout L = make_layout(n, 1, kMsdBits, pairs, false, tile, g_params.portion_tiles);
    if (temp_bytes < L.total) return fail(B200SORT_ETEMP, "temp storage too small");
    char *base = static_cast<char *>(temp);
    uint32_t *done = reinterpret_cast<uint32_t *>(base + L.off_done);  // [0] counts the CTAs of the LAST chunk
    uint32_t *ghist = reinterpret_cast<uint32_t *>(base + L.off_ghist);
    uint32_t *bin_base = reinterpret_cast<uint32_t *>(base + L.off_bin_base);
    uint32_t *desc = reinterpret_cast<uint32_t *>(base + L.off_desc);
    CU(cudaMemsetAsync(base, 0, L.header_bytes, s));
    CU(cudaMemsetAsync(done + 1, 0x40, 4, s));  // [1]: a counter that never reaches a grid size (earlier chunks)

    // ---- upload + per-chunk histogram
    size_t c = 0;
    rc = upload(dk_in, hk_in, bytes, sc, [&](size_t off, size_t len) -> int {
        cudaEvent_t ev = g_host.events[c];
        CU(cudaEventRecord(ev, sc));
        CU(cudaStreamWaitEvent(s, ev, 0));
        const bool last = (off + len == bytes);
        HistArgs h{};
        h.keys = dk_in + off / 4;
        h.n = len / 4;
        h.ghist = ghist;
        h.bin_base = bin_base;
        h.done = last ? done : done + 1;  // only the last chunk's last CTA turns the histogram into bin bases
        h.zero_ptr = last ? reinterpret_cast<uint4 *>(desc) : nullptr;
        h.zero_vecs = last ? L.desc_bytes / 16 : 0;
        h.passes.count = 1;
        h.passes.width = kMsdBits;
        h.passes.shift[0] = (uint8_t)(32 - kMsdBits);
        h.passes.bits[0] = (uint8_t)kMsdBits;
        CU(launch_hist(kMsdBits, false, h, hist_grid(h.n, 1, kMsdBits), s));
        ++c;
        return 0;
    });
    if (rc) return rc;
    if (pairs) {
        if ((rc = upload(dv_in, hv_in, bytes, sc, [](size_t, size_t) { return 0; })) != 0) return rc;
        CU(cudaEventRecord(g_host.events[chunks], sc));
        CU(cudaStreamWaitEvent(s, g_host.events[chunks], 0));
    }
    CU(cudaMemcpyAsync(g_host.h_small, ghist, B * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));

    // ---- MSD step: one stable digit pass on the top kMsdBits bits, dk_in -> dk_out
    const uint64_t portion_keys = L.portion_tiles * (uint64_t)tile;
    for (uint64_t q = 0; q < L.portions; ++q) {
        const uint64_t first = q * portion_keys;
        const uint64_t count = std::min<uint64_t>(portion_keys, n - first);
        PassArgs a{};
        a.keys_in = dk_in + first;
        a.vals_in = pairs ? dv_in + first : nullptr;
        a.keys_out = dk_out;
        a.vals_out = pairs ? dv_out : nullptr;
        a.bin_base = bin_base + (q & 1) * L.bins;
        a.carry_out = (q + 1 < L.portions) ? bin_base + ((q + 1) & 1) * L.bins : nullptr;
        a.desc = desc + (size_t)q * L.portion_tiles * L.bins;
        a.ticket = reinterpret_cast<uint32_t *>(base + L.off_tickets) + q;
        a.bin_dst = nullptr;
        a.n = (uint32_t)count;
        a.num_tiles = (uint32_t)((count + tile - 1) / tile);
        a.shift = 32 - kMsdBits;
        a.mask = B - 1;
        a.parity = 0;
        a.prefetch = prefetch_distance();
        CU(launch_pass(kMsdBits, variant, pairs, false, a, s));
    }
    CU(cudaStreamSynchronize(s));  // bucket sizes are on the host now (the upload is complete as well)

    // ---- per bucket: LSD sort on the remaining bits into the input buffer, then download
    uint64_t offs[B + 1];
    offs[0] = 0;
    for (int b = 0; b < B; ++b) offs[b + 1] = offs[b] + g_host.h_small[b];
    if (offs[B] != n) return fail(B200SORT_EINVAL, "internal: bucket sizes do not add up");
    for (int b = 0; b < B; ++b) {
        const uint64_t cnt = offs[b + 1] - offs[b];
        if (cnt) {
            rc = run_sort(dk_out + offs[b], pairs ? dv_out + offs[b] : nullptr, cnt, dk_in + offs[b],
                          pairs ? dv_in + offs[b] : nullptr, temp, temp_bytes, nbits, s, 32 - kMsdBits);
            if (rc) return rc;
        }
        CU(cudaEventRecord(g_host.events[chunks + 1 + b], s));
    }
    for (int b = 0; b < B; ++b) {
        const uint64_t cnt = offs[b + 1] - offs[b];
        if (!cnt) continue;
        CU(cudaStreamWaitEvent(sc, g_host.events[chunks + 1 + b], 0));
        if ((rc = download(hk_out + offs[b], dk_in + offs[b], cnt * 4, sc)) != 0) return rc;
        if (pairs && (rc = download(hv_out + offs[b], dv_in + offs[b], cnt * 4, sc)) != 0) return rc;
    }
    CU(cudaStreamSynchronize(sc));
    CU(cudaStreamSynchronize(s));
    return B200SORT_OK;
}

int sort_host(const uint32_t *hk_in, const uint32_t *hv_in, uint64_t n, uint32_t *hk_out,
              uint32_t *hv_out, int nbits, int block_size, bool pairs) {
    PassList pl;
    if (!build_pass_list(nbits, pl)) return fail(B200SORT_EINVAL, "nBits must be in 1..16");
    if (block_size <= 0) return fail(B200SORT_EINVAL, "blockSize must be positive");
    if (n > 0xFFFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    if (n == 0) return B200SORT_OK;
    if (!hk_in || !hk_out || (pairs && (!hv_in || !hv_out))) return fail(B200SORT_EINVAL, "null buffer");
    int rc = check_device();
    if (rc) return rc;

    std::lock_guard<std::mutex> lock(g_host.mu);
    const size_t arr = align_up((size_t)n * 4, 256);
    const size_t temp_bytes = temp_upper_bound(n, nbits, pairs);
    const size_t arrays = pairs ? 4 : 2;
    rc = ensure_host_ctx(arrays * arr + temp_bytes);
    if (rc) return rc;
    char *b = static_cast<char *>(g_host.d_buf);
    uint32_t *dk_in = reinterpret_cast<uint32_t *>(b);
    uint32_t *dk_out = reinterpret_cast<uint32_t *>(b + arr);
    uint32_t *dv_in = pairs ? reinterpret_cast<uint32_t *>(b + 2 * arr) : nullptr;
    uint32_t *dv_out = pairs ? reinterpret_cast<uint32_t *>(b + 3 * arr) : nullptr;
    void *temp = b + arrays * arr;
    cudaStream_t s = g_host.stream;

    if (n >= kOverlapMinKeys && nbits == 8 && g_params.host_overlap)
        return sort_host_overlapped(hk_in, hv_in, n, hk_out, hv_out, pairs, dk_in, dk_out, dv_in, dv_out, temp,
                                    temp_bytes, nbits);
    if ((rc = upload(dk_in, hk_in, (size_t)n * 4)) != 0) return rc;
    if (pairs && (rc = upload(dv_in, hv_in, (size_t)n * 4)) != 0) return rc;
    rc = run_sort(dk_in, dv_in, n, dk_out, dv_out, temp, temp_bytes, nbits, s);
    if (rc) return rc;
    if ((rc = download(hk_out, dk_out, (size_t)n * 4)) != 0) return rc;
    if (pairs && (rc = download(hv_out, dv_out, (size_t)n * 4)) != 0) return rc;
    CU(cudaStreamSynchronize(s));
    return B200SORT_OK;
}

}  // namespace

std::atomic<int> g_mgpu_balance_permille{1200};

int set_error(int code, const char *what) { return fail(code, what); }
int set_cuda_error(cudaError_t e, const char *what) { return fail_cuda(e, what); }
void set_error_message(const char *message) { g_last_error = message ? message : ""; }

}  // namespace b200sort

namespace b200sort {
namespace {
template <int THREADS, int VECS>
cudaError_t launch_scan_persistent(const uint32_t *d_in, uint32_t *d_out, uint64_t n, uint64_t *desc, uint32_t tiles,
                                   cudaStream_t s) {
    auto kernel = exclusive_scan_persistent_kernel<THREADS, VECS>;
    static std::atomic<int> resident[kMaxDevices] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    int &slot = *reinterpret_cast<int *>(&resident[dev & (kMaxDevices - 1)]);
    if (slot == 0) {  // every CTA must be resident: tiles spin on tiles held by other CTAs of the launch
        int per_sm = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, 0);
        if (e != cudaSuccess) return e;
        slot = std::max(1, per_sm) * std::max(1, num_sms());
    }
    kernel<<<std::min<unsigned>(tiles, (unsigned)slot), THREADS, 0, s>>>(d_in, d_out, n, desc, tiles);
    return cudaGetLastError();
}

}  // namespace
}  // namespace b200sort

using namespace b200sort;

// =====================================================================================================
extern "C" {

int b200sort_version(void) { return B200SORT_VERSION; }

const char *b200sort_error_string(int code) {
    switch (code) {
    case B200SORT_OK: return "ok";
    case B200SORT_EINVAL: return "invalid argument";
    case B200SORT_ETOOBIG: return "n exceeds 2^32-1 keys";
    case B200SORT_ETEMP: return "temp storage too small or misaligned";
    case B200SORT_EALIAS: return "output aliases input";
    case B200SORT_ENODEVICE: return "no usable sm_100 device";
    case B200SORT_ENOMEM: return "out of device memory";
    case B200SORT_ENOPEER: return "no peer access between the selected devices";
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "unknown error";
}

const char *b200sort_last_error_string(void) { return g_last_error.c_str(); }

uint64_t b200sort_launch_count(void) { return g_launches.load(); }

int b200sort_num_passes(int nBits) {
    PassList pl;
    return build_pass_list(nBits, pl) ? pl.count : B200SORT_EINVAL;
}

uint64_t b200sort_algorithmic_bytes(uint64_t n, int nBits, int pairs) {
    PassList pl;
    if (!build_pass_list(nBits, pl)) return 0;
    const uint64_t P = (uint64_t)pl.count;
    return 4ull * n * (pairs ? 4 * P + 1 : 2 * P + 1);
}

int b200sort_tile_keys(int pairs) { return tile_keys(effective_variant(8, pairs != 0), pairs != 0); }

size_t b200sort_temp_bytes(uint64_t n, int nBits, int pairs) {
    return temp_upper_bound(n, nBits, pairs != 0);
}

int b200sort_set_param(const char *name, int value) {
    if (!name) return B200SORT_EINVAL;
    if (!strcmp(name, "variant")) {
        if (value < -1 || value >= kNumVariants) return B200SORT_EINVAL;
        g_params.variant = value;
        return 0;
    }
    if (!strcmp(name, "portion_tiles")) {
        if (value < 0) return B200SORT_EINVAL;
        g_params.portion_tiles = value;
        return 0;
    }
    if (!strcmp(name, "narrow_variant")) {
        if (value != -1 && value != kBallotVariant && value != kBallotSmallVariant) return B200SORT_EINVAL;
        g_params.narrow_variant = value;
        return 0;
    }
    if (!strcmp(name, "mgpu_balance_permille")) {
        if (value < 0) return B200SORT_EINVAL;
        g_mgpu_balance_permille = value;
        return 0;
    }
    if (!strcmp(name, "hist_ctas_per_sm")) {
        if (value < 1 || value > 4) return B200SORT_EINVAL;
        g_params.hist_ctas_per_sm = value;
        return 0;
    }
    if (!strcmp(name, "safe_rank")) {
        if (value != 0 && value != 1) return B200SORT_EINVAL;
        g_params.safe_rank = value;
        return 0;
    }
    if (!strcmp(name, "host_overlap")) {
        if (value != 0 && value != 1) return B200SORT_EINVAL;
        g_params.host_overlap = value;
        return 0;
    }
    if (!strcmp(name, "dst_bulk")) {
        if (value != 0 && value != 1) return B200SORT_EINVAL;
        g_params.dst_bulk = value;
        return 0;
    }
    if (!strcmp(name, "scan_variant")) {
        if (value < 0 || value >= kScanNumVariants) return B200SORT_EINVAL;
        g_params.scan_variant = value;
        return 0;
    }
    if (!strcmp(name, "prefetch_tiles")) {
        if (value < -1 || value > (1 << 20)) return B200SORT_EINVAL;
        g_params.prefetch_tiles = value;
        return 0;
    }
    if (!strcmp(name, "scan_prefetch_tiles")) {
        if (value < -1 || value > (1 << 20)) return B200SORT_EINVAL;
        g_params.scan_prefetch_tiles = value;
        return 0;
    }
    return B200SORT_EINVAL;
}

int b200sort_get_param(const char *name) {
    if (!name) return B200SORT_EINVAL;
    if (!strcmp(name, "variant")) return g_params.variant;
    if (!strcmp(name, "portion_tiles")) return g_params.portion_tiles;
    if (!strcmp(name, "hist_ctas_per_sm")) return g_params.hist_ctas_per_sm;
    if (!strcmp(name, "mgpu_balance_permille")) return g_mgpu_balance_permille.load();
    if (!strcmp(name, "num_variants")) return kNumVariants;
    if (!strcmp(name, "tuning_build")) return kTuningBuild ? 1 : 0;
    if (!strcmp(name, "safe_rank")) return g_params.safe_rank;
    if (!strcmp(name, "scan_variant")) return g_params.scan_variant;
    if (!strcmp(name, "prefetch_tiles")) return g_params.prefetch_tiles;
    if (!strcmp(name, "scan_prefetch_tiles")) return g_params.scan_prefetch_tiles;
    if (!strcmp(name, "host_overlap")) return g_params.host_overlap;
    if (!strcmp(name, "dst_bulk")) return g_params.dst_bulk;
    if (!strcmp(name, "rank_mode")) return check_device() ? -1 : variant_mode(effective_variant(8));
    if (!strcmp(name, "effective_variant")) return check_device() ? std::max<int>(g_params.variant.load(), 0) : effective_variant(8);
    if (!strcmp(name, "atomic_rank_ok")) return check_device() ? -1 : run_selftest();
    return B200SORT_EINVAL;
}

int b200sort_keys(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp, size_t temp_bytes,
                  int nBits, void *stream) {
    return run_sort(d_in, nullptr, n, d_out, nullptr, d_temp, temp_bytes, nBits, (cudaStream_t)stream);
}

int b200sort_pairs(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                   uint32_t *d_keys_out, uint32_t *d_vals_out, void *d_temp, size_t temp_bytes,
                   int nBits, void *stream) {
    if (n > 0 && (!d_vals_in || !d_vals_out)) return fail(B200SORT_EINVAL, "null value buffer");
    return run_sort(d_keys_in, d_vals_in, n, d_keys_out, d_vals_out, d_temp, temp_bytes, nBits,
                    (cudaStream_t)stream);
}

int b200sort_keys_low_bits(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp, size_t temp_bytes,
                           int nBits, int key_bits, void *stream) {
    if (key_bits < 1 || key_bits > 32) return fail(B200SORT_EINVAL, "key_bits must be in 1..32");
    return run_sort(d_in, nullptr, n, d_out, nullptr, d_temp, temp_bytes, nBits, (cudaStream_t)stream, key_bits);
}

int b200sort_pairs_low_bits(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                            uint32_t *d_keys_out, uint32_t *d_vals_out, void *d_temp, size_t temp_bytes,
                            int nBits, int key_bits, void *stream) {
    if (key_bits < 1 || key_bits > 32) return fail(B200SORT_EINVAL, "key_bits must be in 1..32");
    if (n > 0 && (!d_vals_in || !d_vals_out)) return fail(B200SORT_EINVAL, "null value buffer");
    return run_sort(d_keys_in, d_vals_in, n, d_keys_out, d_vals_out, d_temp, temp_bytes, nBits,
                    (cudaStream_t)stream, key_bits);
}

int b200sort_keys_host(const uint32_t *h_in, uint64_t n, uint32_t *h_out, int nBits, int blockSize) {
    return sort_host(h_in, nullptr, n, h_out, nullptr, nBits, blockSize, false);
}

int b200sort_pairs_host(const uint32_t *h_keys_in, const uint32_t *h_vals_in, uint64_t n,
                        uint32_t *h_keys_out, uint32_t *h_vals_out, int nBits, int blockSize) {
    return sort_host(h_keys_in, h_vals_in, n, h_keys_out, h_vals_out, nBits, blockSize, true);
}

int b200sort_warmup(uint64_t max_n, int pairs) {
    if (max_n > 0xFFFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    int rc = check_device();
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(g_host.mu);
    if (kTuningBuild) run_selftest();  // the atomic-rank kernels (tuning build only) run after this verdict
    const size_t arr = align_up((size_t)std::max<uint64_t>(max_n, 1) * 4, 256);
    rc = ensure_host_ctx((pairs ? 4 : 2) * arr + temp_upper_bound(std::max<uint64_t>(max_n, 1), 8, pairs != 0));
    if (rc) return rc;
    if ((rc = ensure_staging()) != 0) return rc;
    if ((rc = ensure_overlap_ctx(((size_t)max_n * 4 + kStageChunk - 1) / kStageChunk + (1 << kMsdBits) + 2)) != 0) return rc;
    // load the kernels of the default path: one tiny sort of each kind through the device buffers
    char *b = static_cast<char *>(g_host.d_buf);
    uint32_t *k_in = reinterpret_cast<uint32_t *>(b), *k_out = reinterpret_cast<uint32_t *>(b + arr);
    void *temp = b + (pairs ? 4 : 2) * arr;
    const size_t temp_bytes = temp_upper_bound(std::max<uint64_t>(max_n, 1), 8, pairs != 0);
    const uint64_t m = std::min<uint64_t>(std::max<uint64_t>(max_n, 1), 1u << 16);
    CU(cudaMemsetAsync(k_in, 0, m * 4, g_host.stream));
    if ((rc = run_sort(k_in, nullptr, m, k_out, nullptr, temp, temp_bytes, 8, g_host.stream)) != 0) return rc;
    if (pairs) {
        uint32_t *v_in = reinterpret_cast<uint32_t *>(b + 2 * arr), *v_out = reinterpret_cast<uint32_t *>(b + 3 * arr);
        CU(cudaMemsetAsync(v_in, 0, m * 4, g_host.stream));
        if ((rc = run_sort(k_in, v_in, m, k_out, v_out, temp, temp_bytes, 8, g_host.stream)) != 0) return rc;
    }
    if (max_n >= kOverlapMinKeys) {  // the MSD step of the overlapped host path uses the 4-bit kernels
        if ((rc = run_sort(k_in, nullptr, m, k_out, nullptr, temp, temp_bytes, 4, g_host.stream)) != 0) return rc;
    }
    CU(cudaStreamSynchronize(g_host.stream));
    return B200SORT_OK;
}

int b200sort_shutdown(void) {
    b200sort_mgpu_shutdown();
    std::lock_guard<std::mutex> lock(g_host.mu);
    if (g_host.d_buf) cudaFree(g_host.d_buf);
    g_host.d_buf = nullptr;
    g_host.d_bytes = 0;
    for (int i = 0; i < kStageSlots; ++i) {
        if (g_host.stage[i]) cudaFreeHost(g_host.stage[i]);
        if (g_host.stage_ev[i]) cudaEventDestroy(g_host.stage_ev[i]);
        g_host.stage[i] = nullptr;
        g_host.stage_ev[i] = nullptr;
    }
    delete g_host.pool;
    g_host.pool = nullptr;
    if (g_host.stream) cudaStreamDestroy(g_host.stream);
    g_host.stream = nullptr;
    if (g_host.copy_stream) cudaStreamDestroy(g_host.copy_stream);
    g_host.copy_stream = nullptr;
    for (cudaEvent_t ev : g_host.events) cudaEventDestroy(ev);
    g_host.events.clear();
    if (g_host.h_small) cudaFreeHost(g_host.h_small);
    g_host.h_small = nullptr;
    for (cudaEvent_t ev : g_events) cudaEventDestroy(ev);
    g_events.clear();
    g_event_tags.clear();
    g_events_used = 0;
    return 0;
}

int b200sort_histogram(const uint32_t *d_keys, uint64_t n, int shift, int bits, uint32_t *d_hist,
                       void *d_temp, size_t temp_bytes, void *stream) {
    if (bits < 1 || bits > kMaxRadixBits || shift < 0 || shift > 31) return fail(B200SORT_EINVAL, "shift/bits");
    if (!d_hist || (n > 0 && !d_keys)) return fail(B200SORT_EINVAL, "null buffer");
    if (n > 0xFFFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    int rc = check_device();
    if (rc) return rc;
    const int bins = 1 << bits;
    const size_t need = 256 + (size_t)2 * bins * 4;
    if (!d_temp || ((uintptr_t)d_temp & 255u) || temp_bytes < need) return fail(B200SORT_ETEMP, "temp storage");
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_temp, 0, 256, s));
    CU(cudaMemsetAsync(d_hist, 0, (size_t)bins * 4, s));
    if (n == 0) return 0;
    HistArgs h{};
    h.keys = d_keys;
    h.n = n;
    h.ghist = d_hist;
    h.done = static_cast<uint32_t *>(d_temp);
    h.bin_base = reinterpret_cast<uint32_t *>(static_cast<char *>(d_temp) + 256);
    h.zero_ptr = nullptr;
    h.zero_vecs = 0;
    h.passes.count = 1;
    h.passes.width = bits;
    h.passes.shift[0] = (uint8_t)shift;
    h.passes.bits[0] = (uint8_t)std::min(bits, 32 - shift);
    CU(launch_hist(bits, false, h, hist_grid(n, 1, bits), s));
    return 0;
}

int b200sort_digit_pass(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                        uint32_t *d_keys_out, uint32_t *d_vals_out, int shift, int bits,
                        const uint64_t *d_bin_dst, void *d_temp, size_t temp_bytes, void *stream) {
    if (bits < 1 || bits > kMaxRadixBits || shift < 0 || shift > 31) return fail(B200SORT_EINVAL, "shift/bits");
    const bool pairs = d_vals_in != nullptr;
    const bool dst = d_bin_dst != nullptr;
    if (n > 0xFFFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    if (n == 0) return 0;
    if (!d_keys_in || (!dst && !d_keys_out) || (pairs && !dst && !d_vals_out))
        return fail(B200SORT_EINVAL, "null buffer");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;

    int variant = effective_variant(bits, pairs);
    // keys with per-bin destinations (the fused partition + exchange): the column-sweep kernel writes every
    // (tile, bin) run with one shared->global bulk copy (UBLKCP) instead of 4-byte stores
    if (dst && !pairs && g_params.dst_bulk) variant = kColVariant;
    if (dst && !variant_has_dst(bits, variant)) variant = fallback_variant(bits);
    if (variant_mode(variant) == 1 && (g_params.safe_rank || !run_selftest())) variant = kColVariant;
    const int tile = tile_keys(variant, pairs);
    const Layout L = make_layout(n, 1, bits, pairs, false, tile, g_params.portion_tiles);
    if (!d_temp || ((uintptr_t)d_temp & 255u) || temp_bytes < L.total) return fail(B200SORT_ETEMP, "temp storage");
    char *base = static_cast<char *>(d_temp);
    uint32_t *bin_base = reinterpret_cast<uint32_t *>(base + L.off_bin_base);
    uint32_t *desc = reinterpret_cast<uint32_t *>(base + L.off_desc);

    CU(cudaMemsetAsync(base, 0, L.header_bytes, s));
    if (!dst) {
        HistArgs h{};
        h.keys = d_keys_in;
        h.n = n;
        h.ghist = reinterpret_cast<uint32_t *>(base + L.off_ghist);
        h.bin_base = bin_base;
        h.done = reinterpret_cast<uint32_t *>(base + L.off_done);
        h.zero_ptr = reinterpret_cast<uint4 *>(desc);
        h.zero_vecs = L.desc_bytes / 16;
        h.passes.count = 1;
        h.passes.width = bits;
        h.passes.shift[0] = (uint8_t)shift;
        h.passes.bits[0] = (uint8_t)std::min(bits, 32 - shift);
        CU(launch_hist(bits, false, h, hist_grid(n, 1, bits), s));
    } else {
        // Destinations are per-bin arrays: offsets inside a bin start at zero.
        CU(cudaMemsetAsync(bin_base, 0, (size_t)2 * L.bins * 4, s));
        CU(cudaMemsetAsync(desc, 0, L.desc_bytes, s));
    }
    const uint64_t portion_keys = L.portion_tiles * (uint64_t)tile;
    for (uint64_t q = 0; q < L.portions; ++q) {
        const uint64_t first = q * portion_keys;
        const uint64_t count = std::min<uint64_t>(portion_keys, n - first);
        PassArgs a{};
        a.keys_in = d_keys_in + first;
        a.vals_in = pairs ? d_vals_in + first : nullptr;
        a.keys_out = d_keys_out;
        a.vals_out = d_vals_out;
        a.bin_base = bin_base + (q & 1) * L.bins;
        a.carry_out = (q + 1 < L.portions) ? bin_base + ((q + 1) & 1) * L.bins : nullptr;
        a.desc = desc + (size_t)q * L.portion_tiles * L.bins;
        a.ticket = reinterpret_cast<uint32_t *>(base + L.off_tickets) + q;
        a.bin_dst = d_bin_dst;
        a.n = (uint32_t)count;
        a.num_tiles = (uint32_t)((count + tile - 1) / tile);
        a.shift = (uint32_t)shift;
        a.mask = (1u << std::min(bits, 32 - shift)) - 1u;
        a.parity = 0;
        a.prefetch = prefetch_distance();
        CU(launch_pass(bits, variant, pairs, dst, a, s));
    }
    return 0;
}

size_t b200sort_scan_temp_bytes(uint64_t n) {
    return align_up(((n + kScanMinTile - 1) / kScanMinTile + 1) * sizeof(uint64_t), 256) + 256;
}

int b200sort_exclusive_scan(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp, size_t temp_bytes,
                            void *stream) {
    if (n == 0) return 0;
    if (!d_in || !d_out) return fail(B200SORT_EINVAL, "null buffer");
    if (((uintptr_t)d_in | (uintptr_t)d_out) & 15u) return fail(B200SORT_EINVAL, "scan buffers must be 16-byte aligned");
    const int sv = g_params.scan_variant;
    const uint64_t tile_elems = (uint64_t)scan_tile(sv);
    const uint64_t tiles = (n + tile_elems - 1) / tile_elems;
    if (tiles > 0x7FFFFFFFull) return fail(B200SORT_ETOOBIG, "n");
    if (!d_temp || ((uintptr_t)d_temp & 255u) || temp_bytes < b200sort_scan_temp_bytes(n))
        return fail(B200SORT_ETEMP, "temp storage");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_temp, 0, (tiles + 1) * sizeof(uint64_t), s));  // descriptors + the ticket of the persistent form
    g_launches.fetch_add(1, std::memory_order_relaxed);
    uint64_t *desc = static_cast<uint64_t *>(d_temp);
    const int pfp = g_params.scan_prefetch_tiles.load();
    const uint32_t pf = pfp >= 0 ? (uint32_t)pfp : (uint32_t)((8u << 20) / (tile_elems * 4u));
    switch (sv) {
    case 0: exclusive_scan_kernel<256, 4><<<(unsigned)tiles, 256, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 1: exclusive_scan_kernel<512, 4><<<(unsigned)tiles, 512, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 2: exclusive_scan_kernel<512, 8><<<(unsigned)tiles, 512, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 3: exclusive_scan_kernel<1024, 4><<<(unsigned)tiles, 1024, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 4: exclusive_scan_kernel<1024, 8><<<(unsigned)tiles, 1024, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 5: exclusive_scan_kernel<128, 8><<<(unsigned)tiles, 128, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 6: exclusive_scan_kernel<256, 8><<<(unsigned)tiles, 256, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 7: exclusive_scan_kernel<256, 16><<<(unsigned)tiles, 256, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 8: exclusive_scan_kernel<128, 16><<<(unsigned)tiles, 128, 0, s>>>(d_in, d_out, n, desc, pf); break;
    case 9: CU((launch_scan_persistent<1024, 4>(d_in, d_out, n, desc, (uint32_t)tiles, s))); break;
    case 10: CU((launch_scan_persistent<512, 8>(d_in, d_out, n, desc, (uint32_t)tiles, s))); break;
    case 11: CU((launch_scan_persistent<512, 4>(d_in, d_out, n, desc, (uint32_t)tiles, s))); break;
    default: CU((launch_scan_persistent<256, 8>(d_in, d_out, n, desc, (uint32_t)tiles, s))); break;
    }
    CU(cudaGetLastError());
    return 0;
}

int b200sort_generate(uint32_t *d_out, uint64_t first, uint64_t count, int kind, uint64_t total,
                      const uint32_t *d_zipf_cdf, void *stream) {
    if (kind < GEN_UNIFORM || kind > GEN_IOTA) return fail(B200SORT_EINVAL, "generator kind");
    if (kind == GEN_ZIPF && !d_zipf_cdf) return fail(B200SORT_EINVAL, "zipf needs the cdf table");
    if (count == 0) return 0;
    if (!d_out) return fail(B200SORT_EINVAL, "null buffer");
    int rc = check_device();
    if (rc) return rc;
    const int grid = (int)std::min<uint64_t>((count + 255) / 256, (uint64_t)num_sms() * 16);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    generate_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_out, first, count, kind, total, d_zipf_cdf);
    CU(cudaGetLastError());
    return 0;
}

int b200sort_route(const uint32_t *d_keys, uint64_t n, const uint64_t *d_thresholds, int count, uint32_t *d_route,
                   uint32_t *d_counts, void *stream) {
    if (count < 0 || count > kMaxRouteThresholds) return fail(B200SORT_EINVAL, "threshold count");
    if (n > 0 && (!d_keys || !d_route || (count > 0 && !d_thresholds))) return fail(B200SORT_EINVAL, "null buffer");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (d_counts) CU(cudaMemsetAsync(d_counts, 0, (size_t)(count + 1) * sizeof(uint32_t), s));
    if (n == 0) return 0;
    const uint64_t want = (n + 256ull * 8 - 1) / (256ull * 8);
    const int grid = (int)std::min<uint64_t>(std::max<uint64_t>(want, 1), (uint64_t)num_sms() * 8);
    if (count <= kRouteRegCuts)
        route_kernel<true><<<grid, 256, 0, s>>>(d_keys, n, d_thresholds, count, d_route, d_counts);
    else
        route_kernel<false><<<grid, 256, 0, s>>>(d_keys, n, d_thresholds, count, d_route, d_counts);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    CU(cudaGetLastError());
    return 0;
}

int b200sort_store_probe(uint32_t *d_dst, const uint32_t *d_src, uint64_t n, int vec, int ctas_per_sm, void *stream) {
    if (!d_dst || !d_src || (vec != 1 && vec != 4) || ctas_per_sm < 1) return fail(B200SORT_EINVAL, "store probe arguments");
    int rc = check_device();
    if (rc) return rc;
    g_launches.fetch_add(1, std::memory_order_relaxed);
    store_probe_kernel<<<num_sms() * ctas_per_sm, 256, 0, (cudaStream_t)stream>>>(d_dst, d_src, n, vec);
    CU(cudaGetLastError());
    return 0;
}

int b200sort_verify(const uint32_t *d_keys, uint64_t n, uint64_t *d_result, void *stream) {
    if (!d_result || (n > 0 && !d_keys)) return fail(B200SORT_EINVAL, "null buffer");
    int rc = check_device();
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemsetAsync(d_result, 0, 4 * sizeof(uint64_t), s));
    if (n == 0) return 0;
    const int grid = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)num_sms() * 16);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    verify_kernel<<<grid, 256, 0, s>>>(d_keys, n, reinterpret_cast<unsigned long long *>(d_result));
    CU(cudaGetLastError());
    return 0;
}

int b200sort_device_banner(char *buf, size_t len) {
    if (!buf || len == 0) return B200SORT_EINVAL;
    int rc = check_device();
    if (rc) return rc;
    int dev = 0;
    cudaDeviceProp p;
    CU(cudaGetDevice(&dev));
    CU(cudaGetDeviceProperties(&p, dev));
    // same fields and layout as printDeviceInfo(), SourceCode/Parallel7.cu:664-677
    snprintf(buf, len,
             "**********GPU info**********\n"
             "Name: %s\n"
             "Compute capability: %d.%d\n"
             "Num SMs: %d\n"
             "Max num threads per SM: %d\n"
             "Max num warps per SM: %d\n"
             "GMEM: %zu byte\n"
             "SMEM per SM: %zu byte\n"
             "SMEM per block: %zu byte\n"
             "****************************\n",
             p.name, p.major, p.minor, p.multiProcessorCount, p.maxThreadsPerMultiProcessor,
             p.maxThreadsPerMultiProcessor / p.warpSize, p.totalGlobalMem, p.sharedMemPerMultiprocessor,
             p.sharedMemPerBlock);
    return 0;
}

int b200sort_profile_enable(int on) {
    g_profile = on != 0;
    g_events_used = 0;
    return 0;
}

int b200sort_profile_read(float *ms, int *tag, int capacity) {
    if (!ms || !tag || capacity < 0) return B200SORT_EINVAL;
    int out = 0;
    if (g_events_used >= 2) {
        CU(cudaEventSynchronize(g_events[g_events_used - 1]));
        for (int i = 1; i < g_events_used && out < capacity; ++i) {
            if (g_event_tags[i] < 0) continue;  // a new sort starts here
            CU(cudaEventElapsedTime(&ms[out], g_events[i - 1], g_events[i]));
            tag[out] = g_event_tags[i];
            ++out;
        }
    }
    g_events_used = 0;
    return out;
}

}  // extern "C"
