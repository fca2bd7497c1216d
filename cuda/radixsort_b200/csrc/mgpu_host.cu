// mgpu_host.cu -- b200sort_mgpu_*_host: one host thread drives G GPUs of one node through the
// library's own C ABI (b200sort_histogram, b200sort_digit_pass with per-bin destinations,
// b200sort_keys / b200sort_pairs).  No kernels of its own.
//
// Same algorithm as the one-process-per-GPU driver in cuda/radixsort_b200/mgpu.py (SURVEY.md
// section 8e): the reference's sortByDevice (SourceCode/Parallel7.cu:530-639) has no multi-GPU
// form; this keeps its host-array contract and shards the work.
//
//   shard g -> its GPU (every GPU has its own PCIe link)
//   histogram of the partition byte on every GPU -> count matrix [src][bin] on the host
//   owner[bin]: bin edges closest to the ideal cuts j*n/G  (stable range partition)
//   one digit pass per GPU writing bin b straight into owner[b]'s receive buffer (peer stores
//   over NVLink); ranges are laid out source-major so equal keys stay in global index order
//   event-wait for all partitions -> local LSD sort of the received range -> D2H at its offset
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/b200sort.h"
#include "copy_pool.h"
#include "internal.h"

namespace b200sort {
namespace {

constexpr size_t kChunk = 16u << 20;  // bytes per pinned staging buffer
constexpr int kSlots = 2;
constexpr int kMaxDevices = 64;
constexpr int kPartBits = 8;
constexpr int kPartBins = 1 << kPartBits;

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e_ = (call);                                        \
        if (e_ != cudaSuccess) return set_cuda_error(e_, #call);        \
    } while (0)
#define RC(call)                \
    do {                        \
        int rc_ = (call);       \
        if (rc_ != 0) return rc_; \
    } while (0)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Buf {
    void *p = nullptr;
    size_t bytes = 0;
    int ensure(size_t want) {
        if (want <= bytes) return 0;
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return set_error(B200SORT_ENOMEM, "cudaMalloc of a multi-GPU shard buffer");
        }
        bytes = want;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

enum { EV_START = 0, EV_UPLOADED, EV_HIST, EV_PLANNED, EV_PARTITIONED, EV_EXCHANGED, EV_SORTED, EV_DOWNLOADED, EV_COUNT };

struct Shard {
    int dev = -1;
    cudaStream_t stream = nullptr;
    Buf in_k, in_v, recv_k, recv_v, out_k, out_v, temp, small, route, dump;
    void *stage[kSlots] = {nullptr, nullptr};
    cudaEvent_t stage_ev[kSlots] = {nullptr, nullptr};
    cudaEvent_t ev[EV_COUNT] = {};
    uint32_t *h_counts = nullptr;   // pinned [kPartBins]
    uint64_t *h_bin_dst = nullptr;  // pinned [4 * kPartBins]
    // per call
    uint64_t first = 0, count = 0;  // shard of the input
    uint64_t recv = 0, out_first = 0;
    const uint32_t *range_k = nullptr, *range_v = nullptr;  // what the local sort reads
    uint32_t *sorted_k = nullptr, *sorted_v = nullptr;
};

struct State {
    std::mutex mu;
    std::vector<Shard> shards;
    CopyPool *pool = nullptr;
    double stats[B200SORT_MGPU_STATS] = {};
    int stats_valid = 0;
} g;

void release_shard(Shard &s) {
    if (s.dev < 0) return;
    cudaSetDevice(s.dev);
    if (s.stream) cudaStreamSynchronize(s.stream);
    s.in_k.release(); s.in_v.release(); s.recv_k.release(); s.recv_v.release();
    s.out_k.release(); s.out_v.release(); s.temp.release(); s.small.release();
    s.route.release(); s.dump.release();
    for (int i = 0; i < kSlots; ++i) {
        if (s.stage[i]) cudaFreeHost(s.stage[i]);
        if (s.stage_ev[i]) cudaEventDestroy(s.stage_ev[i]);
        s.stage[i] = nullptr;
        s.stage_ev[i] = nullptr;
    }
    for (auto &e : s.ev) {
        if (e) cudaEventDestroy(e);
        e = nullptr;
    }
    if (s.h_counts) cudaFreeHost(s.h_counts);
    if (s.h_bin_dst) cudaFreeHost(s.h_bin_dst);
    s.h_counts = nullptr;
    s.h_bin_dst = nullptr;
    if (s.stream) cudaStreamDestroy(s.stream);
    s.stream = nullptr;
    s.dev = -1;
    cudaGetLastError();
}

int prepare_shard(Shard &s, int dev) {
    if (s.dev != dev) release_shard(s);
    CU(cudaSetDevice(dev));
    s.dev = dev;
    if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    for (auto &e : s.ev)
        if (!e) CU(cudaEventCreate(&e));
    if (!s.h_counts) CU(cudaMallocHost(reinterpret_cast<void **>(&s.h_counts), kPartBins * sizeof(uint32_t)));
    if (!s.h_bin_dst) CU(cudaMallocHost(reinterpret_cast<void **>(&s.h_bin_dst), 4 * kPartBins * sizeof(uint64_t)));
    RC(s.small.ensure(8192));  // [hist: kPartBins u32 | pad | bin_dst: 2*kPartBins u64 (4*bins by value) | thresholds @6144]
    return 0;
}

int ensure_staging(Shard &s) {
    CU(cudaSetDevice(s.dev));
    for (int i = 0; i < kSlots; ++i) {
        if (!s.stage[i]) {
            cudaError_t e = cudaMallocHost(&s.stage[i], kChunk);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return set_error(B200SORT_ENOMEM, "cudaMallocHost of pinned staging buffers");
            }
        }
        if (!s.stage_ev[i]) CU(cudaEventCreateWithFlags(&s.stage_ev[i], cudaEventDisableTiming));
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

// Host array -> the `in` buffer of every shard.  Pinned arrays: one asynchronous copy per GPU,
// all links busy at once.  Pageable arrays: chunks go through each shard's pinned ring, the
// shards taking turns so that every link always has a chunk in flight.
int upload_all(std::vector<Shard> &sh, const uint32_t *h, bool vals) {
    const bool staged = is_pageable(h);
    if (!staged) {
        for (auto &s : sh) {
            if (!s.count) continue;
            CU(cudaSetDevice(s.dev));
            CU(cudaMemcpyAsync(vals ? s.in_v.p : s.in_k.p, h + s.first, s.count * 4, cudaMemcpyHostToDevice, s.stream));
        }
        return 0;
    }
    if (!g.pool) g.pool = new CopyPool();
    size_t max_bytes = 0;
    for (auto &s : sh) {
        RC(ensure_staging(s));
        max_bytes = std::max<size_t>(max_bytes, s.count * 4);
    }
    int c = 0;
    for (size_t off = 0; off < max_bytes; off += kChunk, ++c) {
        const int slot = c % kSlots;
        for (auto &s : sh) {
            const size_t bytes = s.count * 4;
            if (off >= bytes) continue;
            const size_t len = std::min(kChunk, bytes - off);
            CU(cudaSetDevice(s.dev));
            CU(cudaEventSynchronize(s.stage_ev[slot]));  // the copy that last used this slot is done
            g.pool->copy(s.stage[slot], reinterpret_cast<const char *>(h + s.first) + off, len);
            CU(cudaMemcpyAsync(static_cast<char *>(vals ? s.in_v.p : s.in_k.p) + off, s.stage[slot], len,
                               cudaMemcpyHostToDevice, s.stream));
            CU(cudaEventRecord(s.stage_ev[slot], s.stream));
        }
    }
    return 0;
}

// Sorted ranges -> host array.  Asynchronous for pinned destinations; staged ones return with
// the data in place.
int download_all(std::vector<Shard> &sh, uint32_t *h, bool vals) {
    const bool staged = is_pageable(h);
    if (!staged) {
        for (auto &s : sh) {
            if (!s.recv) continue;
            CU(cudaSetDevice(s.dev));
            CU(cudaMemcpyAsync(h + s.out_first, vals ? s.sorted_v : s.sorted_k, s.recv * 4, cudaMemcpyDeviceToHost, s.stream));
        }
        return 0;
    }
    if (!g.pool) g.pool = new CopyPool();
    size_t max_bytes = 0;
    for (auto &s : sh) {
        RC(ensure_staging(s));
        max_bytes = std::max<size_t>(max_bytes, s.recv * 4);
    }
    auto issue = [&](Shard &s, size_t c) -> int {
        const size_t bytes = s.recv * 4, off = c * kChunk;
        if (off >= bytes) return 0;
        const size_t len = std::min(kChunk, bytes - off);
        const int slot = (int)(c % kSlots);
        CU(cudaSetDevice(s.dev));
        CU(cudaMemcpyAsync(s.stage[slot], reinterpret_cast<const char *>(vals ? s.sorted_v : s.sorted_k) + off, len,
                           cudaMemcpyDeviceToHost, s.stream));
        CU(cudaEventRecord(s.stage_ev[slot], s.stream));
        return 0;
    };
    const size_t chunks = (max_bytes + kChunk - 1) / kChunk;
    for (size_t c = 0; c < std::min<size_t>(chunks, kSlots); ++c)
        for (auto &s : sh) RC(issue(s, c));
    for (size_t c = 0; c < chunks; ++c) {
        for (auto &s : sh) {
            const size_t bytes = s.recv * 4, off = c * kChunk;
            if (off >= bytes) continue;
            const size_t len = std::min(kChunk, bytes - off);
            CU(cudaSetDevice(s.dev));
            CU(cudaEventSynchronize(s.stage_ev[c % kSlots]));
            g.pool->copy(reinterpret_cast<char *>(h + s.out_first) + off, s.stage[c % kSlots], len);
            RC(issue(s, c + kSlots));  // the slot just drained takes the chunk after next
        }
    }
    return 0;
}

// owner[b] for the bins of the partition digit: cut j sits on the bin edge closest to j*total/G.
void choose_owner(const uint64_t *hist, int bins, int G, int *owner) {
    std::vector<uint64_t> csum(bins + 1, 0);
    for (int b = 0; b < bins; ++b) csum[b + 1] = csum[b] + hist[b];
    const uint64_t total = csum[bins];
    std::vector<int> cuts(G + 1, 0);
    for (int j = 1; j < G; ++j) {
        const unsigned __int128 target = (unsigned __int128)total * j;  // compared against csum * G
        int b = 0;
        while (b < bins && (unsigned __int128)csum[b] * G < target) ++b;
        if (b > 0) {
            const unsigned __int128 hi = (unsigned __int128)csum[b] * G, lo = (unsigned __int128)csum[b - 1] * G;
            const unsigned __int128 d_hi = hi > target ? hi - target : target - hi;
            const unsigned __int128 d_lo = lo > target ? lo - target : target - lo;
            if (d_lo <= d_hi) --b;
        }
        cuts[j] = std::max(b, cuts[j - 1]);
    }
    cuts[G] = bins;
    for (int r = 0; r < G; ++r)
        for (int b = cuts[r]; b < cuts[r + 1]; ++b) owner[b] = r;
}

int sync_all(std::vector<Shard> &sh) {
    int rc = 0;
    for (auto &s : sh) {
        if (s.dev < 0 || !s.stream) continue;
        cudaSetDevice(s.dev);
        cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess && rc == 0) rc = set_cuda_error(e, "cudaStreamSynchronize");
    }
    return rc;
}

int mark(Shard &s, int which) {
    CU(cudaSetDevice(s.dev));
    CU(cudaEventRecord(s.ev[which], s.stream));
    return 0;
}

// Value splitters for skewed keys (the host-side twin of mgpu.py:value_splitters): cut j aims at
// the j/G quantile of the pooled sample and sits at a key value v_j.  Keys equal to v_j go left of
// the cut when they come from a shard < split[j], right from a shard > split[j], and inside shard
// split[j] those at local index < pos[j] go left.  Cutting a run of equal keys at a position of the
// global input order keeps ties in that order and lets one heavy value spread over shards.
constexpr uint64_t kTieAllLeft = 1ull << 62;
struct ValueCuts {
    std::vector<uint64_t> value, pos;
    std::vector<int> split;
};
struct ShardSample {
    std::vector<uint32_t> key;
    std::vector<uint64_t> at;  // local index each sample was taken from (ascending)
};
ValueCuts value_splitters(const std::vector<uint32_t> &pool_sorted, const std::vector<ShardSample> &by_shard, int G) {
    ValueCuts c;
    c.value.assign(G > 1 ? G - 1 : 0, 0);
    c.pos.assign(c.value.size(), 0);
    c.split.assign(c.value.size(), 0);
    const size_t m = pool_sorted.size();
    if (m == 0) return c;
    for (int j = 1; j < G; ++j) {
        const size_t q = std::min((size_t)j * m / G, m - 1);
        const uint32_t v = pool_sorted[q];
        const size_t lo = std::lower_bound(pool_sorted.begin(), pool_sorted.end(), v) - pool_sorted.begin();
        size_t left = q - lo;  // sampled copies of v that belong left of the cut
        c.value[j - 1] = v;
        if (left == 0) continue;  // the cut sits at the start of the run: split 0, pos 0 = all of it goes right
        c.split[j - 1] = G;       // default: the whole run goes left
        for (int r = 0; r < G && c.split[j - 1] == G; ++r) {
            const ShardSample &sm = by_shard[r];
            for (size_t i = 0; i < sm.key.size(); ++i) {
                if (sm.key[i] != v) continue;
                if (left == 0) {
                    c.split[j - 1] = r;
                    c.pos[j - 1] = sm.at[i];
                    break;
                }
                --left;
            }
        }
    }
    return c;
}
// Cuts of one source shard for b200sort_route: `count` values followed by `count` tie indices.
std::vector<uint64_t> thresholds_for_shard(const ValueCuts &c, int shard) {
    const size_t count = c.value.size();
    std::vector<uint64_t> t(2 * count);
    for (size_t j = 0; j < count; ++j) {
        t[j] = c.value[j];
        t[count + j] = shard < c.split[j] ? kTieAllLeft : (shard > c.split[j] ? 0 : c.pos[j]);
    }
    return t;
}

constexpr size_t kSamplePerShard = 8192;

// Histogram of the partition digit on every shard, splitters, then one digit pass per shard that
// writes each bin into its owner's receive buffer.  On return every stream has waited for all
// partitions and the shards know their received range.
int partition_and_exchange(std::vector<Shard> &sh, const uint32_t *hk_in, uint64_t n, int nbits, bool pairs,
                           int &part_shift_out, int &part_bits_out, uint64_t &max_recv_out, double &plan_ms_out,
                           bool &by_value_out) {
    const int G = (int)sh.size();
    // ---- partition digit: the highest byte in which the keys differ -----------------------------
    std::vector<uint64_t> counts((size_t)G * kPartBins);  // [src][bin], re-shaped below when the digit narrows
    std::vector<uint64_t> global(kPartBins);
    int shift = 32 - kPartBits;
    for (;;) {
        for (auto &s : sh) {
            CU(cudaSetDevice(s.dev));
            uint32_t *d_hist = static_cast<uint32_t *>(s.small.p);
            RC(b200sort_histogram(static_cast<const uint32_t *>(s.in_k.p), s.count, shift, kPartBits, d_hist,
                                  s.temp.p, s.temp.bytes, s.stream));
            CU(cudaMemcpyAsync(s.h_counts, d_hist, kPartBins * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
            CU(cudaEventRecord(s.ev[EV_HIST], s.stream));
        }
        RC(sync_all(sh));
        std::fill(global.begin(), global.end(), 0);
        for (int i = 0; i < G; ++i)
            for (int b = 0; b < kPartBins; ++b) {
                counts[(size_t)i * kPartBins + b] = sh[i].h_counts[b];
                global[b] += sh[i].h_counts[b];
            }
        const bool one_bin = *std::max_element(global.begin(), global.end()) == n;
        if (!one_bin || shift == 0) break;
        shift -= kPartBits;  // every key shares this byte: partition on the next one
    }

    // ---- plan ----------------------------------------------------------------------------------
    const auto plan_t0 = std::chrono::steady_clock::now();
    int bins = kPartBins, part_bits = kPartBits, part_shift = shift;
    std::vector<int> owner(kPartBins);
    choose_owner(global.data(), kPartBins, G, owner.data());
    std::vector<uint64_t> matrix((size_t)G * G, 0);  // [src][dst]
    auto fill_matrix = [&]() {
        std::fill(matrix.begin(), matrix.end(), 0);
        for (int i = 0; i < G; ++i)
            for (int b = 0; b < bins; ++b) matrix[(size_t)i * G + owner[b]] += counts[(size_t)i * bins + b];
        uint64_t mx = 0;
        for (int d = 0; d < G; ++d) {
            uint64_t tot = 0;
            for (int i = 0; i < G; ++i) tot += matrix[(size_t)i * G + d];
            mx = std::max(mx, tot);
        }
        return mx;
    };
    uint64_t max_recv = fill_matrix();

    const int permille = g_mgpu_balance_permille.load();
    const bool by_value = permille > 0 && (double)max_recv * G > (double)n * permille / 1000.0;
    if (by_value) {
        // Skewed keys: bin edges of one byte cannot balance the shards.  Splitters become key values
        // from a sample of the host array; b200sort_route turns every key into its destination, and
        // that route array is the KEY of the partition pass below, which carries the real keys.
        std::vector<ShardSample> by_shard(G);
        std::vector<uint32_t> pool;
        pool.reserve((size_t)G * kSamplePerShard);
        for (int g2 = 0; g2 < G; ++g2) {
            const Shard &s = sh[g2];
            if (!s.count) continue;
            const uint64_t stride = std::max<uint64_t>(s.count / kSamplePerShard, 1);
            for (uint64_t i = 0; i < kSamplePerShard; ++i) {
                const uint64_t base = (uint64_t)(((unsigned __int128)i * s.count) / kSamplePerShard);
                const uint64_t at = std::min<uint64_t>(base + ((i * 2654435761ull) & 0xFFFFFFFFull) % stride, s.count - 1);
                by_shard[g2].key.push_back(hk_in[s.first + at]);
                by_shard[g2].at.push_back(at);
            }
            pool.insert(pool.end(), by_shard[g2].key.begin(), by_shard[g2].key.end());
        }
        std::sort(pool.begin(), pool.end());  // <= 64 x 8192 sample keys: planning, not the sort
        const ValueCuts cuts = value_splitters(pool, by_shard, G);
        part_bits = 1;
        while ((1 << part_bits) < G) ++part_bits;
        part_shift = 0;
        bins = 1 << part_bits;
        for (int g2 = 0; g2 < G; ++g2) {
            Shard &s = sh[g2];
            const std::vector<uint64_t> thresholds = thresholds_for_shard(cuts, g2);
            CU(cudaSetDevice(s.dev));
            const size_t bytes = align_up(std::max<uint64_t>(s.count, 1) * 4, 256);
            RC(s.route.ensure(bytes));
            RC(s.dump.ensure(bytes));
            uint64_t *d_thr = reinterpret_cast<uint64_t *>(static_cast<char *>(s.small.p) + 6144);
            std::copy(thresholds.begin(), thresholds.end(), s.h_bin_dst);  // pinned scratch, re-filled below
            CU(cudaMemcpyAsync(d_thr, s.h_bin_dst, thresholds.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, s.stream));
            uint32_t *d_hist = static_cast<uint32_t *>(s.small.p);  // G destination counts, tallied by the route kernel
            RC(b200sort_route(static_cast<const uint32_t *>(s.in_k.p), s.count, d_thr, (int)(thresholds.size() / 2),
                              static_cast<uint32_t *>(s.route.p), d_hist, s.stream));
            CU(cudaMemcpyAsync(s.h_counts, d_hist, G * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.stream));
        }
        RC(sync_all(sh));
        counts.assign((size_t)G * bins, 0);
        for (int i = 0; i < G; ++i)
            for (int b = 0; b < G; ++b) counts[(size_t)i * bins + b] = sh[i].h_counts[b];
        owner.assign(bins, 0);
        for (int b = 0; b < bins; ++b) owner[b] = std::min(b, G - 1);
        max_recv = fill_matrix();
    } else {
        // Splitters on the top log2(G) bits (uniform keys, G a power of two): partition with a
        // log2(G)-bit digit -- G bins instead of 256, long runs per (tile, destination).
        int lg = 0;
        while ((1 << (lg + 1)) <= G) ++lg;
        bool narrow = G > 1 && (1 << lg) == G && lg <= kPartBits;
        for (int b = 0; narrow && b < kPartBins; ++b) narrow = owner[b] == (b >> (kPartBits - lg));
        if (narrow) {
            part_bits = lg;
            part_shift = shift + kPartBits - lg;
            bins = G;
            counts = matrix;  // [src][dst] is the count table of the narrow digit
            owner.resize(G);
            for (int b = 0; b < G; ++b) owner[b] = b;
        }
    }

    uint64_t running_out = 0;
    for (int d = 0; d < G; ++d) {
        uint64_t tot = 0;
        for (int i = 0; i < G; ++i) tot += matrix[(size_t)i * G + d];
        if (tot > 0xFFFFFFFFull) return set_error(B200SORT_ETOOBIG, "received range");
        sh[d].recv = tot;
        sh[d].out_first = running_out;
        running_out += tot;
    }
    // The plan proper ends here (O(G * bins) host work).  Growing the cached receive / output / temp buffers
    // below is cudaMalloc time, paid by the first call of a size only: it is not part of plan_ms.
    const auto plan_t1 = std::chrono::steady_clock::now();
    for (int d = 0; d < G; ++d) {
        Shard &s = sh[d];
        CU(cudaSetDevice(s.dev));
        const size_t bytes = align_up(std::max<uint64_t>(s.recv, 1) * 4, 256);
        RC(s.recv_k.ensure(bytes));
        if (pairs) RC(s.recv_v.ensure(bytes));
        s.range_k = static_cast<const uint32_t *>(s.recv_k.p);
        s.range_v = static_cast<const uint32_t *>(s.recv_v.p);
        // the shard's input buffers are free once every partition is done (Buf::ensure only grows, and in_v
        // is only sized by pairs calls: check it on its own)
        if (bytes <= s.in_k.bytes && (!pairs || bytes <= s.in_v.bytes)) {
            s.sorted_k = static_cast<uint32_t *>(s.in_k.p);
            s.sorted_v = static_cast<uint32_t *>(s.in_v.p);
        } else {
            RC(s.out_k.ensure(bytes));
            if (pairs) RC(s.out_v.ensure(bytes));
            s.sorted_k = static_cast<uint32_t *>(s.out_k.p);
            s.sorted_v = static_cast<uint32_t *>(s.out_v.p);
        }
        size_t t = b200sort_temp_bytes(s.recv, nbits, pairs);
        if (by_value) t = std::max(t, b200sort_temp_bytes(s.count, part_bits, true));
        if (t > s.temp.bytes) RC(s.temp.ensure(align_up(t, 256)));
    }

    // ---- partition fused with the exchange -------------------------------------------------------
    std::vector<uint64_t> src_base(G, 0);  // keys of lower source ranks already placed in each owner's range
    for (int i = 0; i < G; ++i) {
        Shard &s = sh[i];
        std::vector<uint64_t> at(src_base);
        uint64_t local = 0;
        for (int b = 0; b < bins; ++b) {
            const int o = owner[b];
            const uint64_t cnt = counts[(size_t)i * bins + b];
            const uint64_t to_k = reinterpret_cast<uint64_t>(static_cast<uint32_t *>(sh[o].recv_k.p) + at[o]);
            const uint64_t to_v = pairs ? reinterpret_cast<uint64_t>(static_cast<uint32_t *>(sh[o].recv_v.p) + at[o]) : 0;
            if (by_value) {  // keys = route (stays here, in the dump buffer), carried = real keys; values in a second pass
                s.h_bin_dst[b] = reinterpret_cast<uint64_t>(static_cast<uint32_t *>(s.dump.p) + local);
                s.h_bin_dst[bins + b] = to_k;
                s.h_bin_dst[2 * bins + b] = s.h_bin_dst[b];
                s.h_bin_dst[3 * bins + b] = to_v;
            } else {
                s.h_bin_dst[b] = to_k;
                s.h_bin_dst[bins + b] = to_v;
            }
            at[o] += cnt;
            local += cnt;
        }
        for (int d = 0; d < G; ++d) src_base[d] += matrix[(size_t)i * G + d];
        CU(cudaSetDevice(s.dev));
        uint64_t *d_bin_dst = reinterpret_cast<uint64_t *>(static_cast<char *>(s.small.p) + 2048);
        CU(cudaMemcpyAsync(d_bin_dst, s.h_bin_dst, (size_t)(by_value ? 4 : 2) * bins * sizeof(uint64_t),
                           cudaMemcpyHostToDevice, s.stream));
        CU(cudaEventRecord(s.ev[EV_PLANNED], s.stream));
        if (s.count && !by_value)
            RC(b200sort_digit_pass(static_cast<const uint32_t *>(s.in_k.p),
                                   pairs ? static_cast<const uint32_t *>(s.in_v.p) : nullptr, s.count, nullptr, nullptr,
                                   part_shift, part_bits, d_bin_dst, s.temp.p, s.temp.bytes, s.stream));
        if (s.count && by_value) {
            RC(b200sort_digit_pass(static_cast<const uint32_t *>(s.route.p), static_cast<const uint32_t *>(s.in_k.p),
                                   s.count, nullptr, nullptr, 0, part_bits, d_bin_dst, s.temp.p, s.temp.bytes, s.stream));
            if (pairs)
                RC(b200sort_digit_pass(static_cast<const uint32_t *>(s.route.p), static_cast<const uint32_t *>(s.in_v.p),
                                       s.count, nullptr, nullptr, 0, part_bits, d_bin_dst + 2 * bins, s.temp.p,
                                       s.temp.bytes, s.stream));
        }
        CU(cudaEventRecord(s.ev[EV_PARTITIONED], s.stream));
    }
    for (int i = 0; i < G; ++i) {
        CU(cudaSetDevice(sh[i].dev));
        for (int j = 0; j < G; ++j)
            if (j != i) CU(cudaStreamWaitEvent(sh[i].stream, sh[j].ev[EV_PARTITIONED], 0));
        CU(cudaEventRecord(sh[i].ev[EV_EXCHANGED], sh[i].stream));
    }

    plan_ms_out = std::chrono::duration<double, std::milli>(plan_t1 - plan_t0).count();
    part_shift_out = part_shift;
    part_bits_out = part_bits;
    max_recv_out = max_recv;
    by_value_out = by_value;
    return 0;
}

// One shard: nothing to partition, the uploaded array is sorted as it is.
int single_shard(Shard &s, int nbits, bool pairs, uint64_t &max_recv_out) {
    CU(cudaSetDevice(s.dev));
    const size_t bytes = align_up(std::max<uint64_t>(s.count, 1) * 4, 256);
    RC(s.out_k.ensure(bytes));
    if (pairs) RC(s.out_v.ensure(bytes));
    const size_t t = b200sort_temp_bytes(s.count, nbits, pairs);
    if (t > s.temp.bytes) RC(s.temp.ensure(align_up(t, 256)));
    // the "received range" is the input buffer itself
    s.range_k = static_cast<const uint32_t *>(s.in_k.p);
    s.range_v = static_cast<const uint32_t *>(s.in_v.p);
    s.recv = s.count;
    s.out_first = 0;
    s.sorted_k = static_cast<uint32_t *>(s.out_k.p);
    s.sorted_v = static_cast<uint32_t *>(s.out_v.p);
    for (int e : {EV_HIST, EV_PLANNED, EV_PARTITIONED, EV_EXCHANGED}) CU(cudaEventRecord(s.ev[e], s.stream));
    max_recv_out = s.count;
    return 0;
}

int sort_mgpu(const uint32_t *hk_in, const uint32_t *hv_in, uint64_t n, uint32_t *hk_out, uint32_t *hv_out,
              int nbits, int block_size, const int *devices, int num_devices, bool pairs) {
    if (nbits < 1 || nbits > 16) return set_error(B200SORT_EINVAL, "nBits must be in 1..16");
    if (block_size <= 0) return set_error(B200SORT_EINVAL, "blockSize must be positive");
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible <= 0) {
        cudaGetLastError();
        return set_error(B200SORT_ENODEVICE, "cudaGetDeviceCount");
    }
    const int G = num_devices > 0 ? num_devices : visible;
    if (G > kMaxDevices) return set_error(B200SORT_EINVAL, "too many devices");
    std::vector<int> devs(G);
    for (int i = 0; i < G; ++i) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= visible) return set_error(B200SORT_EINVAL, "device ordinal out of range");
    }
    if (n == 0) return B200SORT_OK;
    if (!hk_in || !hk_out || (pairs && (!hv_in || !hv_out))) return set_error(B200SORT_EINVAL, "null buffer");

    std::lock_guard<std::mutex> lock(g.mu);
    g.stats_valid = 0;
    int caller_dev = 0;
    cudaGetDevice(&caller_dev);
    struct Restore {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore{caller_dev};

    if ((int)g.shards.size() < G) g.shards.resize(G);
    for (int i = G; i < (int)g.shards.size(); ++i) release_shard(g.shards[i]);
    g.shards.resize(G);
    std::vector<Shard> &sh = g.shards;

    // ---- devices, peer access, shard buffers --------------------------------------------------
    for (int i = 0; i < G; ++i) {
        RC(prepare_shard(sh[i], devs[i]));
        if (b200sort_get_param("atomic_rank_ok") < 0)  // also validates the device (sm_100 only)
            return set_error(B200SORT_ENODEVICE, "device is not usable by libb200sort");
        for (int j = 0; j < G; ++j) {
            if (devs[j] == devs[i]) continue;
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, devs[i], devs[j]));
            if (!can) return set_error(B200SORT_ENOPEER, "cudaDeviceCanAccessPeer");
            cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) return set_cuda_error(e, "cudaDeviceEnablePeerAccess");
        }
    }
    size_t max_count = 0;
    for (int i = 0; i < G; ++i) {
        Shard &s = sh[i];
        s.first = (uint64_t)(((unsigned __int128)n * i) / G);
        s.count = (uint64_t)(((unsigned __int128)n * (i + 1)) / G) - s.first;
        if (s.count > 0xFFFFFFFFull) return set_error(B200SORT_ETOOBIG, "shard");
        max_count = std::max<size_t>(max_count, s.count);
        CU(cudaSetDevice(s.dev));
        // a little slack so that the buffer can usually take the sorted range afterwards
        const size_t cap = align_up((s.count + s.count / 16 + 1024) * 4, 256);
        RC(s.in_k.ensure(cap));
        if (pairs) RC(s.in_v.ensure(cap));
        const size_t t = std::max(b200sort_temp_bytes(s.count, kPartBits, pairs),
                                  (size_t)256 + 2 * kPartBins * 4);
        RC(s.temp.ensure(align_up(t, 256)));
    }

    // ---- upload --------------------------------------------------------------------------------
    for (auto &s : sh) RC(mark(s, EV_START));
    RC(upload_all(sh, hk_in, false));
    if (pairs) RC(upload_all(sh, hv_in, true));
    for (auto &s : sh) RC(mark(s, EV_UPLOADED));

    int part_shift = 0, part_bits = 0;
    uint64_t max_recv = 0;
    double plan_ms = 0.0;
    bool by_value = false;
    if (G > 1) RC(partition_and_exchange(sh, hk_in, n, nbits, pairs, part_shift, part_bits, max_recv, plan_ms, by_value));
    else RC(single_shard(sh[0], nbits, pairs, max_recv));

    // ---- local sorts, download -------------------------------------------------------------------
    for (auto &s : sh) {
        CU(cudaSetDevice(s.dev));
        if (s.recv) {
            if (pairs)
                RC(b200sort_pairs(s.range_k, s.range_v, s.recv,
                                  s.sorted_k, s.sorted_v, s.temp.p, s.temp.bytes, nbits, s.stream));
            else
                RC(b200sort_keys(s.range_k, s.recv, s.sorted_k, s.temp.p, s.temp.bytes,
                                 nbits, s.stream));
        }
        CU(cudaEventRecord(s.ev[EV_SORTED], s.stream));
    }
    RC(download_all(sh, hk_out, false));
    if (pairs) RC(download_all(sh, hv_out, true));
    for (auto &s : sh) RC(mark(s, EV_DOWNLOADED));
    RC(sync_all(sh));

    // ---- figures ---------------------------------------------------------------------------------
    double *st = g.stats;
    std::fill(st, st + B200SORT_MGPU_STATS, 0.0);
    for (auto &s : sh) {
        cudaSetDevice(s.dev);
        for (int k = 0; k + 1 < EV_COUNT; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s.ev[k], s.ev[k + 1]) == cudaSuccess) st[k] = std::max(st[k], (double)ms);
            else cudaGetLastError();
        }
    }
    // The histogram -> partition gap on a device is mostly the wait for the slowest upload of
    // the node (the splitters need every histogram); report the host's own planning time instead.
    st[2] = plan_ms;
    st[7] = part_shift;
    st[8] = part_bits;
    st[9] = (double)max_recv * G / (double)n;
    st[10] = G;
    st[11] = by_value ? 1.0 : 0.0;
    g.stats_valid = 1;
    return B200SORT_OK;
}

int guarded(const uint32_t *hk_in, const uint32_t *hv_in, uint64_t n, uint32_t *hk_out, uint32_t *hv_out, int nbits,
            int block_size, const int *devices, int num_devices, bool pairs) {
    const int rc = sort_mgpu(hk_in, hv_in, n, hk_out, hv_out, nbits, block_size, devices, num_devices, pairs);
    if (rc != 0) {  // leave no work in flight on buffers the caller is about to reuse
        const std::string first = b200sort_last_error_string();
        sync_all(g.shards);
        cudaGetLastError();
        set_error_message(first.c_str());
    }
    return rc;
}

}  // namespace
}  // namespace b200sort

using namespace b200sort;

extern "C" {

int b200sort_mgpu_keys_host(const uint32_t *h_in, uint64_t n, uint32_t *h_out, int nBits, int blockSize,
                            const int *devices, int num_devices) {
    return guarded(h_in, nullptr, n, h_out, nullptr, nBits, blockSize, devices, num_devices, false);
}

int b200sort_mgpu_pairs_host(const uint32_t *h_keys_in, const uint32_t *h_vals_in, uint64_t n, uint32_t *h_keys_out,
                             uint32_t *h_vals_out, int nBits, int blockSize, const int *devices, int num_devices) {
    return guarded(h_keys_in, h_vals_in, n, h_keys_out, h_vals_out, nBits, blockSize, devices, num_devices, true);
}

int b200sort_plan_owners(const uint64_t *hist, int bins, int num_shards, int *owner) {
    if (!hist || !owner || bins < 1 || num_shards < 1) return set_error(B200SORT_EINVAL, "plan_owners arguments");
    choose_owner(hist, bins, num_shards, owner);
    return 0;
}

int b200sort_plan_value_cuts(const uint32_t *sample_keys, const uint64_t *sample_pos, const uint64_t *shard_offsets,
                             int num_shards, uint64_t *values, int *split_shard, uint64_t *split_pos) {
    if (num_shards < 1 || !shard_offsets || !values || !split_shard || !split_pos)
        return set_error(B200SORT_EINVAL, "plan_value_cuts arguments");
    const uint64_t total = shard_offsets[num_shards];
    if (total && (!sample_keys || !sample_pos)) return set_error(B200SORT_EINVAL, "plan_value_cuts arguments");
    std::vector<ShardSample> by_shard(num_shards);
    std::vector<uint32_t> pool(sample_keys, sample_keys + total);
    for (int r = 0; r < num_shards; ++r) {
        if (shard_offsets[r + 1] < shard_offsets[r]) return set_error(B200SORT_EINVAL, "shard_offsets");
        by_shard[r].key.assign(sample_keys + shard_offsets[r], sample_keys + shard_offsets[r + 1]);
        by_shard[r].at.assign(sample_pos + shard_offsets[r], sample_pos + shard_offsets[r + 1]);
    }
    std::sort(pool.begin(), pool.end());
    const ValueCuts c = value_splitters(pool, by_shard, num_shards);
    for (int j = 0; j + 1 < num_shards; ++j) {
        values[j] = c.value[j];
        split_shard[j] = c.split[j];
        split_pos[j] = c.pos[j];
    }
    return 0;
}

int b200sort_mgpu_last_stats(double *out, int capacity) {
    std::lock_guard<std::mutex> lock(g.mu);
    if (!out || capacity <= 0 || !g.stats_valid) return 0;
    const int k = std::min<int>(capacity, B200SORT_MGPU_STATS);
    std::copy(g.stats, g.stats + k, out);
    return k;
}

int b200sort_mgpu_shutdown(void) {
    std::lock_guard<std::mutex> lock(g.mu);
    int dev = 0;
    const bool have = cudaGetDevice(&dev) == cudaSuccess;
    for (auto &s : g.shards) release_shard(s);
    g.shards.clear();
    delete g.pool;
    g.pool = nullptr;
    g.stats_valid = 0;
    if (have) cudaSetDevice(dev);
    cudaGetLastError();
    return 0;
}

}  // extern "C"
