/*
 * b200sort.h -- C ABI of libb200sort.so, a B200 (sm_100a) LSD radix sort for uint32 keys and
 * stable uint32 key/value pairs.
 *
 * This is the drop-in boundary for the device sort path of truongchauhien/CUDA.RadixSort.
 * The reference has no FFI layer: its boundary is a set of free C++ functions in one
 * translation unit.  Each entry point below names the reference
 * interface it replaces (paths relative to the reference checkout).
 *
 * Conventions
 *   - every function returns 0 on success or a negative B200SORT_E* code / positive
 *     cudaError_t; the library never prints and never calls exit() (the reference's CHECK
 *     macro, SourceCode/common/common.h:6-16, prints and exits -- the C++ shim
 *     include/radix_sort_compat.hpp reproduces that on top of these return codes);
 *   - there is NO CPU path: every sort runs the hand-written sm_100a kernels or fails;
 *   - n is 64-bit in the ABI (the reference's `int n` caps it at 2^31-1); one call sorts
 *     at most 2^32-1 keys on one GPU;
 *   - nBits is the reference's digit width (SourceCode/Parallel7.cu:644): 1..16.  Widths
 *     1..8 run one kernel pass per digit; widths 9..16 run each digit as two stable
 *     sub-digit passes (the sorted output of an LSD sort does not depend on the width);
 *   - blockSize is accepted for signature compatibility (SourceCode/Parallel7.cu:645) and
 *     validated (> 0) but advisory: the kernels fix their own tile geometry and the output
 *     never depends on it.
 */
#ifndef B200SORT_H_
#define B200SORT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SORT_VERSION 100

enum {
    B200SORT_OK = 0,
    B200SORT_EINVAL = -1,     /* bad argument (null pointer, nBits outside 1..16, blockSize <= 0) */
    B200SORT_ETOOBIG = -2,    /* n > 2^32-1 */
    B200SORT_ETEMP = -3,      /* temp buffer too small or misaligned */
    B200SORT_EALIAS = -4,     /* d_out aliases d_in */
    B200SORT_ENODEVICE = -5,  /* no sm_100 device / CUDA runtime unusable */
    B200SORT_ENOMEM = -6,     /* host-pointer wrapper could not allocate */
    B200SORT_ENOPEER = -7     /* multi-GPU entry points: a selected device cannot reach a peer */
};

/* ---------------------------------------------------------------------------------------
 * Host-pointer entry points: same semantics as the reference's
 *   void sortByDevice(const uint32_t *h_input, int n, uint32_t *h_output, int numBits,
 *                     int blockSize)                       -- SourceCode/Parallel7.cu:530
 * (called from sort(), SourceCode/Parallel7.cu:641-662).  The caller owns the host arrays
 * (pageable or pinned); h_in is not modified; h_out is fully overwritten; h_out == h_in is
 * allowed.  Blocking.  Device buffers are cached by the library between calls and released
 * by b200sort_shutdown().  One call at a time per process (internally serialised).
 * Where sortByDevice copies the whole array in, sorts, and copies it out (Parallel7.cu:549, :624),
 * arrays of 2^24 keys and more with 8-bit digits are pipelined: chunked upload with the histogram of the
 * top 4 bits accumulated per chunk, one stable digit pass on those bits (16 buckets), then every bucket
 * is sorted on its low 28 bits and downloaded while the next one is being sorted -- the same stable
 * order (b200sort_set_param("host_overlap", 0) restores the single-shot form).
 */
int b200sort_keys_host(const uint32_t *h_in, uint64_t n, uint32_t *h_out, int nBits,
                       int blockSize);

/* Key/value extension named by the north star (the reference has no pairs path).  Stable:
 * equal keys keep their input order. */
int b200sort_pairs_host(const uint32_t *h_keys_in, const uint32_t *h_vals_in, uint64_t n,
                        uint32_t *h_keys_out, uint32_t *h_vals_out, int nBits, int blockSize);

/* Optional: pay the one-time costs of the host-pointer entry points now instead of inside the first sort --
 * device buffers for arrays of up to max_n keys (and values when pairs != 0), the pinned staging ring, streams,
 * the ranking self test of the device, and the first launch of every kernel of the default path.  The
 * reference pays them inside sortByDevice on every call (cudaMalloc x 6 + 3 static buffers, Parallel7.cu:547-559,
 * :203-218); here they are cached, so without this call only the first sort is slow. */
int b200sort_warmup(uint64_t max_n, int pairs);

/* Frees the cached device/pinned buffers of the host-pointer entry points (single- and
 * multi-GPU). */
int b200sort_shutdown(void);

/* ---------------------------------------------------------------------------------------
 * Single-process multi-GPU host-pointer sort: the same contract as b200sort_keys_host /
 * sortByDevice (SourceCode/Parallel7.cu:530), with the array sharded over `num_devices`
 * GPUs of one node driven from the calling thread.  The reference has no multi-GPU path;
 * this is the form its single-process main() (SourceCode/Parallel7.cu:696-775) could call.
 *   shard g = elements [g*n/G, (g+1)*n/G)  ->  H2D over each GPU's own PCIe link
 *   -> digit histogram of the partition byte (b200sort_histogram) -> splitters on bin edges
 *   -> one stable digit pass whose per-bin destinations are peer addresses in the owners'
 *      receive buffers (b200sort_digit_pass with d_bin_dst: partition fused with the NVLink
 *      exchange) -> local sort of each received range (b200sort_keys / b200sort_pairs)
 *   -> D2H into h_out at the range's global offset.
 * devices: `num_devices` CUDA ordinals (an ordinal may repeat: the shards then share that
 * GPU, which exercises the whole path on a one-GPU box); NULL = 0 .. num_devices-1;
 * num_devices <= 0 = every visible device.  Needs peer access between distinct devices
 * (B200SORT_ENOPEER otherwise).  n may exceed 2^32-1 as long as every shard and every
 * received range stays below 2^32 keys.  Blocking; buffers cached until
 * b200sort_mgpu_shutdown() / b200sort_shutdown().  The multi-process form of the same
 * algorithm (one rank per GPU over torch.distributed) is cuda/radixsort_b200/mgpu.py. */
int b200sort_mgpu_keys_host(const uint32_t *h_in, uint64_t n, uint32_t *h_out, int nBits,
                            int blockSize, const int *devices, int num_devices);
int b200sort_mgpu_pairs_host(const uint32_t *h_keys_in, const uint32_t *h_vals_in, uint64_t n,
                             uint32_t *h_keys_out, uint32_t *h_vals_out, int nBits,
                             int blockSize, const int *devices, int num_devices);

/* Host-side planning of the multi-GPU sorts, exported so that it can be exercised without a
 * GPU (tests compare it with the Python driver's planner, cuda/radixsort_b200/mgpu.py).
 * plan_owners: owner[b] = shard that receives bin b of the partition digit; cut j sits on the
 *   bin edge closest to j * total / num_shards, owners are non-decreasing in b.
 * plan_value_cuts: value splitters from a sample (see b200sort_route).  sample_keys/sample_pos
 *   hold the sampled keys of all shards, shard after shard (shard r = entries shard_offsets[r]
 *   .. shard_offsets[r+1]), with the local index each was taken from (ascending per shard).
 *   Writes num_shards-1 cuts: the key value, the shard whose run of that value is cut, and the
 *   local index in that shard from which equal keys go right (lower shards: left, higher: right;
 *   split_shard == num_shards: the whole run goes left). */
int b200sort_plan_owners(const uint64_t *hist, int bins, int num_shards, int *owner);
int b200sort_plan_value_cuts(const uint32_t *sample_keys, const uint64_t *sample_pos,
                             const uint64_t *shard_offsets, int num_shards, uint64_t *values,
                             int *split_shard, uint64_t *split_pos);

/* Figures of the last b200sort_mgpu_*_host call: out[0..6] = milliseconds of upload,
 * histogram, plan (host clock: splitters, receive buffers), partition kernel, wait for the
 * peers' partitions, local sort, download -- device-event times, each the maximum over the
 * devices, so they need not add up to the call's wall time; out[7] = partition shift,
 * out[8] = partition bits, out[9] = largest received range / (n / G), out[10] = devices, out[11] = 1 when
 * the splitters were key values from a sample (skewed keys; see b200sort_route) instead of bin
 * edges of the partition byte.
 * Returns the number of values written (<= capacity). */
enum { B200SORT_MGPU_STATS = 12 };
int b200sort_mgpu_last_stats(double *out, int capacity);
int b200sort_mgpu_shutdown(void);

/* ---------------------------------------------------------------------------------------
 * Device-resident entry points: the per-digit loop of sortByDevice
 * (SourceCode/Parallel7.cu:561-622: sortLocallyDataBlocks + histogram + transpose/scan/
 * transpose + scatter) without its cudaMalloc/H2D/D2H shell.  Asynchronous on `stream`
 * (a cudaStream_t; NULL = default stream).  The caller owns every buffer.  d_in is not
 * modified; d_out must not overlap d_in.  d_temp must be 256-byte aligned and at least
 * b200sort_temp_bytes() long.  n == 0 is a no-op (undefined in the reference,
 * SourceCode/Parallel7.cu:157).
 */
size_t b200sort_temp_bytes(uint64_t n, int nBits, int pairs);

int b200sort_keys(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp,
                  size_t temp_bytes, int nBits, void *stream);

int b200sort_pairs(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                   uint32_t *d_keys_out, uint32_t *d_vals_out, void *d_temp,
                   size_t temp_bytes, int nBits, void *stream);

/* The same sorts for keys that are known to agree in all bits >= key_bits (1..32): only the digits
 * below key_bits are sorted -- the bucket of an MSD partition, a shard of the multi-GPU sort whose
 * splitters fixed the top bits.  The result is the full sort if the promise holds, otherwise the keys
 * are ordered by their low key_bits bits only (stable).  No counterpart in the reference, whose loop
 * always covers all 32 bits (SourceCode/Parallel7.cu:561); temp storage as for the full sort. */
int b200sort_keys_low_bits(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp,
                           size_t temp_bytes, int nBits, int key_bits, void *stream);

int b200sort_pairs_low_bits(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                            uint32_t *d_keys_out, uint32_t *d_vals_out, void *d_temp,
                            size_t temp_bytes, int nBits, int key_bits, void *stream);

/* ---------------------------------------------------------------------------------------
 * Building blocks, exported for the multi-GPU driver (histogram -> splitters -> partition ->
 * exchange -> local sort) and as the public form of the reference's stage wrappers.
 */

/* Digit histogram of (key >> shift) & (2^bits - 1), bits in 1..8; d_hist[2^bits] uint32 is
 * overwritten.  Replaces histogram() + histogramKernel, SourceCode/Parallel7.cu:318-359,
 * reduced over tiles (the tile x bin table itself never exists in this design). */
int b200sort_histogram(const uint32_t *d_keys, uint64_t n, int shift, int bits,
                       uint32_t *d_hist, void *d_temp, size_t temp_bytes, void *stream);

/* One stable counting-sort pass on the digit (key >> shift) & (2^bits - 1), bits in 1..8:
 * out = keys ordered by that digit, ties in input order.  This is one iteration of the
 * reference's digit loop (SourceCode/Parallel7.cu:561-622) and, with shift = 32 - bits, the
 * MSD range partition of the multi-GPU sort.  d_vals_in/d_vals_out may both be NULL.
 * If d_bin_dst is non-NULL it holds 2^bits device addresses (uint64): the keys of bin d
 * are written to ((uint32_t*)d_bin_dst[d])[0 .. count_d) instead of d_keys_out (and values
 * to d_bin_dst[2^bits + d]) -- the addresses may be peer (NVLink) memory, which fuses the
 * partition with the exchange.  Temp size: b200sort_temp_bytes(n, bits, pairs). */
int b200sort_digit_pass(const uint32_t *d_keys_in, const uint32_t *d_vals_in, uint64_t n,
                        uint32_t *d_keys_out, uint32_t *d_vals_out, int shift, int bits,
                        const uint64_t *d_bin_dst, void *d_temp, size_t temp_bytes,
                        void *stream);

/* d_route[i] = number of cuts j with (value_j, tie_j) <= (d_keys[i], i) in lexicographic order:
 * the destination shard of every key under VALUE splitters.  d_thresholds holds `count` (<= 255)
 * values in [0, 2^32] followed by `count` tie indices; a key equal to value_j is at or above cut
 * j from local index tie_j on (tie 0: the whole run of equal keys goes right; tie >= n: left).
 * Cuts must be non-decreasing.  Used by the multi-GPU drivers when the bin-edge splitters of the
 * partition byte leave the shards unbalanced (skewed keys): the route array is then the key of a
 * digit pass (b200sort_digit_pass, shift 0) that carries the real keys as values to their
 * owners; the tie index cuts a run of equal keys at a position, which keeps ties in input order.
 * d_counts (may be NULL): count+1 uint32, overwritten with the number of keys routed to each
 * destination -- the row of the exchange matrix this shard contributes.
 * No counterpart in the reference (single GPU). */
int b200sort_route(const uint32_t *d_keys, uint64_t n, const uint64_t *d_thresholds, int count,
                   uint32_t *d_route, uint32_t *d_counts, void *stream);

/* Device-wide exclusive prefix sum of uint32 (mod 2^32), one pass, decoupled look-back.  The
 * public form of the reference's scan stage: scan() + scanBlocks + addScannedBlockSumsToScannedBlocks
 * with its host round trip (SourceCode/Parallel7.cu:408-528), and of Docs/Snippets/
 * PrefixSum-WorkEfficient.cu.  d_in / d_out 16-byte aligned; d_out may equal d_in.
 * Temp: b200sort_scan_temp_bytes(n), 256-byte aligned. */
size_t b200sort_scan_temp_bytes(uint64_t n);
int b200sort_exclusive_scan(const uint32_t *d_in, uint64_t n, uint32_t *d_out, void *d_temp,
                            size_t temp_bytes, void *stream);

/* ---------------------------------------------------------------------------------------
 * Measurement and test utilities (device side of SURVEY.md section 8d's workloads).
 */
enum {
    B200SORT_GEN_UNIFORM = 0, B200SORT_GEN_ZIPF = 1, B200SORT_GEN_UNIQUE16 = 2,
    B200SORT_GEN_ALL_EQUAL = 3, B200SORT_GEN_SORTED = 4, B200SORT_GEN_REVERSED = 5,
    B200SORT_GEN_IOTA = 6
};

/* d_out[j] = key(first + j), j < count; `total` scales the sorted/reversed ramps;
 * d_zipf_cdf (65536 uint32) is only read for B200SORT_GEN_ZIPF. */
int b200sort_generate(uint32_t *d_out, uint64_t first, uint64_t count, int kind,
                      uint64_t total, const uint32_t *d_zipf_cdf, void *stream);

/* d_result[0] = number of i with keys[i-1] > keys[i]; d_result[1..3] = order-independent
 * multiset fingerprint (sum key, sum sm64(key), xor sm64(key)).  d_result: 4 x uint64,
 * overwritten. */
int b200sort_verify(const uint32_t *d_keys, uint64_t n, uint64_t *d_result, void *stream);

/* Bandwidth probe used to size the fused multi-GPU exchange: copies n uint32 from d_src to d_dst
 * (which may be peer memory) with 4-byte (vec = 1) or 16-byte (vec = 4) stores per lane. */
int b200sort_store_probe(uint32_t *d_dst, const uint32_t *d_src, uint64_t n, int vec, int ctas_per_sm,
                         void *stream);

/* Per-kernel device timing: while enabled, every sort records CUDA events on its own stream
 * around each kernel.  b200sort_profile_read() synchronises on the last recorded event,
 * drains the records of all sorts since the previous read and returns how many it wrote:
 * ms[i] is a kernel duration, tag[i] says which kernel (0 = histogram kernel, p+1 = digit
 * pass p, all launches of that pass together). */
int b200sort_profile_enable(int on);
int b200sort_profile_read(float *ms, int *tag, int capacity);

/* Number of kernels this library has launched since load (the bench's gpu_launches). */
uint64_t b200sort_launch_count(void);

/* Tuning hooks (bench/test use): "variant" selects the digit-pass kernel by its number in
 * csrc/launch.h (-1 = automatic: the column sweep with two ranking chains for digits of >= 4 bits,
 * ballot rank below), "portion_tiles" caps tiles per launch (0 = default; tests use it to
 * exercise the multi-launch path at small n), "hist_ctas_per_sm", "safe_rank" (1 = an
 * atomic-rank request is served by the column sweep), "prefetch_tiles" (L2 prefetch distance
 * of the default kernel in tiles; -1 = one per SM, 0 = off), "dst_bulk", "host_overlap",
 * "scan_variant", "scan_prefetch_tiles" (the scan's L2 prefetch distance in tiles; -1 = 8 MiB
 * ahead, 0 = off).  Returns B200SORT_EINVAL for an unknown name or value. */
int b200sort_set_param(const char *name, int value);
int b200sort_get_param(const char *name);

/* Geometry of the selected digit-pass kernel (keys per tile), algorithmic byte count
 * 4n(2P+1) / 4n(4P+1) of SURVEY.md section 8d, and number of digit passes actually run. */
int b200sort_tile_keys(int pairs);
uint64_t b200sort_algorithmic_bytes(uint64_t n, int nBits, int pairs);
int b200sort_num_passes(int nBits);

/* Device banner in the layout of the reference's printDeviceInfo() (SourceCode/Parallel7.cu:
 * 664-677), for drivers that reproduce its log. */
int b200sort_device_banner(char *buf, size_t len);

int b200sort_version(void);
const char *b200sort_error_string(int code);
const char *b200sort_last_error_string(void);

#ifdef __cplusplus
}
#endif
#endif /* B200SORT_H_ */
