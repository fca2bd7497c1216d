/*
 * radix_oracle.c -- CPU restatement of the reference's LSD radix sort.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker (or as the timed CPU baseline), never as a fallback for
 * the CUDA path.  The product library (libb200sort.so) has no CPU sort in it.
 *
 * Parity status: PINNED for keys.  tests/test_oracle.py checks this file against the two
 * known-answer vectors the reference's own test procedure defines (glibc rand(), never
 * seeded; SourceCode/Baseline1.cu:140-160, DEBUG and default configs; fingerprints in
 * SURVEY.md section 8c) and, when /root/reference is present, against the reference's own
 * sortByHost compiled unmodified into oracle/_ref/ (see oracle/build_oracle.py).
 * Pairs: the reference has no key/value path, so pair parity is "unpinned by the
 * reference"; oracle_sort_pairs is the same counting sort carrying a payload and is
 * cross-checked against a stable sort by key in the tests.
 *
 * Each function cites the reference lines it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* Number of digit passes the reference loop performs:
 * `for (bit = 0; bit < 32; bit += nBits)` -- SourceCode/Baseline1.cu:30. */
ORACLE_API int oracle_num_passes(int nbits) {
    if (nbits <= 0) return 0;
    return (32 + nbits - 1) / nbits;
}

/*
 * One stable counting-sort pass on the digit (key >> shift) & (bins - 1).
 * Restates the loop body of sortByHost, SourceCode/Baseline1.cu:31-49:
 *   histogram (:32-36), exclusive scan (:39-42), stable scatter in index order (:45-49).
 * `vsrc`/`vdst` may be NULL (keys only).
 */
static void counting_pass(const uint32_t *ksrc, uint32_t *kdst, const uint32_t *vsrc,
                          uint32_t *vdst, int64_t n, int shift, int64_t bins,
                          int64_t *count, int64_t *cursor) {
    const uint32_t digit_mask = (uint32_t)(bins - 1);
    memset(count, 0, (size_t)bins * sizeof(int64_t));
    for (int64_t i = 0; i < n; ++i) count[(ksrc[i] >> shift) & digit_mask] += 1;

    int64_t running = 0;
    for (int64_t b = 0; b < bins; ++b) {
        cursor[b] = running;
        running += count[b];
    }

    if (vsrc) {
        for (int64_t i = 0; i < n; ++i) {
            const uint32_t k = ksrc[i];
            const int64_t at = cursor[(k >> shift) & digit_mask]++;
            kdst[at] = k;
            vdst[at] = vsrc[i];
        }
    } else {
        for (int64_t i = 0; i < n; ++i) {
            const uint32_t k = ksrc[i];
            kdst[cursor[(k >> shift) & digit_mask]++] = k;
        }
    }
}

/*
 * Whole sort.  Restates sortByHost, SourceCode/Baseline1.cu:15-64: copy the input to a
 * scratch array (:22-24), ping-pong between scratch and `out` once per digit (:30-55),
 * copy the final source array to `out` (:58).  nbits in 1..16; the last digit is narrower
 * when 32 % nbits != 0 because the shift never reaches 32 (:30).
 * Returns 0, or -1 on bad arguments / allocation failure.
 */
static int sort_impl(const uint32_t *kin, const uint32_t *vin, int64_t n, uint32_t *kout,
                     uint32_t *vout, int nbits) {
    if (nbits < 1 || nbits > 16 || n < 0) return -1;
    if (n == 0) return 0;
    const int64_t bins = (int64_t)1 << nbits;
    int64_t *count = (int64_t *)malloc((size_t)bins * sizeof(int64_t));
    int64_t *cursor = (int64_t *)malloc((size_t)bins * sizeof(int64_t));
    uint32_t *kscratch = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
    uint32_t *vscratch = vin ? (uint32_t *)malloc((size_t)n * sizeof(uint32_t)) : NULL;
    if (!count || !cursor || !kscratch || (vin && !vscratch)) {
        free(count); free(cursor); free(kscratch); free(vscratch);
        return -1;
    }
    memcpy(kscratch, kin, (size_t)n * sizeof(uint32_t));
    if (vin) memcpy(vscratch, vin, (size_t)n * sizeof(uint32_t));

    uint32_t *ka = kscratch, *kb = kout, *va = vscratch, *vb = vout;
    for (int shift = 0; shift < 32; shift += nbits) {
        counting_pass(ka, kb, vin ? va : NULL, vb, n, shift, bins, count, cursor);
        uint32_t *t = ka; ka = kb; kb = t;
        t = va; va = vb; vb = t;
    }
    if (ka != kout) {
        memcpy(kout, ka, (size_t)n * sizeof(uint32_t));
        if (vin) memcpy(vout, va, (size_t)n * sizeof(uint32_t));
    }
    free(count); free(cursor); free(kscratch); free(vscratch);
    return 0;
}

ORACLE_API int oracle_sort_keys(const uint32_t *in, int64_t n, uint32_t *out, int nbits) {
    return sort_impl(in, NULL, n, out, NULL, nbits);
}

/* Key/value variant: same passes, payload moved with its key (stable). */
ORACLE_API int oracle_sort_pairs(const uint32_t *kin, const uint32_t *vin, int64_t n,
                                 uint32_t *kout, uint32_t *vout, int nbits) {
    if (!vin || !vout) return -1;
    return sort_impl(kin, vin, n, kout, vout, nbits);
}

/*
 * Intermediate-state oracle: the tile x bin histogram table and its bin-major exclusive
 * scan for ONE digit, as the reference's GPU algorithm defines them.
 * Restates SourceCode/Baseline4.cu:103-114 (per-tile histogram) and :127-138 (scan in
 * column-major order: all tiles of bin 0, then bin 1, ...), i.e.
 *   scan[t][d] = sum_{d'<d} sum_{t'} cnt[t'][d'] + sum_{t'<t} cnt[t'][d].
 * Both tables are row-major [tiles][bins].  Returns the number of tiles, or -1.
 */
ORACLE_API int64_t oracle_tile_table(const uint32_t *in, int64_t n, int64_t tile, int shift,
                                     int nbits, uint32_t *table, uint32_t *scan) {
    if (tile <= 0 || nbits < 1 || nbits > 16 || n < 0) return -1;
    const int64_t bins = (int64_t)1 << nbits;
    const int64_t tiles = (n + tile - 1) / tile;
    memset(table, 0, (size_t)(tiles * bins) * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i)
        table[(i / tile) * bins + ((in[i] >> shift) & (uint32_t)(bins - 1))] += 1;
    uint32_t running = 0;
    for (int64_t d = 0; d < bins; ++d)
        for (int64_t t = 0; t < tiles; ++t) {
            scan[t * bins + d] = running;
            running += table[t * bins + d];
        }
    return tiles;
}

/* FNV-1a-64 over uint32 words; the fingerprint SURVEY.md 8c quotes the KATs in. */
ORACLE_API uint64_t oracle_fnv1a64_words(const uint32_t *w, int64_t n) {
    uint64_t h = 1469598103934665603ULL;
    for (int64_t i = 0; i < n; ++i) {
        h ^= (uint64_t)w[i];
        h *= 1099511628211ULL;
    }
    return h;
}

/*
 * The reference's own test input: `input[i] = rand()` (or `rand() & 0xFF` under DEBUG),
 * never seeded -- SourceCode/Baseline1.cu:152-158.  glibc's unseeded rand() is srand(1).
 */
ORACLE_API void oracle_fill_glibc_rand(uint32_t *out, int64_t n, uint32_t and_mask) {
    srand(1);
    for (int64_t i = 0; i < n; ++i) out[i] = (uint32_t)rand() & and_mask;
}

/* ------------------------------------------------------------------------------------ */
/* Synthetic workloads of SURVEY.md 8d, generated from a counter hash so the host and   */
/* the device generator (csrc/generators.cuh) produce the same bytes.                   */

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

enum { GEN_UNIFORM = 0, GEN_ZIPF = 1, GEN_UNIQUE16 = 2, GEN_ALL_EQUAL = 3, GEN_SORTED = 4,
       GEN_REVERSED = 5, GEN_IOTA = 6 };

#define ZIPF_VALUES 65536

/* Zipf(1.0) over ZIPF_VALUES ranks by inverse CDF on a precomputed table (see
 * oracle_zipf_cdf); value of rank r is splitmix64(r) >> 32. */
ORACLE_API void oracle_zipf_cdf(uint32_t *cdf /* [ZIPF_VALUES] */) {
    double total = 0.0;
    for (int r = 0; r < ZIPF_VALUES; ++r) total += 1.0 / (double)(r + 1);
    double acc = 0.0;
    for (int r = 0; r < ZIPF_VALUES; ++r) {
        acc += 1.0 / (double)(r + 1);
        double q = acc / total * 4294967296.0;
        cdf[r] = q >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)q;
    }
    cdf[ZIPF_VALUES - 1] = 0xFFFFFFFFu;
}

static uint32_t zipf_rank(const uint32_t *cdf, uint32_t u) {
    /* first r with cdf[r] >= u */
    int lo = 0, hi = ZIPF_VALUES - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cdf[mid] >= u) hi = mid; else lo = mid + 1;
    }
    return (uint32_t)lo;
}

/*
 * key[i] for global index first+i.  `cdf` is only read for GEN_ZIPF.
 *   uniform   : sm64(0x5EED0001 + i) >> 32
 *   zipf      : v[rank], rank by inverse CDF from u = sm64(0x5EED0004 + i) >> 32
 *   unique16  : v[sm64(0x5EED0005 + i) & 15]
 *   all_equal : 0xDEADBEEF
 *   sorted    : i scaled onto the 32-bit range (non-decreasing), reversed: its complement
 *   iota      : (uint32) i
 */
ORACLE_API int oracle_generate(uint32_t *out, int64_t first, int64_t count, int kind,
                               int64_t total, const uint32_t *cdf) {
    for (int64_t j = 0; j < count; ++j) {
        const uint64_t i = (uint64_t)(first + j);
        uint32_t k;
        switch (kind) {
        case GEN_UNIFORM: k = (uint32_t)(splitmix64(0x5EED0001ULL + i) >> 32); break;
        case GEN_ZIPF:
            if (!cdf) return -1;
            k = (uint32_t)(splitmix64((uint64_t)zipf_rank(
                    cdf, (uint32_t)(splitmix64(0x5EED0004ULL + i) >> 32))) >> 32);
            break;
        case GEN_UNIQUE16:
            k = (uint32_t)(splitmix64(splitmix64(0x5EED0005ULL + i) & 15ULL) >> 32);
            break;
        case GEN_ALL_EQUAL: k = 0xDEADBEEFu; break;
        case GEN_SORTED:
        case GEN_REVERSED: {
            /* floor(i * 2^32 / total) without overflow for total <= 2^32 */
            unsigned __int128 s = ((unsigned __int128)i << 32) / (uint64_t)(total > 0 ? total : 1);
            k = (uint32_t)s;
            if (kind == GEN_REVERSED) k = ~k;
            break;
        }
        case GEN_IOTA: k = (uint32_t)i; break;
        default: return -1;
        }
        out[j] = k;
    }
    return 0;
}

/* Order-independent multiset fingerprint (sum of keys, sum of sm64(key), xor of sm64(key)),
 * used to check the 2^32-key sharded sort without holding it on one host. */
ORACLE_API void oracle_multiset_fingerprint(const uint32_t *k, int64_t n, uint64_t out[3]) {
    uint64_t s = 0, h = 0, x = 0;
    for (int64_t i = 0; i < n; ++i) {
        const uint64_t m = splitmix64((uint64_t)k[i]);
        s += k[i]; h += m; x ^= m;
    }
    out[0] = s; out[1] = h; out[2] = x;
}

/* 1 if non-decreasing. */
ORACLE_API int oracle_is_sorted(const uint32_t *k, int64_t n) {
    for (int64_t i = 1; i < n; ++i) if (k[i - 1] > k[i]) return 0;
    return 1;
}
