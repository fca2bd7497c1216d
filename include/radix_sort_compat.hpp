// radix_sort_compat.hpp -- header-only C++ shim that gives a program written against
// truongchauhien/CUDA.RadixSort the SAME free functions it calls today, on top of the C ABI of
// libb200sort.so (include/b200sort.h).
//
//   reference (SourceCode/Parallel7.cu)                      this header
//   ----------------------------------------------------     -----------------------------------
//   typedef enum {SORT_BY_HOST, SORT_BY_THRUST,              same enum, same values (:22)
//                 SORT_BY_DEVICE} Implementation;
//   void sortByDevice(const uint32_t*, int, uint32_t*,       same signature (:530) ->
//                     int numBits, int blockSize);             b200sort_keys_host
//   void sort(const uint32_t* in, int n, uint32_t* out,      same signature and defaults (:641-645),
//             Implementation = SORT_BY_HOST,                   same banner / "Time: %.3f ms" output
//             int numBits = 4, int blockSize = 1);             (:650-661)
//   -- (north star spelling)                                  void sort(in, n, out, bool useDevice,
//                                                                       int blockSize)
//   -- (no multi-GPU path)                                    b200compat::device_count() = G makes
//                                                               sortByDevice shard over G GPUs
//                                                               (b200sort_mgpu_keys_host)
//
// Error behaviour follows the reference's CHECK macro (SourceCode/common/common.h:6-16): on
// failure print "Error: file:line, code: N, reason: ..." to stderr and exit(EXIT_FAILURE).
//
// SORT_BY_HOST / SORT_BY_THRUST / useDevice = false: the product library has no CPU path and
// links no Thrust.  A harness that wants them (the reference's own main() calls all three)
// registers its own implementation with b200compat::set_host_sort() (csrc/radixsort_cli.cpp
// registers std::stable_sort, or the reference's sortByHost from a user-named shared object);
// without a registered function those modes are a CHECK-style fatal error, never a silent
// substitute.
#ifndef RADIX_SORT_COMPAT_HPP_
#define RADIX_SORT_COMPAT_HPP_

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "b200sort.h"

typedef enum { SORT_BY_HOST, SORT_BY_THRUST, SORT_BY_DEVICE } Implementation;

namespace b200compat {

typedef void (*HostSortFn)(const uint32_t *in, int n, uint32_t *out, int nBits);

inline HostSortFn &host_sort_slot() {
    static HostSortFn fn = nullptr;
    return fn;
}
// Registers the function that serves SORT_BY_HOST (and SORT_BY_THRUST) in test harnesses.
inline void set_host_sort(HostSortFn fn) { host_sort_slot() = fn; }

// Digit width used by the bool spelling of sort(), which has no nBits argument.
inline int &default_nbits() {
    static int nbits = 8;
    return nbits;
}

// Number of GPUs sortByDevice() shards the array over (1 = the single-GPU entry point;
// 0 = every visible device).  The reference is single-GPU; this is the one-line opt-in.
inline int &device_count() {
    static int count = 1;
    return count;
}

inline void check(int rc, const char *file, int line) {
    if (rc != B200SORT_OK) {
        std::fprintf(stderr, "Error: %s:%d, ", file, line);
        std::fprintf(stderr, "code: %d, reason: %s\n", rc, b200sort_last_error_string());
        std::exit(EXIT_FAILURE);
    }
}

}  // namespace b200compat

#define B200_CHECK(call) ::b200compat::check((call), __FILE__, __LINE__)

// Drop-in for the reference's sortByDevice (SourceCode/Parallel7.cu:530).
inline void sortByDevice(const uint32_t *h_input, int n, uint32_t *h_output, int numBits, int blockSize) {
    const uint64_t count = n < 0 ? 0 : (uint64_t)n;
    if (b200compat::device_count() == 1)
        B200_CHECK(b200sort_keys_host(h_input, count, h_output, numBits, blockSize));
    else
        B200_CHECK(b200sort_mgpu_keys_host(h_input, count, h_output, numBits, blockSize, nullptr,
                                           b200compat::device_count()));
}

// Drop-in for the reference's sort() (SourceCode/Parallel7.cu:641-662), including its output.
inline void sort(const uint32_t *in, int n, uint32_t *out, Implementation implementation = SORT_BY_HOST,
                 int numBits = 4, int blockSize = 1) {
    const auto t0 = std::chrono::steady_clock::now();
    if (implementation == SORT_BY_DEVICE) {
        std::printf("\nRadix Sort by device:\n");
        sortByDevice(in, n, out, numBits, blockSize);
    } else {
        std::printf(implementation == SORT_BY_HOST ? "\nRadix Sort by host\n" : "\nRadix Sort by Thrust library\n");
        if (b200compat::host_sort_slot() == nullptr) {
            std::fprintf(stderr, "Error: %s:%d, ", __FILE__, __LINE__);
            std::fprintf(stderr, "code: %d, reason: %s\n", B200SORT_EINVAL,
                         "SORT_BY_HOST/SORT_BY_THRUST requested but libb200sort has no CPU or Thrust path "
                         "(register one with b200compat::set_host_sort)");
            std::exit(EXIT_FAILURE);
        }
        b200compat::host_sort_slot()(in, n, out, numBits);
    }
    const std::chrono::duration<float, std::milli> dt = std::chrono::steady_clock::now() - t0;
    std::printf("Time: %.3f ms\n", dt.count());
}

// The north star's spelling: sort(in, n, out, useDevice, blockSize).
inline void sort(const uint32_t *in, int n, uint32_t *out, bool useDevice, int blockSize) {
    sort(in, n, out, useDevice ? SORT_BY_DEVICE : SORT_BY_HOST, b200compat::default_nbits(), blockSize);
}

#endif  // RADIX_SORT_COMPAT_HPP_
