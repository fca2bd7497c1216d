// radixsort_cli.cpp -- command line driver with the reference's interface and log format
// (SourceCode/Parallel7.cu:696-775: `prog [blockSize] [numBits]`, device banner, "Input size",
// "Block size", "Digit width", "Radix Sort by ...", "Time: ... ms", "CORRECT :)"), running the
// sm_100a library through include/radix_sort_compat.hpp.  C++ host code above the C ABI.
//
//   radixsort_cli [blockSize=512] [numBits=8] [n=(1<<24)+1]
//
// Differences from the reference main(): a third optional argument n; the process exit code is
// non-zero when a check prints INCORRECT (the reference always returns 0); "by host" is served
// by std::stable_sort, or -- when the environment variable B200SORT_REF_LIB names a shared
// object built from the reference's Baseline1.cu -- by the reference's own sortByHost.  Either
// way it is a checker in this harness only; the library itself has no CPU path.  The environment
// variable B200SORT_GPUS=G shards the device sort over G GPUs of this node (0 = all of them).
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "radix_sort_compat.hpp"

extern "C" int b200sort_device_banner(char *buf, size_t len);

static void host_sort_std(const uint32_t *in, int n, uint32_t *out, int) {
    std::copy(in, in + n, out);
    std::stable_sort(out, out + n);
}

static bool checkCorrectness(const uint32_t *out, const uint32_t *correctOut, int n) {
    for (int i = 0; i < n; i++) {
        if (out[i] != correctOut[i]) {
            printf("INCORRECT :(\n");
            return false;
        }
    }
    printf("CORRECT :)\n");
    return true;
}

int main(int argc, char **argv) {
    char banner[1024];
    if (b200sort_device_banner(banner, sizeof banner) != 0) {
        fprintf(stderr, "Error: %s:%d, code: %d, reason: %s\n", __FILE__, __LINE__, -5, b200sort_last_error_string());
        return EXIT_FAILURE;
    }
    fputs(banner, stdout);

    int blockSize = argc > 1 ? atoi(argv[1]) : 512;
    int numBits = argc > 2 ? atoi(argv[2]) : 8;
    int n = argc > 3 ? atoi(argv[3]) : (1 << 24) + 1;
    printf("\nInput size: %d\n", n);
    std::vector<uint32_t> input(n), output(n), correct(n);
    for (int i = 0; i < n; i++) input[i] = rand();      // unseeded, like the reference
    printf("Block size: %d\n", blockSize);
    printf("Digit width: %d-bit\n", numBits);
    if (const char *gpus = getenv("B200SORT_GPUS")) {
        b200compat::device_count() = atoi(gpus);
        printf("GPUs: %s\n", b200compat::device_count() > 0 ? gpus : "all");
    }

    // "by host": the reference's own function when the caller points at a build of it
    std::string ref = getenv("B200SORT_REF_LIB") ? getenv("B200SORT_REF_LIB") : "";
    b200compat::HostSortFn host = nullptr;
    if (!ref.empty()) {
        if (void *h = dlopen(ref.c_str(), RTLD_NOW | RTLD_LOCAL))
            host = (b200compat::HostSortFn)dlsym(h, "_Z10sortByHostPKjiPji");   // SourceCode/Baseline1.cu:15
    }
    printf("Host reference: %s\n", host ? ref.c_str() : "std::stable_sort");
    b200compat::set_host_sort(host ? host : host_sort_std);

    sort(input.data(), n, correct.data(), SORT_BY_HOST, numBits);

    std::fill(output.begin(), output.end(), 0u);
    sort(input.data(), n, output.data(), SORT_BY_DEVICE, numBits, blockSize);
    bool ok = checkCorrectness(output.data(), correct.data(), n);

    // the north star's bool spelling, same data
    std::fill(output.begin(), output.end(), 0u);
    b200compat::default_nbits() = numBits;
    sort(input.data(), n, output.data(), true, blockSize);
    ok = checkCorrectness(output.data(), correct.data(), n) && ok;

    b200sort_shutdown();
    return ok ? EXIT_SUCCESS : EXIT_FAILURE;
}
