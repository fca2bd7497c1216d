// internal.h -- helpers shared by the translation units of libb200sort.so (not part of the ABI).
#pragma once
#include <atomic>

#include <cuda_runtime.h>

namespace b200sort {

// Record the message b200sort_last_error_string() returns on this thread; returns `code`.
int set_error(int code, const char *what);
int set_cuda_error(cudaError_t e, const char *what);
void set_error_message(const char *message);

// b200sort_set_param("mgpu_balance_permille"): bin-edge splitters that leave one shard above this
// share of the mean (in 1/1000) switch b200sort_mgpu_*_host to value splitters; 0 = never.
extern std::atomic<int> g_mgpu_balance_permille;

}  // namespace b200sort
