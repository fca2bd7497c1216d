// internal.h -- helpers shared by the translation units of libb200sort.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

namespace b200sort {

// Record the message b200sort_last_error_string() returns on this thread; returns `code`.
int set_error(int code, const char *what);
int set_cuda_error(cudaError_t e, const char *what);
void set_error_message(const char *message);

}  // namespace b200sort
