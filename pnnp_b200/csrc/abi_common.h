// Shared host-side plumbing of the C ABI: thread-local error string, launch counter.
#pragma once
#include <atomic>
#include <cstdint>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif

namespace pnnp {
int fail(const char* msg);                 // records msg, returns 1
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);
void count_launch(uint64_t n = 1);
}  // namespace pnnp

#define PNNP_CUDA(expr)                                                             \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::pnnp::fail_cuda(_e, #expr, __FILE__, __LINE__); \
    } while (0)
