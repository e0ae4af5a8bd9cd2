// Shared host-side plumbing of the C ABI: thread-local error string, launch counter.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdlib>
#ifndef PNNP_HOST_EMUL
#include <cuda_runtime.h>
#endif

namespace pnnp {
int fail(const char* msg);                 // records msg, returns 1
int fail_cuda(cudaError_t e, const char* what, const char* file, int line);
void count_launch(uint64_t n = 1);
// Kernel variants measured on a B200 in round 2 (tools/r02_sweep.sh, profiles/r02_sweep_summary.txt) and promoted to defaults;
// NAME=0 in the environment switches one off again (A/B timing, bisecting).  Read per launch: tests flip them.
inline bool variant_on(const char* name) { const char* e = getenv(name); return !e || atoi(e) > 0; }
}  // namespace pnnp

#define PNNP_CUDA(expr)                                                             \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) return ::pnnp::fail_cuda(_e, #expr, __FILE__, __LINE__); \
    } while (0)
