#include "abi_common.h"
#include <cstdio>
#include <cstring>
#include "../../include/pnnp_b200.h"

namespace pnnp {
static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

int fail(const char* msg) {
    std::snprintf(g_err, sizeof(g_err), "%s", msg);
    return 1;
}
int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    std::snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d in `%s`", (int)e, cudaGetErrorString(e), file, line, what);
    return 2;
}
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace pnnp

extern "C" const char* pnnp_last_error(void) { return pnnp::g_err; }
extern "C" int pnnp_abi_version(void) { return PNNP_ABI_VERSION; }
extern "C" uint64_t pnnp_launch_count(void) { return pnnp::g_launches.load(); }
// kernels launched through a replayed CUDA graph (captured launches are counted once, at capture time)
extern "C" void pnnp_count_graph_launches(uint64_t n) { pnnp::count_launch(n); }
