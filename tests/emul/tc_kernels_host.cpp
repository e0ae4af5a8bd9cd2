// pnnp_b200/csrc/conv_tc.cu ITSELF — the tcgen05 / TMA implicit-GEMM convolution kernel, its launcher and its tensor-map set-up —
// compiled for the HOST against the functional tensor-core model (tc_host_model.h) and run on the fibre SIMT emulator (simt_host.h).
// TEST INFRASTRUCTURE ONLY (tests/test_device_tc_on_cpu.py).
#define PNNP_HOST_EMUL 1
#include "cuda_host_shim.h"
#include "simt_host.h"
#include "tc_host_model.h"

namespace pnnp {
__attribute__((aligned(1024))) uint8_t smem_raw[kSmemBytes];
static char g_err_msg[512] = "";
int fail(const char* msg) { std::snprintf(g_err_msg, sizeof g_err_msg, "%s", msg); return 1; }
int fail_cuda(cudaError_t e, const char* what, const char* file, int line) {
    std::snprintf(g_err_msg, sizeof g_err_msg, "CUDA error %d at %s:%d in `%s`", (int)e, file, line, what);
    return 2;
}
static uint64_t g_launches = 0;
void count_launch(uint64_t n) { g_launches += n; }
}  // namespace pnnp

#include "../../pnnp_b200/csrc/conv_tc.cu"
#include "../../pnnp_b200/csrc/wgrad_nhwc_tc.cu"
#include "../../pnnp_b200/csrc/conv_first.cu"

extern "C" {
int emul_conv2d_tc_ex(const pnnp_conv_desc* d) { return pnnp::conv_layer_launch(*d, nullptr); }
const char* emul_tc_last_error(void) { return pnnp::g_err_msg; }
int emul_conv_pipeline_error(void) { return pnnp_conv_pipeline_error(); }
int emul_wgrad_nhwc(int mode, const void* g, int co, int co_stride, const void* x, int ci, int ci_stride, int n, int h, int w, float* dw,
                    int ci_off, int ci_total, int co_pad) {
    return pnnp_wgrad_nhwc(mode, g, co, co_stride, x, ci, ci_stride, n, h, w, dw, ci_off, ci_total, co_pad, nullptr);
}
int emul_wgrad_pipeline_error(void) { return pnnp_wgrad_nhwc_pipeline_error(); }
int emul_conv_first_nchw(const float* in, const float* w, const float* b, void* out, int n, int cin, int h, int wd, int cout, int act) {
    return pnnp_conv_first_nchw(in, w, b, out, n, cin, h, wd, cout, act, nullptr);
}
int emul_conv_first_pipeline_error(void) { return pnnp_conv_first_pipeline_error(); }
// counters of the model since the last call: [mma instructions, TMA loads, zero-filled elements, mbarrier waits]
unsigned long emul_tc_trailing_commits(void) { const unsigned long v = pnnp::g_trailing_commits; pnnp::g_trailing_commits = 0; return v; }
void emul_tc_stats(unsigned long* out4) {
    out4[0] = pnnp::g_tc_stats.mma; out4[1] = pnnp::g_tc_stats.tma; out4[2] = pnnp::g_tc_stats.tma_oob_elems; out4[3] = pnnp::g_tc_stats.waits;
    pnnp::g_tc_stats = pnnp::TcStats();
}
}
