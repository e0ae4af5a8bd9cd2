// The device core of the noise synthesis (pnnp_b200/csrc/noise_core.cuh) compiled for the HOST — TEST INFRASTRUCTURE ONLY.
// Built by tests/test_device_core_on_cpu.py with  g++ -O2 -ffp-contract=off -shared -fPIC ; never linked into the product.
#define PNNP_HOST_EMUL 1
#include "cuda_host_shim.h"
#include "../../pnnp_b200/csrc/noise_core.cuh"
#include "../../pnnp_b200/csrc/pack_core.cuh"

using namespace pnnp;

extern "C" {

// Philox block exactly as the kernels form it: RngCtx::block(index, stream, sub) under (seed, offset)
void emul_philox_blocks(uint64_t seed, uint64_t offset, const uint64_t* index, uint32_t stream, uint32_t sub, int n, uint32_t* out4) {
    const PhiloxKeys rk = philox_round_keys(seed);
    const RngCtx rng{rk, (uint32_t)offset, (uint32_t)(offset >> 32)};
    for (int i = 0; i < n; ++i) {
        const uint4 b = rng.block(index[i], stream, sub);
        out4[4 * i] = b.x; out4[4 * i + 1] = b.y; out4[4 * i + 2] = b.z; out4[4 * i + 3] = b.w;
    }
}

void emul_normal_icdf(const uint32_t* w, int n, float* out) { for (int i = 0; i < n; ++i) out[i] = normal_icdf(w[i]); }

void emul_tukey_lambda(const uint32_t* w, float lam, int n, float* out) {
    const float inv = lam != 0.f ? 1.0f / lam : 0.f;
    for (int i = 0; i < n; ++i) out[i] = tukey_lambda_ppf(w[i], lam, inv);
}

// T: the 161 x 37 float table the launcher builds (noise_synth.cu: ensure_poisson_table)
void emul_poisson(const float* lam, const uint32_t* w, const float* T, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = poisson_sample(lam[i], w[i], T);
}

void emul_quant_draws(const uint32_t* mix, int n, double* q64, float* q32) {
    for (int i = 0; i < n; ++i) { q64[i] = quant_draw_f64(mix[i]); q32[i] = quant_draw_f32(mix[i]); }
}

// noise_replay_kernel (noise_synth.cu) for ONE crop of c x h x w: the deterministic tail on caller-supplied draws
void emul_replay(const float* clean, float* noisy, const pnnp_noise_params* row, int c, int h, int w, uint32_t code, int chain,
                 int ori, int clip, float post_lo, float post_hi, const float* d_shot, const float* d_read, const float* d_rowz,
                 const double* d_q) {
    const RowP p = load_row_params(row);
    for (int ch = 0; ch < c; ++ch)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                const size_t i = ((size_t)ch * h + y) * w + x;
                const float shot = d_shot ? d_shot[i] : 0.f, read = d_read ? d_read[i] : 0.f;
                const float rowz = d_rowz ? d_rowz[(size_t)ch * h + y] : 0.f;
                const double dq = d_q ? d_q[i] : 0.0;
                float out;
                if (chain == PNNP_CHAIN_NUMPY) {
                    const double bias_c = (code & PNNP_CODE_D) ? row->bias[ch & 3] : 0.0;
                    out = tail_numpy(clean[i], p, code, ori != 0, clip != 0, shot, read, rowz, dq, bias_c);
                } else {
                    out = tail_torch(p, code, ori != 0, clip != 0, shot, read, rowz, (float)dq);
                }
                noisy[i] = fminf(fmaxf(out, post_lo), post_hi);
            }
}

// the specialised kernel's per-element arithmetic (noise_synth.cu: noise_synth_fast_kernel phases 1 and 3) on supplied draws:
// Markstein divisions instead of IEEE divisions; must equal emul_replay bit for bit for 'pgrq' with float64 K / sigR
void emul_fast_tail(const float* clean, float* noisy, float* rate_out, const pnnp_noise_params* t, int n, const float* cnt,
                    const float* d_read, float rowz, const double* d_q, float post_lo, float post_hi) {
    const FastC f = fast_constants(t, post_lo, post_hi);
    const double row64 = __dmul_rn((double)rowz, f.sigR);
    for (int i = 0; i < n; ++i) {
        rate_out[i] = fast_rate(f, clean[i]);
        noisy[i] = fast_tail(f, (int)cnt[i], d_read[i], row64, d_q[i]);
    }
}

void emul_div_by_const_f32(const float* a, const float* b, int n, float* out) {
    for (int i = 0; i < n; ++i) out[i] = div_rn_by_const(a[i], b[i], __frcp_rn(b[i]));
}
void emul_div_by_const_f64(const double* a, const double* b, int n, double* out) {
    for (int i = 0; i < n; ++i) out[i] = div_rn_by_const(a[i], b[i], __drcp_rn(b[i]));
}

// pack.cu's per-sample arithmetic over arrays (one black level per call: the kernels pick black[c] by plane)
void emul_norm_one(const float* v, int n, double black, double wp, int norm, int clip, float* out) {
    for (int i = 0; i < n; ++i) out[i] = norm_one(v[i], black, wp, norm, clip);
}
void emul_quant_one(const float* v, int n, float span, float bl, uint16_t* out) {
    for (int i = 0; i < n; ++i) out[i] = (uint16_t)quant_one(v[i], span, bl);
}

}  // extern "C"
