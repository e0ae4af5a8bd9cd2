// Lock-step SIMT emulator for the host — TEST INFRASTRUCTURE ONLY (tests/test_device_simt_on_cpu.py), never part of the product.
// Include after cuda_host_shim.h.  It runs ONE CTA at a time: every CUDA thread of the CTA is a fibre (ucontext) on the calling OS
// thread; a fibre runs until it reaches a collective (__syncthreads, __syncwarp, __ballot_sync, __shfl_sync, __shfl_xor_sync),
// parks there, and the scheduler moves on to the next fibre.  The last lane to arrive completes the collective and everyone
// resumes with the exchanged values — exactly the convergence the full-mask *_sync forms demand of the device code, so kernels
// that use warp ballots, shuffles, shared-memory queues and block barriers run unchanged.  Shared memory is whatever static
// storage the kernel source names (PNNP_SMEM / extern arrays defined by the including file); atomics are plain read-modify-writes
// (one fibre runs at a time).  Deterministic: fibres are resumed in thread order.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __shared__
#define __align__(n) __attribute__((aligned(n)))

namespace simt {

struct WarpState {
    uint64_t val[2][32];
    unsigned count = 0, gen = 0;
};
struct Cta {
    unsigned nthreads = 0, cur = 0;
    std::vector<ucontext_t> ctx;
    std::vector<char> done;
    std::vector<WarpState> warps;
    ucontext_t sched;
    unsigned bar_count = 0, bar_gen = 0;
    std::function<void()> body;
    unsigned long switches = 0;
};
static Cta* g_cta = nullptr;

static inline void yield() {
    Cta* c = g_cta;
    swapcontext(&c->ctx[c->cur], &c->sched);
}
static void trampoline() {
    Cta* c = g_cta;
    c->body();
    c->done[c->cur] = 1;
    swapcontext(&c->ctx[c->cur], &c->sched);
}

// Runs body() once per thread of a CTA of `nthreads` threads (blockIdx / gridDim / blockDim are set by the caller).
static inline void run_cta(unsigned nthreads, std::function<void()> body, size_t stack_bytes = 256 * 1024) {
    if (nthreads % 32) { std::fprintf(stderr, "simt: CTA size must be a multiple of 32\n"); std::abort(); }
    Cta c;
    c.nthreads = nthreads;
    c.ctx.resize(nthreads); c.done.assign(nthreads, 0); c.warps.resize(nthreads / 32);
    static std::vector<char*> pool;                      // fibre stacks, kept across CTAs (never zeroed, never freed)
    static size_t pool_bytes = 0;
    if (pool_bytes != stack_bytes) { for (char* q : pool) std::free(q); pool.clear(); pool_bytes = stack_bytes; }
    while (pool.size() < nthreads) pool.push_back(static_cast<char*>(std::malloc(stack_bytes)));
    c.body = std::move(body);
    g_cta = &c;
    for (unsigned t = 0; t < nthreads; ++t) {
        getcontext(&c.ctx[t]);
        c.ctx[t].uc_stack.ss_sp = pool[t];
        c.ctx[t].uc_stack.ss_size = stack_bytes;
        c.ctx[t].uc_link = &c.sched;
        makecontext(&c.ctx[t], trampoline, 0);
    }
    unsigned live = nthreads;
    while (live) {
        unsigned long before = c.switches;
        for (unsigned t = 0; t < nthreads; ++t) {
            if (c.done[t]) continue;
            c.cur = t;
            threadIdx.x = t;
            ++c.switches;
            swapcontext(&c.sched, &c.ctx[t]);
            if (c.done[t]) --live;
        }
        if (c.switches == before) break;
        if (c.switches > 4000000000ul) { std::fprintf(stderr, "simt: runaway kernel\n"); std::abort(); }
    }
    g_cta = nullptr;
}

// All 32 lanes of the calling fibre's warp exchange one 64-bit value; returns the generation's buffer.
static inline const uint64_t* warp_exchange(uint64_t mine) {
    Cta* c = g_cta;
    const unsigned tid = c->cur, lane = tid & 31;
    WarpState& w = c->warps[tid >> 5];
    const unsigned gen = w.gen;
    w.val[gen & 1][lane] = mine;
    if (++w.count == 32) { w.count = 0; ++w.gen; }
    else while (w.gen == gen) { yield(); threadIdx.x = tid; }
    return w.val[gen & 1];
}
static inline void cta_barrier() {
    Cta* c = g_cta;
    const unsigned tid = c->cur, gen = c->bar_gen;
    if (++c->bar_count == c->nthreads) { c->bar_count = 0; ++c->bar_gen; }
    else while (c->bar_gen == gen) { yield(); threadIdx.x = tid; }
}
template <typename T> static inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, ""); std::memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> static inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

}  // namespace simt

static inline void __syncthreads() { simt::cta_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { simt::warp_exchange(0); }
static inline unsigned __ballot_sync(unsigned mask, bool pred) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "simt: partial-mask ballot\n"); std::abort(); }
    const uint64_t* v = simt::warp_exchange(pred ? 1u : 0u);
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) m |= (unsigned)(v[l] & 1u) << l;
    return m;
}
template <typename T> static inline T __shfl_sync(unsigned mask, T val, int src) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "simt: partial-mask shuffle\n"); std::abort(); }
    return simt::from_bits<T>(simt::warp_exchange(simt::to_bits(val))[src & 31]);
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T val, int lane_mask) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "simt: partial-mask shuffle\n"); std::abort(); }
    const unsigned lane = simt::g_cta->cur & 31;
    return simt::from_bits<T>(simt::warp_exchange(simt::to_bits(val))[(lane ^ (unsigned)lane_mask) & 31]);
}
// width-limited forms: lanes are grouped in segments of `width`; a source outside the caller's segment returns the caller's own value
template <typename T> static inline T __shfl_up_sync(unsigned mask, T val, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "simt: partial-mask shuffle\n"); std::abort(); }
    const unsigned lane = simt::g_cta->cur & 31;
    const uint64_t* v = simt::warp_exchange(simt::to_bits(val));
    return (lane % (unsigned)width) >= delta ? simt::from_bits<T>(v[lane - delta]) : val;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T val, unsigned delta, int width = 32) {
    if (mask != 0xffffffffu) { std::fprintf(stderr, "simt: partial-mask shuffle\n"); std::abort(); }
    const unsigned lane = simt::g_cta->cur & 31;
    const uint64_t* v = simt::warp_exchange(simt::to_bits(val));
    return (lane % (unsigned)width) + delta < (unsigned)width ? simt::from_bits<T>(v[lane + delta]) : val;
}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
template <typename T> static inline T atomicAdd(T* p, T v) { const T old = *p; *p = old + v; return old; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }

// launch: CTAs one after the other
#define SIMT_LAUNCH(grid, block, call)                                                        \
    do {                                                                                      \
        gridDim.x = (grid); blockDim.x = (block);                                             \
        for (unsigned b_ = 0; b_ < (unsigned)(grid); ++b_) { blockIdx.x = b_; simt::run_cta((block), [&]() { call; }); } \
    } while (0)
#define SIMT_LAUNCH3(gx, gy, gz, block, call)                                                 \
    do {                                                                                      \
        gridDim.x = (gx); gridDim.y = (gy); gridDim.z = (gz); blockDim.x = (block);           \
        for (unsigned z_ = 0; z_ < (unsigned)(gz); ++z_)                                      \
            for (unsigned y_ = 0; y_ < (unsigned)(gy); ++y_)                                  \
                for (unsigned x_ = 0; x_ < (unsigned)(gx); ++x_) {                            \
                    blockIdx.x = x_; blockIdx.y = y_; blockIdx.z = z_;                        \
                    simt::run_cta((block), [&]() { call; });                                  \
                }                                                                             \
        gridDim.y = gridDim.z = 1; blockIdx.y = blockIdx.z = 0;                               \
    } while (0)
