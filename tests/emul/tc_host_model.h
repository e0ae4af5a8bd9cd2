// FUNCTIONAL host model of the tcgen05 / TMA / mbarrier wrappers of pnnp_b200/csrc/tc_common.cuh, plus the few CUDA driver /
// runtime names the conv launcher touches — TEST INFRASTRUCTURE ONLY (tests/test_device_tc_on_cpu.py), never part of the product.
// Include after cuda_host_shim.h and simt_host.h.  With it the CPU suite compiles conv_tc.cu ITSELF for the host — kernel, launcher,
// tensor-map set-up — and runs it on the fibre SIMT emulator.
//
// What is modelled (cta_group::1, kind::f16 — all the conv and weight-gradient kernels use):
//   * shared memory = pnnp::smem_raw (1024-aligned), shared "addresses" = offset + kSmemBase;
//   * mbarrier: phase bit, pending arrivals, outstanding transaction bytes (init / arrive / arrive.expect_tx / complete_tx /
//     try_wait.parity); a failed wait yields to the other fibres of the CTA;
//   * TMA tiled loads (3-D / 4-D, element strides, zero fill outside the tensor, 32 / 64 / 128-byte swizzle = XOR of address bits
//     [4,7) with bits [7,10) limited to the swizzle span) completing bytes on an mbarrier;
//   * tcgen05.mma: D[M x N] (+)= A[M x 16] * B[N x 16]^T in fp32 from the two shared-memory descriptors (start, LBO, SBO, swizzle
//     mode; K-major and MN-major canonical layouts), M = 128 rows -> TMEM lanes; tcgen05.commit arrives after the MMAs queued before it;
//   * TMEM = 128 lanes x 512 columns; tcgen05.ld 32x32b.x16 with the warp-quadrant lane restriction checked.
// The model is calibrated by the kernels that are parity-green on a B200: under it the default instantiations reproduce torch's
// convolutions; the same semantics then check the opt-in instantiations before they are given GPU time.  TMA loads, MMAs and
// commits are asynchronous here too — queued at issue, executed as late as possible, TMA destinations poisoned meanwhile — so a
// consumer that does not wait for the announcing barrier is caught; PNNP_EMUL_ASYNC picks which agent lags (fifo / tma_first /
// tc_first).  Not modelled: timing, and alignment faults beyond the checks below.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>

#define __grid_constant__

// ------------------------------------------------------------------------------------------ driver / runtime names
typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0;
typedef int CUresult;
constexpr CUresult CUDA_SUCCESS = 0;
typedef uint32_t cuuint32_t;
typedef uint64_t cuuint64_t;
enum CUtensorMapDataType { CU_TENSOR_MAP_DATA_TYPE_FLOAT32 = 7, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 = 9 };
enum CUtensorMapInterleave { CU_TENSOR_MAP_INTERLEAVE_NONE = 0 };
enum CUtensorMapSwizzle { CU_TENSOR_MAP_SWIZZLE_NONE = 0, CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_SWIZZLE_128B };
enum CUtensorMapL2promotion { CU_TENSOR_MAP_L2_PROMOTION_NONE = 0, CU_TENSOR_MAP_L2_PROMOTION_L2_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B };
enum CUtensorMapFloatOOBfill { CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE = 0 };
struct CUtensorMap {            // what cuTensorMapEncodeTiled was told (the real one is an opaque 128-byte blob)
    const uint8_t* base; int rank, elem_bytes, swizzle_bytes;
    uint64_t dims[5], strides[5];            // strides in bytes; strides[0] = element size
    uint32_t box[5], estride[5];
};
static inline CUresult emul_cuTensorMapEncodeTiled(CUtensorMap* tm, CUtensorMapDataType dt, cuuint32_t rank, void* ptr, const cuuint64_t* dims,
                                                   const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* es, CUtensorMapInterleave il,
                                                   CUtensorMapSwizzle sw, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    if (dt != CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 || rank < 3 || rank > 5 || il != CU_TENSOR_MAP_INTERLEAVE_NONE) return 1;
    if (reinterpret_cast<uintptr_t>(ptr) & 15) return 1;                                    // globalAddress: 16-byte aligned
    tm->base = static_cast<const uint8_t*>(ptr); tm->rank = (int)rank; tm->elem_bytes = 2;
    tm->swizzle_bytes = sw == CU_TENSOR_MAP_SWIZZLE_128B ? 128 : (sw == CU_TENSOR_MAP_SWIZZLE_64B ? 64 : (sw == CU_TENSOR_MAP_SWIZZLE_32B ? 32 : 0));
    for (uint32_t d = 0; d < rank; ++d) {
        tm->dims[d] = dims[d]; tm->box[d] = box[d]; tm->estride[d] = es[d];
        tm->strides[d] = d == 0 ? 2 : strides[d - 1];
        if (box[d] < 1 || box[d] > 256 || es[d] < 1 || es[d] > 8 || dims[d] < 1) return 1;
        if (d > 0 && (strides[d - 1] & 15)) return 1;                                       // globalStrides: multiples of 16 bytes
    }
    if (tm->swizzle_bytes && box[0] * 2u > (uint32_t)tm->swizzle_bytes) return 1;           // inner box extent <= swizzle span
    return 0;
}
enum cudaDriverEntryPointQueryResult { cudaDriverEntryPointSuccess = 0 };
constexpr int cudaEnableDefault = 0;
static inline cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, int, cudaDriverEntryPointQueryResult* q) {
    *fn = std::strcmp(name, "cuTensorMapEncodeTiled") ? nullptr : reinterpret_cast<void*>(&emul_cuTensorMapEncodeTiled);
    *q = cudaDriverEntryPointSuccess;
    return *fn ? cudaSuccess : 1;
}
static inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {             // a small "GPU": the grid, not the SM count, is what a test varies
    *v = std::getenv("PNNP_EMUL_SMS") ? std::atoi(std::getenv("PNNP_EMUL_SMS")) : 3;
    return cudaSuccess;
}
template <typename T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = static_cast<T*>(std::malloc(n)); return *p ? cudaSuccess : 2; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
enum cudaMemcpyKind { cudaMemcpyDeviceToHost = 2 };
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int bytes) { return bytes <= 227 * 1024 ? cudaSuccess : 1; }
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
enum cudaLaunchAttributeID { cudaLaunchAttributeProgrammaticStreamSerialization = 4 };
struct cudaLaunchAttribute { cudaLaunchAttributeID id; struct { int programmaticStreamSerializationAllowed; } val; };
struct cudaLaunchConfig_t { dim3 gridDim, blockDim; size_t dynamicSmemBytes = 0; cudaStream_t stream = nullptr; cudaLaunchAttribute* attrs = nullptr; unsigned numAttrs = 0; };

namespace pnnp {

constexpr uint32_t kSpinLimit = 1u << 27;
constexpr uint32_t kSmemBase = 0x4000;                       // shared "address" of smem_raw[0]
constexpr size_t kSmemBytes = 227 * 1024;
extern __attribute__((aligned(1024))) uint8_t smem_raw[];   // defined by the including file
static uint32_t g_tmem[128][512];
static unsigned long g_stalled_polls = 0;                    // consecutive failed mbarrier polls of the running CTA (deadlock detector)
struct TcStats { unsigned long mma = 0, tma = 0, tma_oob_elems = 0, waits = 0; };
static TcStats g_tc_stats;
// Asynchronous agents (TMA unit, tensor pipe): an operation is queued at issue and EXECUTED AS LATE AS POSSIBLE — one queued operation
// per failed mbarrier poll, in issue order — so device code that consumes a result without waiting for the barrier that announces it
// (shared-memory stage, accumulator, reused slot) reads poison / stale data here instead of passing by luck of issue order.
static std::deque<std::function<void()>> g_async_tma, g_async_tc;     // per agent, each in issue order (commits ride the tensor-pipe queue)
static std::deque<int> g_async_order;                                  // issue order across both agents: 0 = TMA, 1 = tensor pipe
static unsigned long g_async_work = 0;                      // queued TMA loads / MMAs (commits are not counted)
static unsigned long g_trailing_commits = 0;                // commits that were still queued when their CTA exited (no work before them)
// Which queued operation happens next when a fibre's barrier poll fails — PNNP_EMUL_ASYNC = fifo (default: issue order across both
// agents), tma_first (the TMA unit runs ahead, the tensor pipe lags as far as the barriers allow) or tc_first (the reverse).
static inline int async_policy() {
    const char* e = std::getenv("PNNP_EMUL_ASYNC");          // read per operation: a test switches it between launches
    return !e ? 0 : (!std::strcmp(e, "tma_first") ? 1 : (!std::strcmp(e, "tc_first") ? 2 : 0));
}
static inline void async_push(int agent, std::function<void()> op) { (agent ? g_async_tc : g_async_tma).push_back(std::move(op)); g_async_order.push_back(agent); }
static inline bool async_run_one() {
    if (g_async_order.empty()) return false;
    int agent = g_async_order.front();
    const int pol = async_policy();
    if (pol == 1 && !g_async_tma.empty()) agent = 0;
    if (pol == 2 && !g_async_tc.empty()) agent = 1;
    for (auto it = g_async_order.begin(); it != g_async_order.end(); ++it) if (*it == agent) { g_async_order.erase(it); break; }
    auto& q = agent ? g_async_tc : g_async_tma;
    std::function<void()> op = std::move(q.front());
    q.pop_front();
    op();
    return true;
}
static inline void tc_model_fail(const char* what) { std::fprintf(stderr, "tc_host_model: %s\n", what); std::abort(); }
static inline uint8_t* smem_ptr(uint32_t addr, size_t bytes) {
    if (addr < kSmemBase || (size_t)(addr - kSmemBase) + bytes > kSmemBytes) tc_model_fail("shared-memory address out of range");
    return smem_raw + (addr - kSmemBase);
}
static inline uint32_t smem_u32(const void* p) {
    const ptrdiff_t off = static_cast<const uint8_t*>(p) - smem_raw;
    if (off < 0 || (size_t)off >= kSmemBytes) tc_model_fail("smem_u32 of a pointer outside the CTA's shared memory");
    return kSmemBase + (uint32_t)off;
}
static inline uint32_t swizzle_addr(uint32_t addr, int span) {          // cute Swizzle<B,4,3>, B = log2(span / 16)
    const uint32_t mask = span == 128 ? 7u : (span == 64 ? 3u : (span == 32 ? 1u : 0u));
    return addr ^ (((addr >> 7) & mask) << 4);
}

// ---- mbarrier (the 8 bytes in shared memory hold the model's state)
struct MbarState { uint32_t phase : 1; uint32_t expected : 15; uint32_t pending : 16; int32_t tx; };
static_assert(sizeof(MbarState) == 8, "mbarrier state must fit the 64-bit barrier object");
static inline MbarState* mbar_at(uint32_t bar) { if (bar & 7) tc_model_fail("mbarrier not 8-byte aligned"); return reinterpret_cast<MbarState*>(smem_ptr(bar, 8)); }
static inline void mbar_settle(MbarState* b) { if (b->pending == 0 && b->tx == 0) { b->phase ^= 1; b->pending = b->expected; g_stalled_polls = 0; } }
static inline void mbar_init(uint32_t bar, uint32_t count) { MbarState* b = mbar_at(bar); b->phase = 0; b->expected = count; b->pending = count; b->tx = 0; }
static inline void mbar_arrive(uint32_t bar) { MbarState* b = mbar_at(bar); if (!b->pending) tc_model_fail("mbarrier arrive beyond its count"); --b->pending; mbar_settle(b); }
static inline void mbar_expect_tx(uint32_t bar, uint32_t bytes) { MbarState* b = mbar_at(bar); b->tx += (int32_t)bytes; if (!b->pending) tc_model_fail("mbarrier arrive beyond its count"); --b->pending; mbar_settle(b); }
static inline void mbar_complete_tx(uint32_t bar, uint32_t bytes) { MbarState* b = mbar_at(bar); b->tx -= (int32_t)bytes; mbar_settle(b); }
static inline void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
    ++g_tc_stats.waits;
    while (true) {
        if ((mbar_at(bar)->phase & 1u) != (parity & 1u)) return;              // the phase with this parity has completed
        if (*err != 0) return;
        if (async_run_one()) { g_stalled_polls = 0; continue; }               // let the oldest queued TMA / MMA / commit happen, poll again
        if (++g_stalled_polls > 400000ul) { *err = code; return; }            // every fibre parked, nothing in flight: the device would time out here
        const unsigned tid = simt::g_cta->cur;
        simt::yield();
        threadIdx.x = tid;
    }
}
static inline bool elect_one() { return (simt::g_cta->cur & 31u) == 0; }

// ---- TMA tiled load: box of the tensor map at element coordinates c[], written row-major (dim 0 fastest) with the map's swizzle
static inline void tma_load(uint32_t dst, const CUtensorMap* tm, uint32_t bar, const int* c) {
    ++g_tc_stats.tma;
    if (dst & 127) tc_model_fail("TMA shared-memory destination must be 128-byte aligned");
    if (tm->swizzle_bytes && (dst % (8u * (uint32_t)tm->swizzle_bytes))) tc_model_fail("swizzled TMA destination not aligned to its 8-row atom");
    uint32_t cnt[5] = {1, 1, 1, 1, 1};
    size_t total = 1;
    for (int d = 0; d < tm->rank; ++d) { cnt[d] = (tm->box[d] + tm->estride[d] - 1) / tm->estride[d]; total *= cnt[d]; }
    const uint16_t poison = 0x7FC0;                                           // bf16 NaN: the unit may write the box any time from now on
    for (size_t lin = 0; lin < total; ++lin) std::memcpy(smem_ptr(swizzle_addr(dst + (uint32_t)(lin * 2), tm->swizzle_bytes), 2), &poison, 2);
    const CUtensorMap map = *tm;
    int cc[5];
    for (int d = 0; d < 5; ++d) cc[d] = c[d];
    ++g_async_work;
    async_push(0, [=]() {
        --g_async_work;
        size_t lin = 0;
        for (uint32_t i4 = 0; i4 < cnt[4]; ++i4) for (uint32_t i3 = 0; i3 < cnt[3]; ++i3) for (uint32_t i2 = 0; i2 < cnt[2]; ++i2)
        for (uint32_t i1 = 0; i1 < cnt[1]; ++i1) for (uint32_t i0 = 0; i0 < cnt[0]; ++i0, ++lin) {
            const uint32_t idx[5] = {i0, i1, i2, i3, i4};
            bool inside = true;
            size_t goff = 0;
            for (int d = 0; d < map.rank; ++d) {
                const long long g = (long long)cc[d] + (long long)idx[d] * map.estride[d];
                if (g < 0 || g >= (long long)map.dims[d]) { inside = false; break; }
                goff += (size_t)g * map.strides[d];
            }
            uint16_t v = 0;
            if (inside) std::memcpy(&v, map.base + goff, 2); else ++g_tc_stats.tma_oob_elems;
            std::memcpy(smem_ptr(swizzle_addr(dst + (uint32_t)(lin * 2), map.swizzle_bytes), 2), &v, 2);
        }
        mbar_complete_tx(bar, (uint32_t)(total * 2));
    });
}
static inline void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    if (tm->rank != 4) tc_model_fail("tma_load_4d on a tensor map of another rank");
    const int c[5] = {c0, c1, c2, c3, 0};
    tma_load(dst, tm, bar, c);
}
static inline void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2) {
    if (tm->rank != 3) tc_model_fail("tma_load_3d on a tensor map of another rank");
    const int c[5] = {c0, c1, c2, 0, 0};
    tma_load(dst, tm, bar, c);
}

// ---- tcgen05
static inline void tc_fence_before() {}
static inline void tc_fence_after() {}
static inline void tc_ld_wait() {}
static inline void prefetch_tensormap(const CUtensorMap*) {}
static inline void fence_mbarrier_init() {}
static inline void fence_proxy_async() {}
static inline void pdl_wait_then_release() {}
static inline void tmem_alloc(uint32_t slot, uint32_t cols) {
    if (cols < 32 || cols > 512 || (cols & (cols - 1))) tc_model_fail("tcgen05.alloc: column count must be a power of two in [32, 512]");
    const uint32_t base = 0;
    std::memcpy(smem_ptr(slot, 4), &base, 4);
    if ((simt::g_cta->cur & 31u) == 0) for (auto& lane : g_tmem) for (auto& w : lane) w = 0x7FC00000u;     // unwritten accumulators read as NaN
}
static inline void tmem_dealloc(uint32_t, uint32_t) {}
static inline void tc_commit(uint32_t bar) { async_push(1, [=]() { mbar_arrive(bar); }); }     // arrives once every MMA queued before it has executed
static inline float bf16_at(uint32_t addr) { uint16_t h; std::memcpy(&h, smem_ptr(addr, 2), 2); return __uint_as_float((uint32_t)h << 16); }
static inline void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    ++g_tc_stats.mma;
    const int M = (int)((idesc >> 24) & 0x1Fu) << 4, N = (int)((idesc >> 17) & 0x3Fu) << 3;
    if (M != 128) tc_model_fail("tcgen05.mma: only M = 128 (cta_group::1) is modelled");
    if (N < 16 || N > 256 || (N & 15)) tc_model_fail("tcgen05.mma: N must be a multiple of 16 in [16, 256] for M = 128");
    if ((idesc & ((1u << 4) | (1u << 7) | (1u << 10))) != ((1u << 4) | (1u << 7) | (1u << 10))) tc_model_fail("tcgen05.mma: expected f32 accumulate, bf16 A and B");
    const bool a_mn = (idesc >> 15) & 1u, b_mn = (idesc >> 16) & 1u;          // operand major-ness: 0 = K-major, 1 = MN-major
    struct Op { uint32_t start, sbo, lbo; int span; bool mn; };
    auto decode = [](uint64_t desc, bool mn) {
        Op o;
        o.start = (uint32_t)(desc & 0x3FFFu) << 4; o.lbo = (uint32_t)((desc >> 16) & 0x3FFFu) << 4; o.sbo = (uint32_t)((desc >> 32) & 0x3FFFu) << 4;
        o.mn = mn;
        const unsigned layout = (unsigned)(desc >> 61);
        o.span = layout == 2 ? 128 : (layout == 4 ? 64 : (layout == 6 ? 32 : 0));
        if (!o.span) tc_model_fail("shared-memory descriptor: only the 32 / 64 / 128-byte swizzled layouts are modelled");
        // K-major operands: the hardware applies the swizzle to the ABSOLUTE address start + (row / 8) * SBO + (row % 8) * span + 2 k
        // (measured on a B200, tools/ubench_umma_offset.cu, r02: start offsets of whole rows and SBO = 10 rows read exactly these rows
        // with the base-offset field 0), so any whole-row start and any whole-row SBO is modelled; MN-major operands keep the atom checks.
        if (mn ? o.sbo != 8u * (uint32_t)o.span : (o.sbo % (uint32_t)o.span) != 0 || o.sbo < 8u * (uint32_t)o.span)
            tc_model_fail("shared-memory descriptor: SBO must be 8 rows of one swizzle span (MN-major) / a whole number >= 8 of rows (K-major)");
        if (mn ? (o.start % (8u * (uint32_t)o.span)) != 0 : (o.start % (uint32_t)o.span) + 32u > (uint32_t)o.span)
            tc_model_fail("shared-memory descriptor: start address not at the head of a swizzle atom (MN-major) / K slice outside its row (K-major)");
        if (mn && (o.lbo % (8u * (uint32_t)o.span))) tc_model_fail("shared-memory descriptor: MN-major LBO must be a whole number of swizzle atoms");
        return o;
    };
    // element (i = M or N index, k = 0..15) of an operand.  K-major: 8-row atoms of one span per row, k contiguous.  MN-major (cute
    // canonical ((T,span/16,m),(8,k)) : ((1,T,LBO),(span/2,SBO)) in elements): span/2 consecutive MN elements per K row, further MN
    // blocks LBO bytes apart, 8 K rows per atom, atoms SBO bytes apart.
    auto addr = [](const Op& o, int i, int k) {
        const uint32_t e = (uint32_t)o.span / 2;
        const uint32_t a = o.mn ? o.start + ((uint32_t)i % e) * 2 + ((uint32_t)i / e) * o.lbo + (uint32_t)(k & 7) * (uint32_t)o.span + (uint32_t)(k >> 3) * o.sbo
                                : o.start + (uint32_t)(i >> 3) * o.sbo + (uint32_t)(i & 7) * (uint32_t)o.span + (uint32_t)k * 2;
        return swizzle_addr(a, o.span);
    };
    const Op oa = decode(adesc, a_mn), ob = decode(bdesc, b_mn);
    const uint32_t col0 = d_tmem & 0xFFFFu;
    if ((d_tmem >> 16) != 0 || col0 + (uint32_t)N > 512) tc_model_fail("tcgen05.mma: accumulator outside TMEM");
    ++g_async_work;
    async_push(1, [=]() {
        --g_async_work;
        static float a[128][16], b[256][16];
        for (int m = 0; m < M; ++m) for (int k = 0; k < 16; ++k) a[m][k] = bf16_at(addr(oa, m, k));
        for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) b[n][k] = bf16_at(addr(ob, n, k));
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            float s = 0.f;
            for (int k = 0; k < 16; ++k) s += a[m][k] * b[n][k];
            uint32_t& d = g_tmem[m][col0 + n];
            d = __float_as_uint(accumulate ? __uint_as_float(d) + s : s);
        }
    });
}
static inline void tc_mma_tf32(uint32_t, uint64_t, uint64_t, uint32_t, uint32_t) { tc_model_fail("tcgen05.mma.kind::tf32 is not modelled (the fp32-storage variant is checked on the GPU only)"); }
static inline void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    const unsigned tid = simt::g_cta->cur, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lane0 = taddr >> 16, col = taddr & 0xFFFFu;
    if (lane0 != 32u * (warp & 3u)) tc_model_fail("tcgen05.ld: a warp may only read the TMEM lane quadrant of its warp id % 4");
    if (col + 16 > 512) tc_model_fail("tcgen05.ld: column out of range");
    for (int i = 0; i < 16; ++i) v[i] = g_tmem[lane0 + lane][col + i];
}

static inline void st_global_256(void* p, const uint32_t (&v)[8]) { std::memcpy(p, v, 32); }
static inline void ld_global_256(const void* p, uint32_t (&v)[8]) { std::memcpy(v, p, 32); }

// packed fp32 pairs
static inline uint64_t f2_pack(float a, float b) { return (uint64_t)__float_as_uint(a) | ((uint64_t)__float_as_uint(b) << 32); }
static inline void f2_unpack(uint64_t v, float& a, float& b) { a = __uint_as_float((uint32_t)v); b = __uint_as_float((uint32_t)(v >> 32)); }
static inline uint64_t f2_add(uint64_t x, uint64_t y) { float a, b, c, d; f2_unpack(x, a, b); f2_unpack(y, c, d); return f2_pack(a + c, b + d); }
static inline uint64_t f2_fma(uint64_t x, uint64_t y, uint64_t z) { float a, b, c, d, e, f; f2_unpack(x, a, b); f2_unpack(y, c, d); f2_unpack(z, e, f); return f2_pack(fmaf(a, c, e), fmaf(b, d, f)); }

static inline uint64_t umma_desc_hi(int swz) {
    const uint64_t layout = swz == 128 ? 2ull : (swz == 64 ? 4ull : 6ull);
    const uint64_t sbo = (uint64_t)((8 * swz) >> 4);
    return (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
static inline uint64_t umma_desc(uint64_t hi, uint32_t saddr) { return hi | (uint64_t)((saddr >> 4) & 0x3FFFu); }

// launch on the SIMT emulator: CTAs one after the other
static inline void emul_launch_1d(int grid, int threads, const std::function<void()>& body) {
    if (grid < 1 || threads < 32 || threads > 1024) tc_model_fail("kernel launch with an invalid grid / block size");
    if (std::getenv("PNNP_EMUL_PLAN_ONLY")) return;          // launcher dry run: variant selection, smem / TMEM plan, tensor maps — no compute
    gridDim.x = (unsigned)grid; blockDim.x = (unsigned)threads;
    for (int b = 0; b < grid; ++b) {
        blockIdx.x = (unsigned)b; g_stalled_polls = 0;
        simt::run_cta((unsigned)threads, body);
        if (g_async_work) tc_model_fail("CTA exited with TMA loads / MMAs still in flight (nobody waited for them)");
        while (!g_async_order.empty()) { ++g_trailing_commits; async_run_one(); }    // a commit with nothing before it: completes at once on the device
    }
}

}  // namespace pnnp

template <typename... KArgs, typename... Args>
static inline cudaError_t cudaLaunchKernelEx(const cudaLaunchConfig_t* cfg, void (*kernel)(KArgs...), Args&&... args) {
    pnnp::emul_launch_1d((int)cfg->gridDim.x, (int)cfg->blockDim.x, [&]() { kernel(args...); });
    return cudaSuccess;
}
