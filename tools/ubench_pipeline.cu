// Micro-benchmarks of the producer <-> consumer hand-shake that bounds the small-K conv layers: how many cycles does
// one trip through a full/empty mbarrier ring cost on a B200 SM, per signalling scheme?  Developer tool for round 2 — build with
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/_bin/ubench_pipeline tools/ubench_pipeline.cu
// and run under gpurun:  tools/_bin/ubench_pipeline  > gpurun_out/ubench_pipeline.txt
//
// One CTA = warp 0 (producer) + warp 1 (consumer) + `idle_warps` warps parked at the final barrier.  Ring of S stages.
//   producer: wait empty[s] -> signal full[s]          consumer: wait full[s] -> signal empty[s]
// Schemes (producer signal / consumer signal):
//   0  warp-uniform loop, try_wait, elect.sync + mbarrier.arrive, __syncwarp            (what conv_tc.cu does with MMAs off)
//   1  as 0 with test_wait (pure polling)
//   2  one thread per role (lane 0 only), try_wait + arrive, no elect / syncwarp
//   3  as 0, consumer signals with tcgen05.commit (nothing outstanding: measures the commit -> mbarrier path)
//   4  as 0, producer "loads": cp.async.bulk global -> shared of `bytes` with complete_tx on full[s]   (TMA latency / rate)
//   5  as 4 with the consumer signalling through tcgen05.commit
// Output: cycles per iteration for S = 1 (latency) .. 8 (throughput), one CTA per SM and two.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
template <bool TEST>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, it = 0;
    while (true) {
        if (TEST) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        else asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
        if (++it > (1u << 24)) return false;            // broken protocol: give up instead of hanging the GPU
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\tselp.b32 %0, 1, 0, px;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kMaxStages = 8;

template <int SCHEME>
__global__ void __launch_bounds__(1024) pingpong(int stages, int iters, int bytes, const uint8_t* src, size_t src_bytes, long long* cycles, int* failed) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr bool kLoad = SCHEME == 4 || SCHEME == 5, kCommit = SCHEME == 3 || SCHEME == 5, kTest = SCHEME == 1, kSingle = SCHEME == 2;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) { mbar_init(smem_u32(&full_bar[s]), 1); mbar_init(smem_u32(&empty_bar[s]), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t full0 = smem_u32(full_bar), empty0 = smem_u32(empty_bar), buf0 = smem_u32(smem);
    bool ok = true;
    if (warp == 0 && (!kSingle || lane == 0)) {
        uint32_t stage = 0, phase = 0;
        const uint8_t* base = src + (size_t)blockIdx.x * (size_t)bytes * 64 % (src_bytes - (size_t)bytes * 64);
        for (int i = 0; i < iters && ok; ++i) {
            ok = mbar_wait<kTest>(empty0 + 8 * stage, phase ^ 1);
            if (kSingle) mbar_arrive(full0 + 8 * stage);
            else {
                if (elect_one()) {
                    if (kLoad) {
                        mbar_expect_tx(full0 + 8 * stage, (uint32_t)bytes);
                        bulk_load(buf0 + stage * (uint32_t)bytes, base + (size_t)(i & 63) * bytes, (uint32_t)bytes, full0 + 8 * stage);
                    } else mbar_arrive(full0 + 8 * stage);
                }
                __syncwarp();
            }
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == 1 && (!kSingle || lane == 0)) {
        uint32_t stage = 0, phase = 0;
        const long long t0 = clock64();
        for (int i = 0; i < iters && ok; ++i) {
            ok = mbar_wait<kTest>(full0 + 8 * stage, phase);
            if (kSingle) mbar_arrive(empty0 + 8 * stage);
            else {
                if (elect_one()) { if (kCommit) tc_commit(empty0 + 8 * stage); else mbar_arrive(empty0 + 8 * stage); }
                __syncwarp();
            }
            if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
        }
        const long long t1 = clock64();
        if (lane == 0) cycles[blockIdx.x] = t1 - t0;
    }
    if (!ok) atomicExch(failed, 1);
    __syncthreads();
}

template <int SCHEME>
static void run(const char* name, int ctas_per_sm, int idle_warps, int bytes, const uint8_t* src, size_t src_bytes, long long* d_cycles, int* d_failed, int sms) {
    const int iters = 20000;
    const int grid = sms * ctas_per_sm;
    const int threads = 64 + 32 * idle_warps;
    printf("%-58s ctas/SM %d warps %2d bytes %5d :", name, ctas_per_sm, threads / 32, bytes);
    for (int stages : {1, 2, 4, 8}) {
        const size_t smem = (size_t)stages * (bytes ? bytes : 16) + 1024;
        if (smem * ctas_per_sm > 220 * 1024) { printf("  S=%d n/a", stages); continue; }
        CK(cudaFuncSetAttribute(pingpong<SCHEME>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(d_failed, 0, sizeof(int)));
        pingpong<SCHEME><<<grid, threads, smem>>>(stages, 200, bytes, src, src_bytes, d_cycles, d_failed);      // warm-up
        pingpong<SCHEME><<<grid, threads, smem>>>(stages, iters, bytes, src, src_bytes, d_cycles, d_failed);
        CK(cudaDeviceSynchronize());
        static long long h[4096];
        int failed = 0;
        CK(cudaMemcpy(h, d_cycles, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&failed, d_failed, sizeof(int), cudaMemcpyDeviceToHost));
        double sum = 0, mx = 0;
        for (int i = 0; i < grid; ++i) { sum += (double)h[i]; if ((double)h[i] > mx) mx = (double)h[i]; }
        printf("  S=%d %6.0f (max %6.0f)%s", stages, sum / grid / iters, mx / iters, failed ? " FAILED" : "");
    }
    printf("   cycles/iteration\n");
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    long long* d_cycles; int* d_failed; uint8_t* src;
    const size_t src_bytes = (size_t)512 << 20;                         // larger than L2: the bulk loads come from HBM
    CK(cudaMalloc(&d_cycles, sizeof(long long) * 4096));
    CK(cudaMalloc(&d_failed, sizeof(int)));
    CK(cudaMalloc(&src, src_bytes));
    CK(cudaMemset(src, 1, src_bytes));
    printf("SMs %d\n", sms);
    for (int cps : {1, 2}) {
        for (int idle : {0, 8}) {
            run<0>("0 try_wait + elect + arrive (conv_tc.cu, MMAs off)", cps, idle, 0, src, src_bytes, d_cycles, d_failed, sms);
            run<1>("1 test_wait + elect + arrive", cps, idle, 0, src, src_bytes, d_cycles, d_failed, sms);
            run<2>("2 single thread per role, try_wait + arrive", cps, idle, 0, src, src_bytes, d_cycles, d_failed, sms);
        }
        for (int bytes : {5120, 10240, 20480, 36864})
            run<4>("4 producer = cp.async.bulk load, consumer arrive", cps, 8, bytes, src, src_bytes, d_cycles, d_failed, sms);
    }
    fflush(stdout);
    // the tcgen05.commit schemes last: if committing with nothing outstanding (and no TMEM allocation) faults, the rest is on file
    for (int cps : {1, 2}) {
        run<3>("3 consumer signals with tcgen05.commit", cps, 8, 0, src, src_bytes, d_cycles, d_failed, sms);
        fflush(stdout);
        for (int bytes : {5120, 20480})
            run<5>("5 producer = cp.async.bulk load, consumer tcgen05.commit", cps, 8, bytes, src, src_bytes, d_cycles, d_failed, sms);
    }
    return 0;
}
