// Experiment (developer tool, not part of the library): does tcgen05.mma accept a K-major SWIZZLE_64B A operand whose start address
// is NOT aligned to an 8-row core-matrix group, and an SBO that is not 8 rows?  If it does, a 3x3 conv can take all nine taps from
// ONE haloed TMA box (tile 8 wide: a tile row = one 8-row group, SBO = box row pitch, tap = start offset (dy * 10 + dx) rows).
// A[r][k] holds distinct small integers, B = identity, so D[m][n] tells which physical row / chunk the tensor core read for row m.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/_bin/ubench_umma_offset tools/ubench_umma_offset.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../pnnp_b200/csrc/tc_common.cuh"
using namespace pnnp;

constexpr int kRows = 256;                       // physical rows of 64 bytes (32 bf16) in the A region
__device__ __host__ inline float aval(int r, int k) { return (float)((r * 7 + k * 3) % 251); }
__device__ inline uint32_t swz64(int r, int c) { return (uint32_t)(r * 64 + ((c ^ ((r >> 1) & 3)) << 4)); }

struct Test { int off_rows, sbo_bytes, base_offset; };

__global__ void k(const Test* tests, int n_tests, float* out, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sA = smem;                              // kRows x 64 B
    uint8_t* sB = smem + kRows * 64;                 // 32 x 64 B (16 KB offset: 1024-aligned)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 64);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < kRows * 32; i += blockDim.x) {
        const int r = i / 32, kk = i % 32;
        *reinterpret_cast<__nv_bfloat16*>(sA + swz64(r, kk >> 3) + (kk & 7) * 2) = __float2bfloat16_rn(aval(r, kk));
    }
    for (int i = tid; i < 32 * 32; i += blockDim.x) {
        const int r = i / 32, kk = i % 32;
        *reinterpret_cast<__nv_bfloat16*>(sB + swz64(r, kk >> 3) + (kk & 7) * 2) = __float2bfloat16_rn(r == kk ? 1.f : 0.f);
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); fence_mbarrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(slot), 32);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;
    for (int t = 0; t < n_tests; ++t) {
        const Test ts = tests[t];
        if (tid == 0) {
            const uint64_t layout = 4ull;            // SWIZZLE_64B
            const uint64_t hi_b = (1ull << 16) | ((uint64_t)(512 >> 4) << 32) | (1ull << 46) | (layout << 61);
            const uint64_t hi_a = (1ull << 16) | ((uint64_t)(ts.sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)(ts.base_offset & 7) << 49) | (layout << 61);
            const uint32_t a0 = smem_u32(sA) + ts.off_rows * 64, b0 = smem_u32(sB);
            tc_fence_after();
            for (int ks = 0; ks < 2; ++ks)
                tc_mma_bf16(tmem, hi_a | (uint64_t)(((a0 + ks * 32) >> 4) & 0x3FFFu), hi_b | (uint64_t)(((b0 + ks * 32) >> 4) & 0x3FFFu), idesc, ks ? 1u : 0u);
            tc_commit(smem_u32(bar));
        }
        mbar_wait(smem_u32(bar), phase, err, 900 + t);
        phase ^= 1;
        tc_fence_after();
        uint32_t v[16];
        for (int j = 0; j < 2; ++j) {
            tc_ld16(tmem + ((uint32_t)(warp * 32) << 16) + j * 16, v);
            tc_ld_wait();
            for (int i = 0; i < 16; ++i) out[((size_t)t * 128 + tid) * 32 + j * 16 + i] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 32); }
}

int main() {
    const Test tests[] = {{0, 512, 0}, {1, 512, 0}, {1, 512, 1}, {2, 512, 0}, {2, 512, 1}, {3, 512, 0}, {4, 512, 0}, {4, 512, 2}, {0, 640, 0}, {11, 640, 0},
                          {11, 640, 5}, {10, 640, 0}, {10, 640, 5}, {1, 640, 0}, {21, 640, 0}, {22, 640, 0}, {8, 640, 0}};
    const int nt = sizeof(tests) / sizeof(tests[0]);
    Test* d_t; float* d_o; int* d_e;
    cudaMalloc(&d_t, sizeof(tests)); cudaMalloc(&d_o, sizeof(float) * nt * 128 * 32); cudaMalloc(&d_e, 4);
    cudaMemcpy(d_t, tests, sizeof(tests), cudaMemcpyHostToDevice); cudaMemset(d_e, 0, 4);
    const int smem = 1024 + kRows * 64 + 32 * 64 + 64;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k<<<1, 128, smem>>>(d_t, nt, d_o, d_e);
    cudaError_t e = cudaDeviceSynchronize();
    int herr = 0; cudaMemcpy(&herr, d_e, 4, cudaMemcpyDeviceToHost);
    printf("sync: %s, pipeline error word %d\n", cudaGetErrorString(e), herr);
    static float o[32 * 128 * 32];
    cudaMemcpy(o, d_o, sizeof(float) * nt * 128 * 32, cudaMemcpyDeviceToHost);
    for (int t = 0; t < nt; ++t) {
        const int G = tests[t].sbo_bytes / 64;
        int bad = 0, bad_rows = 0;
        for (int m = 0; m < 128; ++m) {
            const int r = tests[t].off_rows + (m / 8) * G + m % 8;
            int rb = 0;
            for (int n = 0; n < 32; ++n) if (o[((size_t)t * 128 + m) * 32 + n] != aval(r, n)) { ++bad; rb = 1; }
            bad_rows += rb;
        }
        printf("test %2d: start +%2d rows, SBO %3d B, base_offset %d -> %4d of 4096 elements differ from the linear-address model (%3d rows)", t,
               tests[t].off_rows, tests[t].sbo_bytes, tests[t].base_offset, bad, bad_rows);
        // which row did the hardware read for m = 0, 1, 7, 8, 9 (decode from n = 0: value = (r * 7) % 251, search r)
        printf("   rows read for m = 0,1,7,8,9,127:");
        const int ms[6] = {0, 1, 7, 8, 9, 127};
        for (int q = 0; q < 6; ++q) {
            int found = -1;
            for (int r = 0; r < kRows && found < 0; ++r) {
                bool ok = true;
                for (int n = 0; n < 8 && ok; ++n) ok = o[((size_t)t * 128 + ms[q]) * 32 + n] == aval(r, n);
                if (ok) found = r;
            }
            printf(" %d", found);
        }
        printf("\n");
    }
    return 0;
}
