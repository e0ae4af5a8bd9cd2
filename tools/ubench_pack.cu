// Micro-benchmark of the pack kernel's vector path (pnnp_b200/csrc/pack_kernels.cuh) on 64 crops of 1024 x 1024 uint16 codes; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DPACK_TAG='"cs"' [-DPNNP_PACK_ST=__stcg -DPNNP_PACK_LD=__ldcg] -o tools/_bin/ubench_pack_cs tools/ubench_pack.cu;
// built several times with different load / store flavours (-DPNNP_PACK_ST=..., -DPNNP_PACK_LD=...), run with <blocks per SM> <threads>.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ void st_plain(float4* p, float4 v) { *p = v; }       // -DPNNP_PACK_ST=st_plain -DPNNP_PACK_LD=ld_plain
__device__ __forceinline__ uint4 ld_plain(const uint4* p) { return *p; }
__device__ __forceinline__ float4 ld_plain(const float4* p) { return *p; }
#include "../pnnp_b200/csrc/pack_kernels.cuh"
using namespace pnnp;
int main(int argc, char** argv) {
    const int n = 64, H = 1024, W = 1024;
    const int bps = argc > 1 ? atoi(argv[1]) : 8, threads = argc > 2 ? atoi(argv[2]) : 256;
    uint16_t* raw; float* out;
    cudaMalloc(&raw, (size_t)n * H * W * 2);
    cudaMalloc(&out, (size_t)n * H * W * 4);
    cudaMemset(raw, 1, (size_t)n * H * W * 2);
    const double black[4] = {512, 512, 512, 512};
    const PackArgs a = make_pack_args(raw, out, n, H, W, 16383.0, black, 1, 1);
    const size_t items = (size_t)n * (H / 2) * (W / 8);
    int grid = bps > 0 ? 148 * bps : (int)((items + threads - 1) / threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int r = 0; r < 6; ++r) {
        cudaEventRecord(e0);
        for (int k = 0; k < 10; ++k) pack_norm_kernel<uint16_t, true, true><<<grid, threads>>>(a);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (r && ms / 10 < best) best = ms / 10;
    }
    printf("%s grid %d x %d: %.1f us, %.0f GB/s (%s)\n", PACK_TAG, grid, threads, best * 1e3, (double)n * H * W * 6 / best / 1e6, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
