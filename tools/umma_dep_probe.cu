// Does a chain of tcgen05.mma instructions that accumulate into the SAME TMEM tile run at the pipe's throughput or at its latency?
// Issues 8 x reps MMAs (128 x N x 16, no-swizzle K-major operands) cycling over `nacc` distinct accumulator tiles.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/umma_dep_probe tools/umma_dep_probe.cu && ./tools/umma_dep_probe
#include <cstdio>
#include "../pytorch_nested-unet_b200/csrc/ptx.cuh"
using namespace nunet;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(128, 1) dep_kernel(int N, int nacc, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (uint32_t i = threadIdx.x * 16; i < 128 * 64 * 2 + 256 * 64 * 2 + 2048; i += blockDim.x * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = tmem_base;
    if (threadIdx.x == 0) {
        const uint32_t a_lbo = 128 * 16, b_lbo = N * 16;
        const uint32_t a_lo0 = umma_desc_lo(smem_u32(smem), a_lbo), b_lo0 = umma_desc_lo(smem_u32(smem + 128 * 64 * 2), b_lbo);
        const uint32_t hi = umma_desc_hi(128u);
        const uint32_t idesc = umma_idesc_bf16(128, N, kMajorK, kMajorK);
        long long t0 = clock64();
        int acc = 0;
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                umma_bf16_lohi(tb + acc * N, a_lo0 + (k & 3) * ((2 * a_lbo) >> 4), hi, b_lo0 + (k & 3) * ((2 * b_lbo) >> 4), hi, idesc, 1u);
                acc = (acc + 1 == nacc) ? 0 : acc + 1;
            }
        }
        umma_commit(&bar);
        while (!mbar_try_wait(&bar, 0)) {}
        cycles[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

int main() {
    long long* dc; CK(cudaMalloc(&dc, 8));
    const int smem_bytes = 128 * 64 * 2 + 256 * 64 * 2 + 4096;
    CK(cudaFuncSetAttribute(dep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int Ns[] = {32, 64, 96, 128, 192};
    for (int N : Ns)
        for (int nacc = 1; nacc <= 5 && nacc * N <= 512; ++nacc) {
            const int reps = 2000;
            long long c = 0;
            for (int it = 0; it < 2; ++it) { dep_kernel<<<1, 128, smem_bytes>>>(N, nacc, reps, dc); CK(cudaDeviceSynchronize()); }
            CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
            printf("[dep] M=128 N=%3d K=16, %d accumulator tile(s) in rotation: %.1f cycles per MMA\n", N, nacc, (double)c / (reps * 8.0));
        }
    return 0;
}
