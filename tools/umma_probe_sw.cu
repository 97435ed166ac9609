// Probe: swizzled (64B / 128B) operands whose start address is shifted by whole rows that are NOT multiples of the 8-row
// swizzle atom (tap shift of an implicit GEMM on TMA-written tiles).  Variants: descriptor base_offset = 0 vs (addr>>7)&7.
//   ./umma_probe_sw <test> <shift> <use_base_offset>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../pytorch_nested-unet_b200/csrc/ptx.cuh"
using namespace nunet;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct Args {
    const uint8_t *a_img, *b_img; uint32_t a_bytes, b_bytes;
    uint32_t a_start, a_lbo, a_sbo, a_kstep, a_layout;   // layout: 2 = SW128, 4 = SW64
    uint32_t b_start, b_lbo, b_sbo, b_kstep, b_layout;
    uint32_t idesc, nk, ncols, use_bo;
    float* out; uint32_t* status;
};
__device__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t use_bo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    if (use_bo) d |= (uint64_t)((addr >> 7) & 7) << 49;
    d |= (uint64_t)layout << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) k(Args p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar; __shared__ uint32_t tmem_base;
    uint8_t* sa = smem; uint8_t* sb = smem + ((p.a_bytes + 1023u) & ~1023u);
    for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16) *(uint4*)(sa + i) = *(const uint4*)(p.a_img + i);
    for (uint32_t i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16) *(uint4*)(sb + i) = *(const uint4*)(p.b_img + i);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    uint32_t tb = tmem_base;
    if (threadIdx.x == 0) {
        for (uint32_t kk = 0; kk < p.nk; ++kk) {
            uint64_t ad = mkdesc(smem_u32(sa) + p.a_start + kk * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout, p.use_bo);
            uint64_t bd = mkdesc(smem_u32(sb) + p.b_start + kk * p.b_kstep, p.b_lbo, p.b_sbo, p.b_layout, p.use_bo);
            umma_bf16(tb, ad, bd, p.idesc, kk > 0);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    uint32_t spins = 0; bool ok = true;
    while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { ok = false; break; } }
    if (!ok && threadIdx.x == 0) p.status[0] = 1;
    tc_fence_after();
    if (ok) {
        uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (uint32_t c = 0; c < p.ncols; c += 16) {
            uint32_t v[16]; tmem_ld16(tb + ((warp * 32u) << 16) + c, v); tmem_ld_wait();
            for (int j = 0; j < 16; ++j) p.out[(warp * 32 + lane) * p.ncols + c + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}
static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static float rnd(uint32_t& s) { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 24) % 5) - 2); }

// row-major tile [rows][rowbytes] with the hardware swizzle of absolute (1024B-aligned image) addresses
static uint32_t sw_off(uint32_t row, uint32_t kbyte, uint32_t rowbytes) {
    uint32_t addr = row * rowbytes + kbyte;
    if (rowbytes == 128) addr ^= ((addr >> 7) & 7) << 4;          // Swizzle<3,4,3>
    else if (rowbytes == 64) addr ^= ((addr >> 7) & 3) << 4;      // Swizzle<2,4,3>
    return addr;
}

int main(int argc, char** argv) {
    int test = argc > 1 ? atoi(argv[1]) : 1, shift = argc > 2 ? atoi(argv[2]) : 3, use_bo = argc > 3 ? atoi(argv[3]) : 0;
    Args p{}; p.use_bo = use_bo;
    const int M = 128;
    int N, K, rows_a, rows_b, rba, rbb;   // row bytes
    std::vector<float> A, B;              // logical A[m_total][K], B[n_total][K] (K-major tests) or [pos][ch] (MN tests)
    uint32_t seed = 777 + test * 13 + shift;
    std::vector<uint8_t> ia, ib;
    std::vector<float> ref;
    if (test == 1 || test == 2) {
        // K-major: rows = positions. test1: SW128 (64 ch per row), test2: SW64 (32 ch per row). A shifted by `shift` rows.
        rba = rbb = (test == 1) ? 128 : 64;
        K = rba / 2; N = 64; rows_a = M + shift + 5; rows_b = N;
        A.resize(rows_a * K); B.resize(rows_b * K);
        ia.assign(rows_a * rba + 1024, 0); ib.assign(rows_b * rbb + 1024, 0);
        for (int r = 0; r < rows_a; ++r) for (int c = 0; c < K; ++c) { float v = rnd(seed); A[r * K + c] = v; uint16_t h = f2bf(v); memcpy(&ia[sw_off(r, c * 2, rba)], &h, 2); }
        for (int r = 0; r < rows_b; ++r) for (int c = 0; c < K; ++c) { float v = rnd(seed); B[r * K + c] = v; uint16_t h = f2bf(v); memcpy(&ib[sw_off(r, c * 2, rbb)], &h, 2); }
        p.a_layout = p.b_layout = (test == 1) ? 2 : 4;
        p.a_start = shift * rba; p.a_lbo = 16; p.a_sbo = 8 * rba; p.a_kstep = 32;
        p.b_start = 0; p.b_lbo = 16; p.b_sbo = 8 * rbb; p.b_kstep = 32;
        p.nk = K / 16;
        p.idesc = umma_idesc_bf16(M, N, kMajorK, kMajorK);
        ref.resize(M * N);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int c = 0; c < K; ++c) s += A[(m + shift) * K + c] * B[n * K + c]; ref[m * N + n] = s; }
    } else {
        // MN-major (wgrad): K = positions (rows).  A = X [pos][128 ch] as two [pos][64] SW128 tiles (LBO apart), unshifted.
        // B = gy [pos][32 ch] SW64 tile, start shifted by `shift` rows.  D[ci][co] = sum_pos X[pos][ci] * gy[pos+shift][co]
        const int KP = 64;  // positions in the reduction
        rba = 128; rbb = 64; N = 32; K = KP;
        int pos_b = KP + shift + 5;
        std::vector<float> X(KP * 128), G(pos_b * 32);
        uint32_t half_bytes = KP * 128;
        ia.assign(2 * half_bytes + 1024, 0); ib.assign(pos_b * 64 + 1024, 0);
        for (int q = 0; q < KP; ++q) for (int c = 0; c < 128; ++c) { float v = rnd(seed); X[q * 128 + c] = v; uint16_t h = f2bf(v);
            memcpy(&ia[(c / 64) * half_bytes + sw_off(q, (c % 64) * 2, 128)], &h, 2); }
        for (int q = 0; q < pos_b; ++q) for (int c = 0; c < 32; ++c) { float v = rnd(seed); G[q * 32 + c] = v; uint16_t h = f2bf(v); memcpy(&ib[sw_off(q, c * 2, 64)], &h, 2); }
        p.a_layout = 2; p.b_layout = 4;
        p.a_start = 0; p.a_lbo = half_bytes; p.a_sbo = 1024; p.a_kstep = 16 * 128;   // 16 positions per MMA
        p.b_start = shift * 64; p.b_lbo = 16; p.b_sbo = 512; p.b_kstep = 16 * 64;
        p.nk = KP / 16;
        p.idesc = umma_idesc_bf16(M, N, kMajorMN, kMajorMN);
        ref.resize(M * N);
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int q = 0; q < KP; ++q) s += X[q * 128 + m] * G[(q + shift) * 32 + n]; ref[m * N + n] = s; }
    }
    p.a_bytes = (uint32_t)((ia.size() + 15) & ~15u) - 1008; p.b_bytes = (uint32_t)((ib.size() + 15) & ~15u) - 1008;
    p.ncols = N;
    uint8_t *da, *db; float* dout; uint32_t* dst;
    CK(cudaMalloc(&da, ia.size())); CK(cudaMalloc(&db, ib.size())); CK(cudaMalloc(&dout, 128 * N * 4)); CK(cudaMalloc(&dst, 4));
    CK(cudaMemcpy(da, ia.data(), ia.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, ib.data(), ib.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, 128 * N * 4)); CK(cudaMemset(dst, 0, 4));
    p.a_img = da; p.b_img = db; p.out = dout; p.status = dst;
    uint32_t smem_bytes = ((p.a_bytes + 1023u) & ~1023u) + p.b_bytes + 2048;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    k<<<1, 128, smem_bytes>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * N); uint32_t st;
    CK(cudaMemcpy(out.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
    if (st) { printf("[test %d shift %d bo %d] TIMEOUT\n", test, shift, use_bo); return 1; }
    double md = 0; int bad = 0;
    for (int i = 0; i < M * N; ++i) { double d = fabs(out[i] - ref[i]); if (d > md) md = d; if (d > 0) ++bad; }
    printf("[test %d shift %d base_offset %d] maxdiff %g mismatches %d/%d %s\n", test, shift, use_bo, md, bad, M * N, md == 0 ? "PASS" : "FAIL");
    return md == 0 ? 0 : 1;
}
