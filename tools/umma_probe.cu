// Hardware probe for the tcgen05 shared-memory descriptor conventions the conv kernels rely on:
//   * no-swizzle K-major operands whose start address is shifted by a multiple of 16 B (tap shift of a flattened halo)
//   * no-swizzle MN-major operands (wgrad: K = pixel index) with the same shift trick
//   * 128B-swizzle K-major (fallback layout)
//   * M=64 accumulator layout in TMEM
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu
// Run  : ./umma_probe <test-id>     (one test per process so a fault cannot poison the next one)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../pytorch_nested-unet_b200/csrc/ptx.cuh"

using namespace nunet;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

struct ProbeArgs {
    const uint8_t* a_img; uint32_t a_bytes;
    const uint8_t* b_img; uint32_t b_bytes;
    uint32_t mode;  // 0 = no swizzle, 1 = sw128
    uint32_t a_start, a_lbo, a_sbo, a_kstep;
    uint32_t b_start, b_lbo, b_sbo, b_kstep;
    uint32_t idesc, nk, ncols, col_off;
    float* out;  // [128][ncols]
    uint32_t* status;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 1023u) & ~1023u);
    for (uint32_t i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16) *(uint4*)(sa + i) = *(const uint4*)(p.a_img + i);
    for (uint32_t i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16) *(uint4*)(sb + i) = *(const uint4*)(p.b_img + i);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tb = tmem_base;
    if (threadIdx.x == 0) {
        for (uint32_t k = 0; k < p.nk; ++k) {
            uint32_t aa = smem_u32(sa) + p.a_start + k * p.a_kstep;
            uint32_t ba = smem_u32(sb) + p.b_start + k * p.b_kstep;
            uint64_t ad = p.mode == 0 ? umma_desc_nosw(aa, p.a_lbo, p.a_sbo) : umma_desc_sw128(aa, p.a_sbo);
            uint64_t bd = p.mode == 0 ? umma_desc_nosw(ba, p.b_lbo, p.b_sbo) : umma_desc_sw128(ba, p.b_sbo);
            umma_bf16(tb + p.col_off, ad, bd, p.idesc, k > 0);
        }
        umma_commit(&bar);
    }
    __syncwarp();
    // bounded wait that reports instead of trapping
    uint32_t spins = 0; bool ok = true;
    while (!mbar_try_wait(&bar, 0)) { if (++spins > (1u << 22)) { ok = false; break; } }
    if (!ok) { if (threadIdx.x == 0) p.status[0] = 1; }
    tc_fence_after();
    if (ok) {
        uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (uint32_t c = 0; c < p.ncols; c += 16) {
            uint32_t v[16];
            tmem_ld16(tb + ((warp * 32u) << 16) + p.col_off + c, v);
            tmem_ld_wait();
            for (int j = 0; j < 16; ++j) p.out[(warp * 32 + lane) * p.ncols + c + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

__global__ void __launch_bounds__(128, 1) rate_kernel(ProbeArgs p, int reps, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    // operands: whatever bytes are in smem (zero-filled); only timing matters
    for (uint32_t i = threadIdx.x * 16; i < p.a_bytes + p.b_bytes + 2048; i += blockDim.x * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tb = tmem_base;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 1023u) & ~1023u);
    if (threadIdx.x == 0) {
        const uint32_t a_lo0 = p.mode == 0 ? umma_desc_lo(smem_u32(sa), p.a_lbo) : (uint32_t)(umma_desc_sw128(smem_u32(sa), p.a_sbo) & 0xffffffffu);
        const uint32_t b_lo0 = p.mode == 0 ? umma_desc_lo(smem_u32(sb), p.b_lbo) : (uint32_t)(umma_desc_sw128(smem_u32(sb), p.b_sbo) & 0xffffffffu);
        const uint32_t a_hi = p.mode == 0 ? umma_desc_hi(p.a_sbo) : (uint32_t)(umma_desc_sw128(smem_u32(sa), p.a_sbo) >> 32);
        const uint32_t b_hi = p.mode == 0 ? umma_desc_hi(p.b_sbo) : (uint32_t)(umma_desc_sw128(smem_u32(sb), p.b_sbo) >> 32);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k)
                umma_bf16_lohi(tb + (k & 1) * p.ncols, a_lo0 + (k & 3) * (p.a_kstep >> 4), a_hi, b_lo0 + (k & 3) * (p.b_kstep >> 4), b_hi, p.idesc, 1u);
        }
        umma_commit(&bar);
        while (!mbar_try_wait(&bar, 0)) {}
        long long t1 = clock64();
        cycles[0] = t1 - t0;
    }
    __syncthreads();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

int run_rate_test(const char* name, int M, int N, int amaj, int bmaj) {
    // layouts as in run_gemm_test with nk = 4 K-steps available to cycle through
    ProbeArgs p{};
    const int K = 64;
    if (amaj == 0) { p.a_lbo = M * 16; p.a_sbo = 128; p.a_kstep = 2 * p.a_lbo; p.a_bytes = M * K * 2; }
    else if (amaj == 1) { p.a_lbo = 128; p.a_sbo = K * 16; p.a_kstep = 256; p.a_bytes = M * K * 2; }
    else { p.a_lbo = 16; p.a_sbo = 1024; p.a_kstep = 32; p.a_bytes = M * 128; }
    if (bmaj == 0) { p.b_lbo = N * 16; p.b_sbo = 128; p.b_kstep = 2 * p.b_lbo; p.b_bytes = N * K * 2; }
    else if (bmaj == 1) { p.b_lbo = 128; p.b_sbo = K * 16; p.b_kstep = 256; p.b_bytes = N * K * 2; }
    else { p.b_lbo = 16; p.b_sbo = 1024; p.b_kstep = 32; p.b_bytes = N * 128; }
    p.mode = (amaj == 2) ? 1 : 0;
    p.idesc = umma_idesc_bf16(M, N, amaj == 1 ? kMajorMN : kMajorK, bmaj == 1 ? kMajorMN : kMajorK);
    p.ncols = N;
    long long* dc; CK(cudaMalloc(&dc, 8));
    uint32_t smem_bytes = ((p.a_bytes + 1023u) & ~1023u) + p.b_bytes + 4096;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int reps = 2000;
    for (int it = 0; it < 2; ++it) { rate_kernel<<<1, 128, smem_bytes>>>(p, reps, dc); CK(cudaDeviceSynchronize()); }
    long long c; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
    double per = (double)c / (reps * 8.0);
    printf("[rate %s] M=%d N=%d: %.1f cycles per MMA (K=16)  -> %.0f MAC/cycle/SM (peak 4096)\n", name, M, N, per, (double)M * N * 16 / per);
    cudaFree(dc);
    return 0;
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }  // exact for small ints
static float rnd_small(uint32_t& s) { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 24) % 5) - 2); }


struct Layout {  // byte offset of element (row, k) inside an operand image
    int major;    // 0 K-major nosw, 1 MN-major nosw, 2 sw128 K-major
    int rows_total, k_total;
    uint32_t off(int r, int k) const {
        if (major == 0) return (uint32_t)((k / 8) * rows_total * 16 + r * 16 + (k % 8) * 2);     // [kgroup][row][8]
        if (major == 1) return (uint32_t)((r / 8) * k_total * 16 + k * 16 + (r % 8) * 2);        // [rowgroup][k][8]
        return (uint32_t)(r * 128 + ((((k / 8) ^ (r % 8)) & 7) * 16) + (k % 8) * 2);             // sw128, k_total == 64
    }
    uint32_t bytes() const { return major == 2 ? rows_total * 128 : (uint32_t)rows_total * k_total * 2; }
};

int run_gemm_test(const char* name, int M, int N, int nk, int amaj, int bmaj, int shiftA, int shiftB, int padA, int padB, int col_off) {
    // "rows" = MN index, "k" = reduction index. Shift applies along the dimension that is 16B-pitched in smem:
    //   K-major: rows (pixels are rows);  MN-major: k (pixels are k).
    int K = 16 * nk;
    Layout la, lb;
    la.major = amaj; lb.major = bmaj;
    if (amaj == 0) { la.rows_total = M + shiftA + padA; la.k_total = K; }
    else if (amaj == 1) { la.rows_total = M; la.k_total = K + shiftA + padA; }
    else { la.rows_total = M; la.k_total = 64; }
    if (bmaj == 0) { lb.rows_total = N + shiftB + padB; lb.k_total = K; }
    else if (bmaj == 1) { lb.rows_total = N; lb.k_total = K + shiftB + padB; }
    else { lb.rows_total = N; lb.k_total = 64; }
    std::vector<uint8_t> ia(la.bytes() + 16, 0), ib(lb.bytes() + 16, 0);
    std::vector<float> A((size_t)la.rows_total * la.k_total), B((size_t)lb.rows_total * lb.k_total);
    uint32_t seed = 12345u + M * 7 + N * 13 + nk;
    for (int r = 0; r < la.rows_total; ++r) for (int k = 0; k < la.k_total; ++k) {
        float v = rnd_small(seed); A[(size_t)r * la.k_total + k] = v; uint16_t h = f2bf(v); memcpy(&ia[la.off(r, k)], &h, 2); }
    for (int r = 0; r < lb.rows_total; ++r) for (int k = 0; k < lb.k_total; ++k) {
        float v = rnd_small(seed); B[(size_t)r * lb.k_total + k] = v; uint16_t h = f2bf(v); memcpy(&ib[lb.off(r, k)], &h, 2); }
    ProbeArgs p{};
    p.a_bytes = (la.bytes() + 15) & ~15u; p.b_bytes = (lb.bytes() + 15) & ~15u;
    uint8_t *da, *db; float* dout; uint32_t* dst;
    CK(cudaMalloc(&da, p.a_bytes)); CK(cudaMalloc(&db, p.b_bytes));
    CK(cudaMalloc(&dout, 128 * N * 4)); CK(cudaMalloc(&dst, 4));
    CK(cudaMemcpy(da, ia.data(), p.a_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, ib.data(), p.b_bytes, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, 128 * N * 4)); CK(cudaMemset(dst, 0, 4));
    p.a_img = da; p.b_img = db; p.out = dout; p.status = dst;
    p.mode = (amaj == 2) ? 1 : 0;
    if (amaj == 0) { p.a_start = shiftA * 16; p.a_lbo = la.rows_total * 16; p.a_sbo = 128; p.a_kstep = 2 * p.a_lbo; }
    else if (amaj == 1) { p.a_start = shiftA * 16; p.a_lbo = 128; p.a_sbo = la.k_total * 16; p.a_kstep = 256; }
    else { p.a_start = 0; p.a_lbo = 16; p.a_sbo = 1024; p.a_kstep = 32; }
    if (bmaj == 0) { p.b_start = shiftB * 16; p.b_lbo = lb.rows_total * 16; p.b_sbo = 128; p.b_kstep = 2 * p.b_lbo; }
    else if (bmaj == 1) { p.b_start = shiftB * 16; p.b_lbo = 128; p.b_sbo = lb.k_total * 16; p.b_kstep = 256; }
    else { p.b_start = 0; p.b_lbo = 16; p.b_sbo = 1024; p.b_kstep = 32; }
    p.idesc = umma_idesc_bf16(M, N, amaj == 1 ? kMajorMN : kMajorK, bmaj == 1 ? kMajorMN : kMajorK);
    p.nk = nk; p.ncols = N; p.col_off = col_off;
    uint32_t smem_bytes = ((p.a_bytes + 1023u) & ~1023u) + p.b_bytes + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    probe_kernel<<<1, 128, smem_bytes>>>(p);
    CK(cudaDeviceSynchronize());
    std::vector<float> out(128 * N); uint32_t st;
    CK(cudaMemcpy(out.data(), dout, 128 * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&st, dst, 4, cudaMemcpyDeviceToHost));
    if (st) { printf("[%s] TIMEOUT waiting for MMA commit\n", name); return 1; }
    // expected: D[m][n] = sum_k A(m (+shift), k (+shift)) * B(n, k)
    double maxdiff = 0; int bad = 0;
    std::vector<float> ref((size_t)M * N);
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        float acc = 0;
        for (int k = 0; k < K; ++k) {
            float a = amaj == 0 ? A[(size_t)(m + shiftA) * la.k_total + k] : amaj == 1 ? A[(size_t)m * la.k_total + k + shiftA] : A[(size_t)m * la.k_total + k];
            float b = bmaj == 0 ? B[(size_t)(n + shiftB) * lb.k_total + k] : bmaj == 1 ? B[(size_t)n * lb.k_total + k + shiftB] : B[(size_t)n * lb.k_total + k];
            acc += a * b;
        }
        ref[(size_t)m * N + n] = acc;
    }
    if (M == 128) {
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double d = fabs(out[m * N + n] - ref[(size_t)m * N + n]);
            if (d > maxdiff) maxdiff = d;
            if (d > 0 && bad < 5) { printf("   mismatch m=%d n=%d got %g want %g\n", m, n, out[m * N + n], ref[(size_t)m * N + n]); ++bad; }
        }
        printf("[%s] M=%d N=%d K=%d maxdiff=%g  %s\n", name, M, N, K, maxdiff, maxdiff == 0 ? "PASS" : "FAIL");
    } else {
        // M=64: try candidate lane maps
        const char* names[3] = {"lane=m", "lane=m+ (m/32)*32 (rows 0-31 -> lanes 0-31, 32-63 -> lanes 64-95)", "lane=(m%16)+(m/16)*32"};
        for (int c = 0; c < 3; ++c) {
            double md = 0;
            for (int m = 0; m < 64; ++m) for (int n = 0; n < N; ++n) {
                int lane = c == 0 ? m : c == 1 ? (m % 32) + (m / 32) * 64 : (m % 16) + (m / 16) * 32;
                double d = fabs(out[lane * N + n] - ref[(size_t)m * N + n]); if (d > md) md = d;
            }
            printf("[%s] M=64 candidate '%s': maxdiff=%g %s\n", name, names[c], md, md == 0 ? "PASS" : "");
        }
    }
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dst);
    return maxdiff == 0 ? 0 : 1;
}

int main(int argc, char** argv) {
    int t = argc > 1 ? atoi(argv[1]) : 0;
    switch (t) {
        case 0: return run_gemm_test("nosw KK aligned", 128, 64, 4, 0, 0, 0, 0, 0, 0, 0);
        case 1: return run_gemm_test("nosw KK shiftA=1", 128, 64, 4, 0, 0, 1, 0, 0, 0, 0);
        case 2: return run_gemm_test("nosw KK shiftA=11 odd pitch", 128, 64, 4, 0, 0, 11, 0, 14, 1, 0);
        case 3: return run_gemm_test("nosw KK N=32 col_off=96", 128, 32, 4, 0, 0, 3, 0, 6, 0, 96);
        case 4: return run_gemm_test("nosw KK N=256", 128, 256, 2, 0, 0, 5, 0, 4, 0, 0);
        case 5: return run_gemm_test("nosw MNMN aligned", 128, 64, 4, 1, 1, 0, 0, 0, 0, 0);
        case 6: return run_gemm_test("nosw MNMN shift 11/11 odd pitch", 128, 64, 4, 1, 1, 11, 11, 14, 14, 0);
        case 7: return run_gemm_test("nosw MNMN shiftA=0 shiftB=29 N=96", 128, 96, 4, 1, 1, 0, 29, 9, 4, 0);
        case 8: return run_gemm_test("sw128 KK", 128, 64, 4, 2, 2, 0, 0, 0, 0, 0);
        case 9: return run_gemm_test("nosw KK M=64", 64, 32, 2, 0, 0, 0, 0, 0, 0, 0);
        case 10: return run_gemm_test("nosw A=MN B=K", 128, 64, 4, 1, 0, 7, 0, 2, 0, 0);
        case 11: return run_gemm_test("nosw A=K B=MN", 128, 64, 4, 0, 1, 7, 3, 2, 6, 0);
        case 12: return run_gemm_test("nosw MNMN M=64", 64, 64, 2, 1, 1, 5, 5, 4, 4, 0);
        case 13: return run_gemm_test("nosw KK N=48", 128, 48, 2, 0, 0, 2, 0, 7, 0, 0);
        case 100: run_rate_test("nosw KK", 128, 32, 0, 0); run_rate_test("nosw KK", 128, 64, 0, 0); run_rate_test("nosw KK", 128, 128, 0, 0);
                  return run_rate_test("nosw KK", 128, 256, 0, 0);
        case 101: run_rate_test("sw128 KK", 128, 32, 2, 2); run_rate_test("sw128 KK", 128, 64, 2, 2); run_rate_test("sw128 KK", 128, 128, 2, 2);
                  return run_rate_test("sw128 KK", 128, 256, 2, 2);
        case 102: run_rate_test("nosw MNMN", 128, 32, 1, 1); run_rate_test("nosw MNMN", 128, 64, 1, 1); run_rate_test("nosw MNMN", 128, 128, 1, 1);
                  return run_rate_test("nosw MNMN", 128, 256, 1, 1);
        case 103: run_rate_test("nosw A=K B=MN", 128, 32, 0, 1); return run_rate_test("nosw A=MN B=K", 128, 32, 1, 0);
    }
    return 0;
}
