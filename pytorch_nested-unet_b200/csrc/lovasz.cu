// LovaszHingeLoss (losses.py:49-96,120-129; SURVEY.md 8f row f4 — the loss behind the README's published IoUs): per image, the hinge
// errors e = 1 - logit*sign are sorted in DESCENDING order, the Jaccard gradient of the sorted ground truth is formed from a running
// count (lovasz_grad, losses.py:49-60) and the loss is dot(relu(e_sorted), grad); the batch loss is the mean over images
// (lovasz_hinge with per_image=True, losses.py:62-75).  One CTA per image: a bitonic sort of (sortable key, pixel index) pairs — in
// shared memory up to 16 Ki elements, in a global scratch beyond —, a block-wide scan of the sorted labels, and the loss / gradient in
// one pass.  Ties are ordered by pixel index (the loss value does not depend on the order inside a tie group: the Jaccard increments
// telescope); the gradient w.r.t. each logit, -sign * grad * [e > 0], is written next to the loss for the backward pass.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nunet_b200.h"

namespace nunet {

constexpr int kLovThreads = 1024;
constexpr int kLovSmemElems = 16384;   // 128 KB of (key, index) pairs

__device__ __forceinline__ uint32_t desc_key(float e) {
    uint32_t u = __float_as_uint(e);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending sortable
    return ~u;                                          // ... descending in e when sorted ascending
}

// logits, target: [N][P] fp32 (target in {0,1}); coef: [N][P] fp32 out (d loss_image / d logit); loss_img: [N] fp64 out
__global__ void __launch_bounds__(kLovThreads, 1) lovasz_image_kernel(const float* __restrict__ logits, const float* __restrict__ target, int P,
                                                                      int Ppad, unsigned long long* __restrict__ scratch, float* __restrict__ coef,
                                                                      double* __restrict__ loss_img) {
    extern __shared__ __align__(16) unsigned long long s_pairs[];
    __shared__ int s_warp[32];
    __shared__ double s_red[32];
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* x = logits + (size_t)n * P;
    const float* t = target + (size_t)n * P;
    unsigned long long* a = (Ppad <= kLovSmemElems) ? s_pairs : scratch + (size_t)n * Ppad;
    for (int i = tid; i < Ppad; i += kLovThreads) {
        unsigned long long v = ~0ULL;   // padding sorts to the end
        if (i < P) {
            const float sign = 2.f * t[i] - 1.f;
            const float e = 1.f - x[i] * sign;
            v = ((unsigned long long)desc_key(e) << 32) | (unsigned)i;
        }
        a[i] = v;
    }
    __syncthreads();
    for (int k = 2; k <= Ppad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Ppad; i += kLovThreads) {
                const int l = i ^ j;
                if (l > i) {
                    const unsigned long long u = a[i], w = a[l];
                    const bool up = (i & k) == 0;
                    if ((u > w) == up) { a[i] = w; a[l] = u; }
                }
            }
            __syncthreads();
        }
    }
    // running count of the sorted labels: thread `tid` owns the contiguous chunk [tid*L, tid*L + L)
    const int L = (Ppad + kLovThreads - 1) / kLovThreads;
    const int i0 = tid * L, i1 = min(P, i0 + L);
    int cnt = 0;
    for (int i = i0; i < i1; ++i) cnt += (t[(unsigned)(a[i] & 0xFFFFFFFFu)] > 0.5f) ? 1 : 0;
    int inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += v;
    }
    if ((tid & 31) == 31) s_warp[tid >> 5] = inc;
    __syncthreads();
    if (tid < 32) {
        int v = s_warp[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, v, o);
            if (tid >= o) v += u;
        }
        s_warp[tid] = v;
    }
    __syncthreads();
    const int gts = s_warp[31];
    int cs = inc - cnt + ((tid >> 5) ? s_warp[(tid >> 5) - 1] : 0);   // labels before this chunk
    // jaccard_i = 1 - (gts - cs_i) / (gts + (i+1) - cs_i) with cs_i the inclusive count; grad_i = jaccard_i - jaccard_{i-1}
    auto jac = [&](int cs_incl, int i) { return 1.f - (float)(gts - cs_incl) / (float)(gts + (i + 1) - cs_incl); };
    float prev = (i0 > 0 && i0 <= P) ? jac(cs, i0 - 1) : 0.f;
    double acc = 0.0;
    for (int i = i0; i < i1; ++i) {
        const unsigned long long v = a[i];
        const unsigned idx = (unsigned)(v & 0xFFFFFFFFu);
        const float tv = t[idx];
        cs += (tv > 0.5f) ? 1 : 0;
        const float j = jac(cs, i);
        const float g = (i == 0) ? j : j - prev;
        prev = j;
        const float sign = 2.f * tv - 1.f;
        const float e = 1.f - x[idx] * sign;
        acc += (double)fmaxf(e, 0.f) * (double)g;
        coef[(size_t)n * P + idx] = (e > 0.f) ? -sign * g : 0.f;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((tid & 31) == 0) s_red[tid >> 5] = acc;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int w = 0; w < kLovThreads / 32; ++w) s += s_red[w];
        loss_img[n] = s;
    }
}

__global__ void lovasz_mean_kernel(const double* __restrict__ loss_img, int N, float* __restrict__ loss) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < N; ++i) s += loss_img[i];
        *loss = (float)(s / (double)N);     // mean() over the images (losses.py:28-47)
    }
}

// dlogits = coef * gscale[0] * gmul   (gmul = 1/N)
__global__ void lovasz_bwd_kernel(const float* __restrict__ coef, const float* __restrict__ gscale, float gmul, long long total, float* __restrict__ dlogits) {
    const float s = (gscale ? gscale[0] : 1.f) * gmul;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) dlogits[i] = coef[i] * s;
}

}  // namespace nunet

using namespace nunet;

static int lov_pad(int P) {
    int p = 2;
    while (p < P) p <<= 1;
    return p;
}

extern "C" long long nunet_lovasz_scratch_elems(int32_t N, int32_t P) {
    if (N < 1 || P < 1) return 0;
    const int pp = lov_pad(P);
    return pp <= kLovSmemElems ? 0 : (long long)N * pp;
}

extern "C" int nunet_lovasz_hinge_fwd(const float* logits, const float* target, int32_t N, int32_t P, unsigned long long* scratch, float* coef,
                                      double* loss_img, float* loss, void* stream) {
    if (!logits || !target || !coef || !loss_img || !loss || N < 1 || P < 1) return NUNET_E_BADARG;
    if (P > (1 << 24)) return NUNET_E_UNSUPPORTED;   // fp32 counts stay exact
    const int pp = lov_pad(P);
    const bool in_smem = pp <= kLovSmemElems;
    if (!in_smem && !scratch) return NUNET_E_BADARG;
    const size_t smem = in_smem ? (size_t)pp * 8 : 0;
    cudaError_t e = cudaFuncSetAttribute(lovasz_image_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kLovSmemElems * 8);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    lovasz_image_kernel<<<N, kLovThreads, smem, (cudaStream_t)stream>>>(logits, target, P, pp, scratch, coef, loss_img);
    lovasz_mean_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(loss_img, N, loss);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}

extern "C" int nunet_lovasz_hinge_bwd(const float* coef, const float* gscale, float gmul, long long total, float* dlogits, void* stream) {
    if (!coef || !dlogits || total < 1) return NUNET_E_BADARG;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    lovasz_bwd_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(coef, gscale, gmul, total, dlogits);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
