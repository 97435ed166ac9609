// "Activate": everything the consumers of a raw conv output y (bf16 NHWC) need, produced in ONE pass over y:
//     a      = relu(scale*y + shift)                               nn.BatchNorm2d apply + ReLU      (finished/archs1.py:25-26,29-30)
//     a_pool = 2x2 max-pool of a                                   nn.MaxPool2d(2, 2)               (finished/archs1.py:82)
//     a_up   = bilinear 2x upsample of a, align_corners=True       nn.Upsample                      (finished/archs1.py:83)
//     logits = a (fp32, before the bf16 rounding) . Wh + bh        the 1x1 head of a level-0 node   (finished/archs1.py:106-111,135-142)
// Round 1 ran bn_apply, maxpool2x2, upsample2x and head_fwd as four launches that each re-read y or a; any output may be omitted
// (x4_0 / x3_1 / x2_2 / x1_3 are only ever consumed upsampled, x0_4 only through its head).
// Two kernels: a streaming one (no upsample; 2x2 windows per thread when pooling) and a banded one for the upsampled nodes: a CTA
// stages R+1 activated source rows in shared memory and emits its own rows of a / a_pool plus every output row of a_up whose upper
// source row lies in the band.  The upsample is separable: a thread owns one output column and 8 channels, keeps the horizontally
// blended values of the two source rows in registers and only blends vertically per output row (round 1's kernel re-blended four
// taps per output item and was instruction-bound: 167 instructions per 16-byte item, profiles/r02_ncu_kernels.csv).
#include "conv_common.cuh"
#include "loaders.cuh"

namespace nunet {

struct ActParams {
    const __nv_bfloat16* y;
    int N, H, W, C;
    const float *scale, *shift;
    __nv_bfloat16 *a, *a_pool, *a_up;
    const float *wh, *bh;
    int K;
    float* logits;
};

__device__ __forceinline__ void act8(const uint4& raw, const float* sc, const float* sh, float (&a)[8]) {
    float y[8];
    unpack8(raw, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(fmaf(y[i], sc[i], sh[i]), 0.f);
}

// ---- streaming: a (+ logits).  One item = (pixel, 8 channels); 4 items in flight per thread.
template <bool HEAD>
__global__ void __launch_bounds__(256) activate_stream_kernel(const ActParams p) {
    extern __shared__ float s_coef[];  // [2][C] (+ HEAD: [K][32] weights, [K] bias)
    const int C = p.C, cgs = C >> 3;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = p.scale[i]; s_coef[C + i] = p.shift[i]; }
    float* s_w = s_coef + 2 * C;
    if (HEAD) {
        for (int i = threadIdx.x; i < p.K * 32; i += blockDim.x) s_w[i] = p.wh[i];
        if ((int)threadIdx.x < p.K) s_w[p.K * 32 + threadIdx.x] = p.bh ? p.bh[threadIdx.x] : 0.f;
    }
    __syncthreads();
    const long long ngroups = (long long)p.N * p.H * p.W * cgs;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4* y = reinterpret_cast<const uint4*>(p.y);
    uint4* a = reinterpret_cast<uint4*>(p.a);
    const int HW = p.H * p.W;
    constexpr int U = 4;
    // HEAD: the four lanes of a pixel (C == 32) must stay together -> the loop bound is uniform per group of four lanes (ngroups % 4 == 0)
    for (long long g0 = blockIdx.x * (long long)blockDim.x + threadIdx.x; g0 < ngroups; g0 += U * stride) {
        uint4 r[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (g0 + u * stride < ngroups) r[u] = __ldg(y + g0 + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long gi = g0 + u * stride;
            const bool valid = gi < ngroups;     // no early exit: the head's shuffles need every lane of the warp
            if (!valid) r[u] = make_uint4(0, 0, 0, 0);
            const int c0 = (int)(gi % cgs) * 8;
            float v[8];
            act8(r[u], s_coef + c0, s_coef + C + c0, v);
            if (a && valid) a[gi] = pack8(v);
            if (HEAD) {
                const unsigned pix = (unsigned)(gi >> 2);   // C == 32: four 8-channel groups per pixel, adjacent lanes (N*H*W < 2^31, host-checked)
                const unsigned n = pix / (unsigned)HW, rem = pix - n * (unsigned)HW;
                for (int k = 0; k < p.K; ++k) {
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc = fmaf(v[i], s_w[k * 32 + c0 + i], acc);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                    if (c0 == 0 && valid) p.logits[((size_t)n * p.K + k) * HW + rem] = acc + s_w[p.K * 32 + k];
                }
            }
        }
    }
}

// ---- streaming with 2x2 max-pool: one item = (2x2 window, 8 channels): four loads, (four a stores,) one pooled store.
__global__ void __launch_bounds__(256) activate_pool_kernel(const ActParams p) {
    extern __shared__ float s_coef[];
    const int C = p.C, cgs = C >> 3;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = p.scale[i]; s_coef[C + i] = p.shift[i]; }
    __syncthreads();
    const int h = p.H >> 1, w = p.W >> 1;
    const long long items = (long long)p.N * h * w * cgs;
    const uint4* y = reinterpret_cast<const uint4*>(p.y);
    uint4* a = reinterpret_cast<uint4*>(p.a);
    uint4* ap = reinterpret_cast<uint4*>(p.a_pool);
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(it % cgs);
        long long pix = it / cgs;
        const int x = (int)(pix % w); pix /= w;
        const int yy = (int)(pix % h);
        const long long n = pix / h;
        const size_t base = ((size_t)(n * p.H + 2 * yy) * p.W + 2 * x) * cgs + cg;
        const size_t o01 = cgs, o10 = (size_t)p.W * cgs, o11 = o10 + cgs;
        const uint4 r0 = __ldg(y + base), r1 = __ldg(y + base + o01), r2 = __ldg(y + base + o10), r3 = __ldg(y + base + o11);
        float v0[8], v1[8], v2[8], v3[8], m[8];
        const float *sc = s_coef + cg * 8, *sh = s_coef + C + cg * 8;
        act8(r0, sc, sh, v0); act8(r1, sc, sh, v1); act8(r2, sc, sh, v2); act8(r3, sc, sh, v3);
        if (a) { a[base] = pack8(v0); a[base + o01] = pack8(v1); a[base + o10] = pack8(v2); a[base + o11] = pack8(v3); }
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = fmaxf(fmaxf(v0[i], v1[i]), fmaxf(v2[i], v3[i]));   // max commutes with the (monotone) bf16 rounding
        ap[it] = pack8(m);
    }
}

// ---- banded: a (optional), a_pool (optional), a_up.  W*C/8 is the same at every level below the first (channels double as the
// width halves), so a row is a few KB and a band of kBandRows+1 rows fits shared memory.
constexpr int kBandRows = 8;

__global__ void __launch_bounds__(256) activate_up_kernel(const ActParams p, int R, int bands) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    float* s_coef = reinterpret_cast<float*>(smem_raw);                 // [2][C]
    uint4* s_a = reinterpret_cast<uint4*>(smem_raw + (size_t)2 * p.C * 4);   // [R+1][W][C/8] activated rows (bf16)
    const int C = p.C, cgs = C >> 3, W = p.W, H = p.H;
    const int n = blockIdx.x / bands, y0 = (blockIdx.x - n * bands) * R;
    const int rows = min(R, H - y0);                 // rows owned by this band
    const int ld_rows = min(rows + 1, H - y0);       // + the halo row below (if any)
    for (int i = threadIdx.x; i < C; i += 256) { s_coef[i] = p.scale[i]; s_coef[C + i] = p.shift[i]; }
    __syncthreads();
    const int row_items = W * cgs;
    // phase 1: activate rows y0 .. y0+ld_rows-1 into shared memory (and to `a` for the owned rows)
    {
        const uint4* ysrc = reinterpret_cast<const uint4*>(p.y) + ((size_t)n * H + y0) * row_items;
        uint4* adst = p.a ? reinterpret_cast<uint4*>(p.a) + ((size_t)n * H + y0) * row_items : nullptr;
        const int total = ld_rows * row_items;
        for (int i0 = threadIdx.x; i0 < total; i0 += 256 * 4) {
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i0 + u * 256 < total) r[u] = __ldg(ysrc + i0 + u * 256);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 256;
                if (i >= total) continue;
                const int c0 = (i % cgs) * 8;
                float v[8];
                act8(r[u], s_coef + c0, s_coef + C + c0, v);
                const uint4 o = pack8(v);
                s_a[i] = o;
                if (adst && i < rows * row_items) adst[i] = o;
            }
        }
    }
    __syncthreads();
    // phase 2: 2x2 max-pool of the owned rows (bands start on even rows; R is even)
    if (p.a_pool) {
        const int w2 = W >> 1, prow = rows >> 1;
        uint4* dst = reinterpret_cast<uint4*>(p.a_pool) + ((size_t)n * (H >> 1) + (y0 >> 1)) * w2 * cgs;
        for (int i = threadIdx.x; i < prow * w2 * cgs; i += 256) {
            const int cg = i % cgs, x = (i / cgs) % w2, yy = i / (cgs * w2);
            const uint4 *q = s_a + ((size_t)(2 * yy) * W + 2 * x) * cgs + cg;
            const uint4 r0 = q[0], r1 = q[cgs], r2 = q[(size_t)W * cgs], r3 = q[(size_t)W * cgs + cgs];
            const __nv_bfloat162 *a0 = reinterpret_cast<const __nv_bfloat162*>(&r0), *a1 = reinterpret_cast<const __nv_bfloat162*>(&r1),
                                 *a2 = reinterpret_cast<const __nv_bfloat162*>(&r2), *a3 = reinterpret_cast<const __nv_bfloat162*>(&r3);
            uint4 o;
            __nv_bfloat162* m = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) m[e] = __hmax2(__hmax2(a0[e], a1[e]), __hmax2(a2[e], a3[e]));
            dst[i] = o;
        }
    }
    // phase 3: bilinear 2x (align_corners=True, ATen area_pixel_compute_scale).  Output rows whose upper source row i0 = floor(ry*Y)
    // lies in [y0, y0+rows): the set is contiguous and the bands partition all output rows.
    if (p.a_up) {
        const int H2 = 2 * H, W2 = 2 * W;
        const float ry = (H2 > 1) ? (float)(H - 1) / (float)(H2 - 1) : 0.f;
        const float rx = (W2 > 1) ? (float)(W - 1) / (float)(W2 - 1) : 0.f;
        // first output row of the band: smallest Y with floor(ry*Y) >= y0 (search from 2*y0 - 1 downwards is not needed: ry < 1/2)
        int Ys = (y0 == 0) ? 0 : 2 * y0 - 1;
        while (Ys < H2 && (int)(ry * (float)Ys) < y0) ++Ys;
        uint4* udst = reinterpret_cast<uint4*>(p.a_up) + (size_t)n * H2 * W2 * cgs;
        const int out_items = W2 * cgs;
        for (int it = threadIdx.x; it < out_items; it += 256) {
            const int cg = it % cgs, X = it / cgs;
            const float sx = rx * (float)X;
            const int x0 = (int)sx;
            const int x1 = x0 + (x0 < W - 1);
            const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
            const uint4* col0 = s_a + (size_t)x0 * cgs + cg;
            const uint4* col1 = s_a + (size_t)x1 * cgs + cg;
            float h0[8], h1[8];   // horizontally blended values of source rows `cur` and `cur + 1`
            int cur = -1;
            for (int Y = Ys; Y < H2; ++Y) {
                const float sy = ry * (float)Y;
                const int i0 = (int)sy;
                if (i0 >= y0 + rows) break;
                const int i1 = i0 + (i0 < H - 1);
                const float ly1 = sy - (float)i0, ly0 = 1.f - ly1;
                if (i0 != cur) {
                    // (re)load: rows i0 and i1 relative to the band
                    float ta[8], tb[8];
                    if (i0 == cur + 1 && cur >= 0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) h0[e] = h1[e];
                    } else {
                        unpack8(col0[(size_t)(i0 - y0) * row_items], ta);
                        unpack8(col1[(size_t)(i0 - y0) * row_items], tb);
#pragma unroll
                        for (int e = 0; e < 8; ++e) h0[e] = lx0 * ta[e] + lx1 * tb[e];
                    }
                    unpack8(col0[(size_t)(i1 - y0) * row_items], ta);
                    unpack8(col1[(size_t)(i1 - y0) * row_items], tb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) h1[e] = lx0 * ta[e] + lx1 * tb[e];
                    cur = i0;
                }
                float v[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = ly0 * h0[e] + ly1 * h1[e];
                udst[(size_t)Y * out_items + it] = pack8(v);
            }
        }
    }
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_bn_activate(const nunet_activate_desc_t* d, void* stream) {
    if (!d || !d->y || !d->scale || !d->shift || d->N < 1 || d->H < 1 || d->W < 1 || d->C < 8 || d->C % 8 != 0 || d->C > 4096) return NUNET_E_BADARG;
    if (!d->a && !d->a_pool && !d->a_up && !d->logits) return NUNET_E_BADARG;
    if (d->a_pool && ((d->H | d->W) & 1)) return NUNET_E_BADARG;
    if (d->logits && (d->C != 32 || !d->wh || d->K < 1 || d->K > 8 || d->a_pool || d->a_up || (long long)d->N * d->H * d->W >= (1LL << 31))) return NUNET_E_UNSUPPORTED;
    if ((long long)d->N * d->H * d->W * (d->C / 8) >= (1LL << 40)) return NUNET_E_UNSUPPORTED;
    ActParams p{(const __nv_bfloat16*)d->y, d->N, d->H, d->W, d->C, d->scale, d->shift, (__nv_bfloat16*)d->a, (__nv_bfloat16*)d->a_pool,
                (__nv_bfloat16*)d->a_up, d->wh, d->bh, d->K, d->logits};
    cudaStream_t st = (cudaStream_t)stream;
    const int cgs = d->C / 8;
    if (d->a_up) {
        // band height: even, <= kBandRows, and (R+1) rows of W*C bf16 within ~96 KB of shared memory
        int R = kBandRows;
        const size_t row_bytes = (size_t)d->W * d->C * 2;
        while (R > 2 && (size_t)(R + 1) * row_bytes + 2 * d->C * 4 > 96 * 1024) R -= 2;
        const size_t smem = (size_t)(R + 1) * row_bytes + 2 * (size_t)d->C * 4;
        if (smem > 200 * 1024 || (d->a_pool && (d->H & 1))) return NUNET_E_UNSUPPORTED;
        const int bands = (d->H + R - 1) / R;
        if ((long long)d->N * bands >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
        cudaError_t e = cudaFuncSetAttribute(activate_up_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
        activate_up_kernel<<<(unsigned)(d->N * bands), 256, smem, st>>>(p, R, bands);
    } else if (d->a_pool) {
        const long long items = (long long)d->N * (d->H / 2) * (d->W / 2) * cgs;
        long long blocks = (items + 255) / 256;
        if (blocks > 148 * 8) blocks = 148 * 8;
        activate_pool_kernel<<<(int)blocks, 256, 2 * d->C * sizeof(float), st>>>(p);
    } else {
        const long long ngroups = (long long)d->N * d->H * d->W * cgs;
        long long blocks = (ngroups + 1023) / 1024;
        if (blocks > 148 * 8) blocks = 148 * 8;
        const size_t smem = (2 * (size_t)d->C + (d->logits ? (size_t)d->K * 33 : 0)) * sizeof(float);
        if (d->logits) activate_stream_kernel<true><<<(int)blocks, 256, smem, st>>>(p);
        else activate_stream_kernel<false><<<(int)blocks, 256, smem, st>>>(p);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
