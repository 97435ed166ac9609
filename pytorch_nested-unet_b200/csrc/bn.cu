// BatchNorm statistics: reduction of the conv epilogue's per-CTA partials and emission of the apply coefficients
// consumed by the NEXT conv's loader (train mode), running-stat update, eval-mode folding, and the backward
// coefficient kernel.  All HBM-trivial (C <= 512 channels); latency matters, so one launch does reduce + finalize via
// the last-block-done pattern (deterministic: the last block sums the fixed RS slices in a fixed order, in fp64).
#include "conv_common.cuh"
#include "bn_tail.cuh"

namespace nunet {

constexpr int kRS = 32;  // row slices per channel group
constexpr int kDirectRows = 1024;  // up to here one CTA per 32 channels reduces the partial rows directly

// grid (C/32, kRS), block 256 = 32 channels x 8 row lanes.  part: [rows][2][C].
// sums_out (optional): [2][C] fp64 — the cross-rank SyncBN exchange buffer.  fin.gamma == nullptr -> reduce only.
__global__ void __launch_bounds__(256) bn_reduce_kernel(const float* __restrict__ part, int rows, int C, double* scratch,
                                                        unsigned* counters, double* sums_out, BnFinalizeArgs fin) {
    __shared__ float sh[2][8][32];
    __shared__ bool last;
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int per = (rows + kRS - 1) / kRS;
    const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
    float s = 0.f, ss = 0.f;
    for (int r = r0 + rl; r < r1; r += 8) {
        s += part[(size_t)r * 2 * C + c];
        ss += part[(size_t)r * 2 * C + C + c];
    }
    sh[0][rl][cl] = s;
    sh[1][rl][cl] = ss;
    __syncthreads();
    if (threadIdx.x < 64) {
        const int which = threadIdx.x >> 5;
        double acc = 0.0;
        for (int i = 0; i < 8; ++i) acc += (double)sh[which][i][cl];
        scratch[((size_t)blockIdx.y * 2 + which) * C + c] = acc;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(&counters[blockIdx.x], 1u);
        last = (prev == kRS - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double S = 0.0, SS = 0.0;
        for (int i = 0; i < kRS; ++i) {
            S += scratch[((size_t)i * 2 + 0) * C + c];
            SS += scratch[((size_t)i * 2 + 1) * C + c];
        }
        if (sums_out) { sums_out[c] = S; sums_out[C + c] = SS; }
        if (fin.gamma) bn_finalize_channel(fin, c, S, SS);
    }
    if (threadIdx.x == 0) counters[blockIdx.x] = 0;  // self-resetting for the next launch
}

// Single-pass variant for the row counts the conv / router kernels actually produce (<= a few hundred partial rows): one CTA
// per 32 channels, 32 row lanes x 32 channels, fp64 accumulation per thread in a fixed order, no scratch / atomics / second
// phase (the last-block-done kernel above costs ~12 us per BatchNorm layer, 60 times per step; this one ~3 us).
__global__ void __launch_bounds__(1024) bn_reduce_direct_kernel(const float* __restrict__ part, int rows, int C, double* sums_out,
                                                                 BnFinalizeArgs fin) {
    __shared__ double sh[2][32][33];
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s = 0.0, ss = 0.0;
    for (int r = rl; r < rows; r += 32) {
        s += (double)part[(size_t)r * 2 * C + c];
        ss += (double)part[(size_t)r * 2 * C + C + c];
    }
    sh[0][rl][cl] = s;
    sh[1][rl][cl] = ss;
    __syncthreads();
    if (threadIdx.x < 32) {
        double S = 0.0, SS = 0.0;
        for (int i = 0; i < 32; ++i) { S += sh[0][i][cl]; SS += sh[1][i][cl]; }
        if (sums_out) { sums_out[c] = S; sums_out[C + c] = SS; }
        if (fin.gamma) bn_finalize_channel(fin, c, S, SS);
    }
}

__global__ void bn_finalize_from_sums_kernel(const double* __restrict__ sums, int C, BnFinalizeArgs fin) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) bn_finalize_channel(fin, c, sums[c], sums[C + c]);
}

__global__ void bn_fold_eval_kernel(int C, const float* gamma, const float* beta, const float* bias, const float* rm,
                                    const float* rv, float eps, float* scale, float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float sc = gamma[c] * (1.0f / sqrtf(rv[c] + eps));
    scale[c] = sc;
    shift[c] = beta[c] + ((bias ? bias[c] : 0.f) - rm[c]) * sc;
}

// a = relu(scale*y+shift): one 16-byte group (8 channels) per thread-iteration, fully coalesced, HBM-bound.
__global__ void __launch_bounds__(256) bn_apply_kernel(const uint4* __restrict__ y, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, long long ngroups, int cgs, uint4* __restrict__ a) {
    extern __shared__ float s_coef[];  // [2][C]
    const int C = cgs * 8;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = scale[i]; s_coef[C + i] = shift[i]; }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ngroups; g += 4 * stride) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (g + u * stride < ngroups) r[u] = __ldg(y + g + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long gi = g + u * stride;
            if (gi >= ngroups) continue;
            const int c0 = (int)(gi & (long long)(cgs - 1)) * 8;   // C/8 is a power of two (host-checked)
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r[u]);
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = __bfloat1622float2(h[i]);
                const float v0 = fmaxf(fmaf(t.x, s_coef[c0 + 2 * i], s_coef[C + c0 + 2 * i]), 0.f);
                const float v1 = fmaxf(fmaf(t.y, s_coef[c0 + 2 * i + 1], s_coef[C + c0 + 2 * i + 1]), 0.f);
                __nv_bfloat162 q = __floats2bfloat162_rn(v0, v1);
                o[i] = *reinterpret_cast<uint32_t*>(&q);
            }
            a[gi] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// g_y = k1*g_z + k2*y + k3 per channel; 16-byte groups, coalesced; gy may alias gz.
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const uint4* gz, const uint4* __restrict__ y, const float* __restrict__ k1,
                                                           const float* __restrict__ k2, const float* __restrict__ k3, long long ngroups,
                                                           int cgs, uint4* gy) {
    extern __shared__ float s_coef[];  // [3][C]
    const int C = cgs * 8;
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_coef[i] = k1[i]; s_coef[C + i] = k2[i]; s_coef[2 * C + i] = k3[i]; }
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x; g < ngroups; g += 2 * stride) {
        uint4 rg[2], ry[2];
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (g + u * stride < ngroups) { rg[u] = gz[g + u * stride]; ry[u] = __ldg(y + g + u * stride); }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const long long gi = g + u * stride;
            if (gi >= ngroups) continue;
            const int c0 = (int)(gi & (long long)(cgs - 1)) * 8;
            const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&rg[u]);
            const __nv_bfloat162* hy = reinterpret_cast<const __nv_bfloat162*>(&ry[u]);
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 a = __bfloat1622float2(hg[i]), b = __bfloat1622float2(hy[i]);
                const int c = c0 + 2 * i;
                const float v0 = fmaf(a.x, s_coef[c], fmaf(b.x, s_coef[C + c], s_coef[2 * C + c]));
                const float v1 = fmaf(a.y, s_coef[c + 1], fmaf(b.y, s_coef[C + c + 1], s_coef[2 * C + c + 1]));
                __nv_bfloat162 q = __floats2bfloat162_rn(v0, v1);
                o[i] = *reinterpret_cast<uint32_t*>(&q);
            }
            gy[gi] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// Bilinear 2x upsample (align_corners=True, ATen area_pixel_compute_scale) of an activated bf16 NHWC tensor:
// out[N,2h,2w,C].  One 16-byte group per thread-iteration, fp32 blend, one rounding.  Used when the engine materialises
// the upsampled skip operand once instead of re-blending it in every loader that reads it (forward + Cout/32 wgrad CTAs).
constexpr int kUpRows = 8;   // output rows per CTA
__global__ void __launch_bounds__(256) upsample2x_kernel(const __nv_bfloat16* __restrict__ in, int N, int h, int w, int cgs,
                                                         __nv_bfloat16* __restrict__ out) {
    // CTA = kUpRows consecutive output rows of one image; a thread owns (output column X, channel group) items and walks the
    // rows, so the column indices / weights are computed once and no integer division is needed per item.
    const int H = 2 * h, W = 2 * w, C = cgs * 8;
    const int cg_shift = 31 - __clz(cgs);  // C/8 is a power of two
    const int rgroups = (H + kUpRows - 1) / kUpRows;
    const int n = blockIdx.x / rgroups, Y0 = (blockIdx.x - n * rgroups) * kUpRows;
    const float ry = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const __nv_bfloat16* img = in + (size_t)n * h * w * C;
    __nv_bfloat16* oimg = out + (size_t)n * H * W * C;
    const int row_items = W << cg_shift;
    for (int i = threadIdx.x; i < row_items; i += 256) {
        const int cg = i & (cgs - 1), X = i >> cg_shift;
        const float sx = rx * (float)X;
        const int x0 = (int)sx;
        const int x1 = x0 + (x0 < w - 1);
        const float lx1 = sx - (float)x0, lx0 = 1.f - lx1;
        const __nv_bfloat16* c0 = img + (size_t)x0 * C + cg * 8;
        const __nv_bfloat16* c1 = img + (size_t)x1 * C + cg * 8;
#pragma unroll 2
        for (int r = 0; r < kUpRows; ++r) {
            const int Y = Y0 + r;
            if (Y >= H) break;
            const float sy = ry * (float)Y;
            const int y0 = (int)sy;
            const int y1 = y0 + (y0 < h - 1);
            const float ly1 = sy - (float)y0, ly0 = 1.f - ly1;
            const uint4 ra = __ldg(reinterpret_cast<const uint4*>(c0 + (size_t)y0 * w * C));
            const uint4 rb = __ldg(reinterpret_cast<const uint4*>(c1 + (size_t)y0 * w * C));
            const uint4 rc = __ldg(reinterpret_cast<const uint4*>(c0 + (size_t)y1 * w * C));
            const uint4 rd = __ldg(reinterpret_cast<const uint4*>(c1 + (size_t)y1 * w * C));
            const __nv_bfloat162 *a = reinterpret_cast<const __nv_bfloat162*>(&ra), *b = reinterpret_cast<const __nv_bfloat162*>(&rb),
                                 *c = reinterpret_cast<const __nv_bfloat162*>(&rc), *d = reinterpret_cast<const __nv_bfloat162*>(&rd);
            uint32_t o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 fa = __bfloat1622float2(a[e]), fb = __bfloat1622float2(b[e]), fc = __bfloat1622float2(c[e]), fd = __bfloat1622float2(d[e]);
                const float v0 = ly0 * (lx0 * fa.x + lx1 * fb.x) + ly1 * (lx0 * fc.x + lx1 * fd.x);
                const float v1 = ly0 * (lx0 * fa.y + lx1 * fb.y) + ly1 * (lx0 * fc.y + lx1 * fd.y);
                __nv_bfloat162 q = __floats2bfloat162_rn(v0, v1);
                o[e] = *reinterpret_cast<uint32_t*>(&q);
            }
            *reinterpret_cast<uint4*>(oimg + ((size_t)Y * W + X) * C + cg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

// 2x2 max-pool (stride 2) of an activated bf16 NHWC tensor: out[N,H/2,W/2,C].  Materialised once for the backbone nodes so
// that the pooled consumer's forward and weight-gradient convolutions run in pure copy / TMA mode.
__global__ void __launch_bounds__(256) maxpool2x2_kernel(const uint4* __restrict__ in, int N, int H, int W, int cgs, uint4* __restrict__ out) {
    const int h = H >> 1, w = W >> 1;
    const int cg_shift = 31 - __clz(cgs);
    const unsigned items = (unsigned)N * h * w * cgs;
    for (unsigned it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
        const int cg = it & (cgs - 1);
        unsigned pix = it >> cg_shift;
        const int x = pix % w; pix /= w;
        const int y = pix % h;
        const int n = pix / h;
        const size_t base = (((size_t)n * H + 2 * y) * W + 2 * x) * cgs + cg;
        const uint4 r0 = __ldg(in + base), r1 = __ldg(in + base + cgs), r2 = __ldg(in + base + (size_t)W * cgs), r3 = __ldg(in + base + (size_t)W * cgs + cgs);
        const __nv_bfloat162 *a = reinterpret_cast<const __nv_bfloat162*>(&r0), *b = reinterpret_cast<const __nv_bfloat162*>(&r1),
                             *c = reinterpret_cast<const __nv_bfloat162*>(&r2), *d = reinterpret_cast<const __nv_bfloat162*>(&r3);
        uint4 o;
        __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = __hmax2(__hmax2(a[i], b[i]), __hmax2(c[i], d[i]));
        out[it] = o;
    }
}

// Network input fp32 NCHW [N,C,H,W] (C <= 32) -> bf16 NHWC [N,H,W,32], channels >= C zero: the stem conv then runs as a
// pre-activated 32-channel copy-mode source (forward) and through the TMA weight-gradient kernel.
__global__ void __launch_bounds__(256) input_to_nhwc32_kernel(const float* __restrict__ x, int N, int C, long long HW, uint4* __restrict__ out) {
    const long long items = (long long)N * HW * 4;   // 16-byte groups of 8 channels
    for (long long it = blockIdx.x * (long long)blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const int cg = (int)(it & 3);
        const long long pix = it >> 2;
        const long long n = pix / HW, hw = pix - n * HW;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = cg * 8 + e;
            v[e] = (c < C) ? __ldg(x + ((size_t)n * C + c) * HW + hw) : 0.f;
        }
        uint4 o;
        __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) q[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        out[it] = o;
    }
}

}  // namespace nunet

using namespace nunet;

static inline int cuda_rc() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}

extern "C" int nunet_bn_workspace_bytes(int32_t C) { return (int)(sizeof(double) * kRS * 2 * C); }

extern "C" int nunet_bn_finalize_train(const float* stat_part, int32_t rows, int32_t C, double count, const float* gamma,
                                       const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                                       float momentum, float eps, float* scale, float* shift, float* save_mean,
                                       float* save_invstd, double* scratch, uint32_t* counters, double* sums_out,
                                       void* stream) {
    if (!stat_part || rows <= 0 || C <= 0 || C % 32 != 0 || !scratch || !counters) return NUNET_E_BADARG;
    if (gamma && (!beta || !scale || !shift || count <= 0)) return NUNET_E_BADARG;
    BnFinalizeArgs f{0, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, scale, shift, save_mean, save_invstd};
    if (rows <= kDirectRows)
        bn_reduce_direct_kernel<<<C / 32, 1024, 0, (cudaStream_t)stream>>>(stat_part, rows, C, sums_out, f);
    else
        bn_reduce_kernel<<<dim3(C / 32, kRS), 256, 0, (cudaStream_t)stream>>>(stat_part, rows, C, scratch, counters, sums_out, f);
    return cuda_rc();
}

extern "C" int nunet_bn_finalize_from_sums(const double* sums, int32_t C, double count, const float* gamma, const float* beta,
                                           const float* conv_bias, float* running_mean, float* running_var, float momentum,
                                           float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                                           void* stream) {
    if (!sums || C <= 0 || !gamma || !beta || !scale || !shift || count <= 0) return NUNET_E_BADARG;
    BnFinalizeArgs f{0, count, gamma, beta, conv_bias, running_mean, running_var, momentum, eps, scale, shift, save_mean, save_invstd};
    bn_finalize_from_sums_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, C, f);
    return cuda_rc();
}

extern "C" int nunet_bn_fold_eval(int32_t C, const float* gamma, const float* beta, const float* conv_bias,
                                  const float* running_mean, const float* running_var, float eps, float* scale, float* shift,
                                  void* stream) {
    if (C <= 0 || !gamma || !beta || !running_mean || !running_var || !scale || !shift) return NUNET_E_BADARG;
    bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, conv_bias, running_mean, running_var,
                                                                           eps, scale, shift);
    return cuda_rc();
}

extern "C" int nunet_bn_backward_reduce(const float* part, int32_t rows, int32_t C, double count, const float* gamma,
                                        const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* k1,
                                        float* k2, float* k3, double* scratch, uint32_t* counters, double* sums_out, void* stream) {
    if (!part || rows <= 0 || C <= 0 || C % 32 != 0 || !scratch || !counters) return NUNET_E_BADARG;
    if (gamma && (!save_mean || !save_invstd || !dgamma || !dbeta || !k1 || !k2 || !k3 || count <= 0)) return NUNET_E_BADARG;
    BnFinalizeArgs f{};
    f.mode = 1; f.count = count; f.gamma = gamma; f.save_mean = const_cast<float*>(save_mean);
    f.save_invstd = const_cast<float*>(save_invstd); f.dgamma = dgamma; f.dbeta = dbeta; f.k1 = k1; f.k2 = k2; f.k3 = k3;
    f.grad_scale = 1.f;
    if (rows <= kDirectRows)
        bn_reduce_direct_kernel<<<C / 32, 1024, 0, (cudaStream_t)stream>>>(part, rows, C, sums_out, f);
    else
        bn_reduce_kernel<<<dim3(C / 32, kRS), 256, 0, (cudaStream_t)stream>>>(part, rows, C, scratch, counters, sums_out, f);
    return cuda_rc();
}

extern "C" int nunet_bn_backward_from_sums(const double* sums, int32_t C, double count, const float* gamma, const float* save_mean,
                                           const float* save_invstd, float* dgamma, float* dbeta, float* k1, float* k2, float* k3,
                                           float grad_scale, void* stream) {
    if (!sums || C <= 0 || !gamma || !save_mean || !save_invstd || !dgamma || !dbeta || !k1 || !k2 || !k3 || count <= 0) return NUNET_E_BADARG;
    BnFinalizeArgs f{};
    f.mode = 1; f.count = count; f.gamma = gamma; f.save_mean = const_cast<float*>(save_mean);
    f.save_invstd = const_cast<float*>(save_invstd); f.dgamma = dgamma; f.dbeta = dbeta; f.k1 = k1; f.k2 = k2; f.k3 = k3;
    f.grad_scale = grad_scale;
    bn_finalize_from_sums_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(sums, C, f);
    return cuda_rc();
}

extern "C" int nunet_bn_apply(const void* y, const float* scale, const float* shift, long long npix, int32_t C, void* a, void* stream) {
    if (!y || !scale || !shift || !a || npix <= 0 || C <= 0 || C % 8 != 0 || C > 4096 || ((C / 8) & (C / 8 - 1)) != 0) return NUNET_E_BADARG;
    const long long ngroups = npix * (C / 8);
    long long blocks = (ngroups + 1023) / 1024;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bn_apply_kernel<<<(int)blocks, 256, 2 * C * sizeof(float), (cudaStream_t)stream>>>((const uint4*)y, scale, shift, ngroups, C / 8, (uint4*)a);
    return cuda_rc();
}

extern "C" int nunet_bn_bwd_apply(const void* gz, const void* y, const float* k1, const float* k2, const float* k3, long long npix,
                                  int32_t C, void* gy, void* stream) {
    if (!gz || !y || !k1 || !k2 || !k3 || !gy || npix <= 0 || C <= 0 || C % 8 != 0 || C > 2048 || ((C / 8) & (C / 8 - 1)) != 0) return NUNET_E_BADARG;
    const long long ngroups = npix * (C / 8);
    long long blocks = (ngroups + 511) / 512;
    if (blocks > 148 * 8) blocks = 148 * 8;
    bn_bwd_apply_kernel<<<(int)blocks, 256, 3 * C * sizeof(float), (cudaStream_t)stream>>>((const uint4*)gz, (const uint4*)y, k1, k2, k3, ngroups,
                                                                                          C / 8, (uint4*)gy);
    return cuda_rc();
}

extern "C" int nunet_upsample2x(const void* in, int32_t N, int32_t h, int32_t w, int32_t C, void* out, void* stream) {
    if (!in || !out || N < 1 || h < 1 || w < 1 || C < 8 || C % 8 != 0) return NUNET_E_BADARG;
    const long long items = (long long)N * 4 * h * w * (C / 8);
    if (items >= (1LL << 31) || ((C / 8) & (C / 8 - 1)) != 0) return NUNET_E_UNSUPPORTED;
    const long long blocks = (long long)N * ((2 * h + kUpRows - 1) / kUpRows);
    if (blocks >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    upsample2x_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)in, N, h, w, C / 8, (__nv_bfloat16*)out);
    return cuda_rc();
}

extern "C" int nunet_maxpool2x2(const void* in, int32_t N, int32_t H, int32_t W, int32_t C, void* out, void* stream) {
    if (!in || !out || N < 1 || H < 2 || W < 2 || (H & 1) || (W & 1) || C < 8 || C % 8 != 0 || ((C / 8) & (C / 8 - 1)) != 0) return NUNET_E_BADARG;
    const long long items = (long long)N * (H / 2) * (W / 2) * (C / 8);
    if (items >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    maxpool2x2_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)in, N, H, W, C / 8, (uint4*)out);
    return cuda_rc();
}

extern "C" int nunet_input_to_nhwc32(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, void* out, void* stream) {
    if (!x || !out || N < 1 || C < 1 || C > 32 || H < 1 || W < 1) return NUNET_E_BADARG;
    const long long items = (long long)N * H * W * 4;
    long long blocks = (items + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    input_to_nhwc32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, N, C, (long long)H * W, (uint4*)out);
    return cuda_rc();
}
