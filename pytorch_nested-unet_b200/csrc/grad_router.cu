// Gradient router + BatchNorm/ReLU backward reduction for one conv output (one pass over its pixels).
//
// For a raw conv output y (bf16 NHWC, C channels) whose activation a = relu(scale*y+shift) was consumed by up to six
// later convs (each consumer's dgrad wrote a bf16 NHWC buffer dX at ITS resolution; our channels are a slice of it)
// and possibly by a 1x1 head, this kernel
//   1. gathers  g_a = sum_consumers route(dX slice)     route = identity (SAME) | max-pool argmax routing (POOL,
//      first maximum in window scan order, as ATen) | transposed align_corners bilinear 2x (UP) | dlogits * Wh (head)
//      — the backward of torch.cat / nn.MaxPool2d / nn.Upsample / the head conv (finished/archs1.py:82-83,117-142)
//   2. applies the ReLU mask: g_z = g_a * [a > 0]                                   (finished/archs1.py:26,30)
//   3. stores g_z (bf16) and reduces sum(g_z), sum(g_z * yhat) per channel into per-CTA partials — the two
//      reductions of nn.BatchNorm2d's backward (yhat = (y-mean)*invstd).
// The affine rest of BN backward (g_y = k1*g_z + k2*y + k3) is folded into the loaders of dgrad / wgrad.
#include "conv_common.cuh"

namespace nunet {

struct GradCons {
    const __nv_bfloat16* dx;
    int pitch, coff, kind;
};
struct RouterParams {
    int N, H, W, C;
    const __nv_bfloat16* y;
    const float *scale, *shift, *mean, *invstd;
    int num_cons;
    GradCons cons[6];
    const float* dlogits;  // fp32 NCHW [N,K,H,W] or null
    const float* wh;       // [K][C]
    int K;
    __nv_bfloat16* gz;
    float* part;  // [gridDim.x][2][C]
};

__device__ __forceinline__ void add8(float (&g)[8], const uint4& r, float w) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        g[2 * i] = fmaf(w, t.x, g[2 * i]);
        g[2 * i + 1] = fmaf(w, t.y, g[2 * i + 1]);
    }
}
__device__ __forceinline__ void unpack8f(const uint4& r, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

// weights with which output index D (size 2*h) of an align_corners 2x upsample reads source index s (size h)
__device__ __forceinline__ float up_weight(int D, int s, int h, float r) {
    const float src = r * (float)D;
    const int i0 = (int)src;
    const int i1 = i0 + (i0 < h - 1);
    const float l1 = src - (float)i0;
    return (i0 == s ? 1.f - l1 : 0.f) + (i1 == s ? l1 : 0.f);
}

__global__ void __launch_bounds__(256) grad_router_kernel(const RouterParams p) {
    const int cgs = p.C >> 3;
    const int cg = threadIdx.x % cgs;
    const int HW = p.H * p.W;
    const long long items = (long long)p.N * HW * cgs;
    float sc[8], sh[8], mu[8], rs[8], acc1[8], acc2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = p.scale[cg * 8 + i];
        sh[i] = p.shift[cg * 8 + i];
        mu[i] = p.mean[cg * 8 + i];
        rs[i] = p.invstd[cg * 8 + i];
        acc1[i] = 0.f;
        acc2[i] = 0.f;
    }
    for (long long it = blockIdx.x * 256LL + threadIdx.x; it < items; it += (long long)gridDim.x * 256) {
        const int pix = (int)(it / cgs);
        const int n = pix / HW, rem = pix - n * HW;
        const int yy = rem / p.W, xx = rem - yy * p.W;
        const int ch = cg * 8;
        float yv[8];
        unpack8f(__ldg(reinterpret_cast<const uint4*>(p.y + (size_t)pix * p.C + ch)), yv);
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = 0.f;
        for (int ci = 0; ci < p.num_cons; ++ci) {
            const GradCons& c = p.cons[ci];
            if (c.kind == NUNET_SRC_SAME) {
                add8(g, __ldg(reinterpret_cast<const uint4*>(c.dx + (size_t)pix * c.pitch + c.coff + ch)), 1.f);
            } else if (c.kind == NUNET_SRC_POOL) {
                // consumer at (H/2, W/2): route to the first maximum of the 2x2 window of a = relu(scale*y+shift)
                const int y0 = yy & ~1, x0 = xx & ~1;
                const int me = ((yy & 1) << 1) | (xx & 1);
                const size_t wbase = ((size_t)n * HW + (size_t)y0 * p.W + x0);
                float av[4][8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float t[8];
                    if (k == me) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) t[i] = yv[i];
                    } else {
                        unpack8f(__ldg(reinterpret_cast<const uint4*>(p.y + (wbase + (size_t)(k >> 1) * p.W + (k & 1)) * p.C + ch)), t);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) av[k][i] = fmaxf(fmaf(t[i], sc[i], sh[i]), 0.f);
                }
                const size_t cpix = ((size_t)n * (p.H >> 1) + (yy >> 1)) * (p.W >> 1) + (xx >> 1);
                float d[8];
                unpack8f(__ldg(reinterpret_cast<const uint4*>(c.dx + cpix * c.pitch + c.coff + ch)), d);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    bool win = true;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k < me) win = win && (av[me][i] > av[k][i]);
                        if (k > me) win = win && (av[me][i] >= av[k][i]);
                    }
                    if (win) g[i] += d[i];
                }
            } else {  // NUNET_SRC_UP: consumer at (2H, 2W)
                const int H2 = 2 * p.H, W2 = 2 * p.W;
                const float ry = (float)(p.H - 1) / (float)(H2 - 1), rx = (float)(p.W - 1) / (float)(W2 - 1);
                float wy[8], wx[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int Y = 2 * yy - 3 + k, X = 2 * xx - 3 + k;
                    wy[k] = (Y >= 0 && Y < H2) ? up_weight(Y, yy, p.H, ry) : 0.f;
                    wx[k] = (X >= 0 && X < W2) ? up_weight(X, xx, p.W, rx) : 0.f;
                }
#pragma unroll
                for (int ky = 0; ky < 8; ++ky) {
                    if (wy[ky] == 0.f) continue;
                    const int Y = 2 * yy - 3 + ky;
#pragma unroll
                    for (int kx = 0; kx < 8; ++kx) {
                        if (wx[kx] == 0.f) continue;
                        const int X = 2 * xx - 3 + kx;
                        const size_t cpix = ((size_t)n * H2 + Y) * W2 + X;
                        add8(g, __ldg(reinterpret_cast<const uint4*>(c.dx + cpix * c.pitch + c.coff + ch)), wy[ky] * wx[kx]);
                    }
                }
            }
        }
        if (p.dlogits != nullptr) {
            for (int k = 0; k < p.K; ++k) {
                const float dl = __ldg(p.dlogits + ((size_t)n * p.K + k) * HW + rem);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = fmaf(dl, __ldg(p.wh + k * p.C + ch + i), g[i]);
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = fmaf(yv[i], sc[i], sh[i]);
            const float gz = (z > 0.f) ? g[i] : 0.f;
            g[i] = gz;
            acc1[i] += gz;
            acc2[i] = fmaf(gz, (yv[i] - mu[i]) * rs[i], acc2[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
            o[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p.gz + (size_t)pix * p.C + ch) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    // block reduction: threads with equal cg (tid % cgs) hold the same 8 channels
    __shared__ float red[256][9];
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = pass == 0 ? acc1[i] : acc2[i];
        __syncthreads();
        for (int c = threadIdx.x; c < p.C; c += 256) {
            const int g8 = c >> 3, e = c & 7;
            float s = 0.f;
            for (int t = g8; t < 256; t += cgs) s += red[t][e];
            p.part[((size_t)blockIdx.x * 2 + pass) * p.C + c] = s;
        }
        __syncthreads();
    }
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_grad_router_rows(int32_t N, int32_t H, int32_t W, int32_t C) {
    long long items = (long long)N * H * W * (C / 8);
    long long b = (items + 255) / 256;
    if (b > 148 * 8) b = 148 * 8;
    return (int)(b < 1 ? 1 : b);
}

extern "C" int nunet_grad_router(const nunet_router_desc_t* d, void* stream) {
    if (!d || !d->y || !d->scale || !d->shift || !d->mean || !d->invstd || !d->gz || !d->part) return NUNET_E_BADARG;
    if (d->C % 32 != 0 || d->C > 512 || d->num_cons < 0 || d->num_cons > 6) return NUNET_E_BADARG;
    if ((long long)d->N * d->H * d->W >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    if (d->dlogits && (!d->wh || d->K < 1)) return NUNET_E_BADARG;
    RouterParams p{};
    p.N = d->N; p.H = d->H; p.W = d->W; p.C = d->C;
    p.y = (const __nv_bfloat16*)d->y; p.scale = d->scale; p.shift = d->shift; p.mean = d->mean; p.invstd = d->invstd;
    p.num_cons = d->num_cons;
    for (int i = 0; i < d->num_cons; ++i) {
        const nunet_grad_cons_t& c = d->cons[i];
        if (!c.dx || c.pitch % 8 != 0 || c.coff % 8 != 0) return NUNET_E_BADARG;
        if (c.kind != NUNET_SRC_SAME && c.kind != NUNET_SRC_POOL && c.kind != NUNET_SRC_UP) return NUNET_E_BADARG;
        if (c.kind == NUNET_SRC_POOL && ((d->H | d->W) & 1)) return NUNET_E_UNSUPPORTED;
        p.cons[i].dx = (const __nv_bfloat16*)c.dx; p.cons[i].pitch = c.pitch; p.cons[i].coff = c.coff; p.cons[i].kind = c.kind;
    }
    p.dlogits = d->dlogits; p.wh = d->wh; p.K = d->K;
    p.gz = (__nv_bfloat16*)d->gz; p.part = d->part;
    const int rows = nunet_grad_router_rows(d->N, d->H, d->W, d->C);
    grad_router_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
