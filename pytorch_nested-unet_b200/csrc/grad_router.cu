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
#include "loaders.cuh"
#include <cstdlib>

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
    BnTail tail;  // BatchNorm-backward coefficients by the last CTA
    // fused head parameter gradients (K == 1, C == 32): hpart [gridDim.x][33] per-CTA partials of dWh[c] = sum dl*a, dbh = sum dl;
    // reduced into dwh [32] / dbh [1] by the last CTA (needs tail.counter)
    float* hpart;
    float *dwh, *dbh;
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

// Consumers are ordered by the host: cons[0..n_same) are SAME, then (HAS_POOL) one POOL, then (HAS_UP) one UP.
template <bool HAS_POOL, bool HAS_UP, bool HAS_HEAD>
__global__ void __launch_bounds__(256, 2) grad_router_kernel(const RouterParams p) {
    const int cgs = p.C >> 3;
    const int cg_shift = 31 - __clz(cgs);  // C/8 is a power of two (4..64)
    const int cg = threadIdx.x & (cgs - 1);
    const int ch = cg * 8;
    const int HW = p.H * p.W;
    const int items = p.N * HW * cgs;
    const int n_same = p.num_cons - (HAS_POOL ? 1 : 0) - (HAS_UP ? 1 : 0);
    float sc[8], sh[8], mu[8], rs[8], acc1[8], acc2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = p.scale[ch + i];
        sh[i] = p.shift[ch + i];
        mu[i] = p.mean[ch + i];
        rs[i] = p.invstd[ch + i];
        acc1[i] = 0.f;
        acc2[i] = 0.f;
    }
    constexpr int U = 2;
    const int stride = gridDim.x * 256;
    for (int it0 = blockIdx.x * 256 + threadIdx.x; it0 < items; it0 += stride * U) {
        uint4 ry[U], rd[U][4];
        int pix[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int it = it0 + u * stride;
            pix[u] = (it < items) ? (it >> cg_shift) : -1;
            if (pix[u] >= 0) {
                ry[u] = __ldg(reinterpret_cast<const uint4*>(p.y + (size_t)pix[u] * p.C + ch));
#pragma unroll
                for (int ci = 0; ci < 4; ++ci)
                    if (ci < n_same)
                        rd[u][ci] = __ldg(reinterpret_cast<const uint4*>(p.cons[ci].dx + (size_t)pix[u] * p.cons[ci].pitch + p.cons[ci].coff + ch));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (pix[u] < 0) continue;
            float yv[8], g[8];
            unpack8f(ry[u], yv);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = 0.f;
#pragma unroll
            for (int ci = 0; ci < 4; ++ci)
                if (ci < n_same) add8(g, rd[u][ci], 1.f);
            if constexpr (HAS_POOL || HAS_UP || HAS_HEAD) {
                const int n = pix[u] / HW, rem = pix[u] - n * HW;
                const int yy = rem / p.W, xx = rem - yy * p.W;
                if constexpr (HAS_POOL) {
                    // consumer at (H/2, W/2): route to the first maximum of the 2x2 window of a = relu(scale*y+shift)
                    const GradCons& c = p.cons[n_same];
                    const int y0 = yy & ~1, x0 = xx & ~1;
                    const int me = ((yy & 1) << 1) | (xx & 1);
                    const size_t wbase = (size_t)n * HW + (size_t)y0 * p.W + x0;
                    float am[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) am[i] = fmaxf(fmaf(yv[i], sc[i], sh[i]), 0.f);
                    bool win[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) win[i] = true;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k == me) continue;
                        float t[8];
                        unpack8f(__ldg(reinterpret_cast<const uint4*>(p.y + (wbase + (size_t)(k >> 1) * p.W + (k & 1)) * p.C + ch)), t);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float ak = fmaxf(fmaf(t[i], sc[i], sh[i]), 0.f);
                            win[i] = win[i] && ((k < me) ? (am[i] > ak) : (am[i] >= ak));
                        }
                    }
                    const size_t cpix = ((size_t)n * (p.H >> 1) + (yy >> 1)) * (p.W >> 1) + (xx >> 1);
                    float d[8];
                    unpack8f(__ldg(reinterpret_cast<const uint4*>(c.dx + cpix * c.pitch + c.coff + ch)), d);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (win[i]) g[i] += d[i];
                }
                if constexpr (HAS_UP) {  // consumer at (2H, 2W): transposed align_corners bilinear
                    const GradCons& c = p.cons[n_same + (HAS_POOL ? 1 : 0)];
                    const int H2 = 2 * p.H, W2 = 2 * p.W;
                    const float ry_ = (float)(p.H - 1) / (float)(H2 - 1), rx_ = (float)(p.W - 1) / (float)(W2 - 1);
                    // contributing outputs D of source index s satisfy |r*D - s| < 1  =>  D in (  (s-1)/r , (s+1)/r  ) ~ [2s-2, 2s+3]
#pragma unroll
                    for (int ky = 0; ky < 6; ++ky) {
                        const int Y = 2 * yy - 2 + ky;
                        const float wy = (Y >= 0 && Y < H2) ? up_weight(Y, yy, p.H, ry_) : 0.f;
                        if (wy == 0.f) continue;
#pragma unroll
                        for (int kx = 0; kx < 6; ++kx) {
                            const int X = 2 * xx - 2 + kx;
                            const float wx = (X >= 0 && X < W2) ? up_weight(X, xx, p.W, rx_) : 0.f;
                            if (wx == 0.f) continue;
                            const size_t cpix = ((size_t)n * H2 + Y) * W2 + X;
                            add8(g, __ldg(reinterpret_cast<const uint4*>(c.dx + cpix * c.pitch + c.coff + ch)), wy * wx);
                        }
                    }
                }
                if constexpr (HAS_HEAD) {
                    for (int k = 0; k < p.K; ++k) {
                        const float dl = __ldg(p.dlogits + ((size_t)n * p.K + k) * HW + rem);
#pragma unroll
                        for (int i = 0; i < 8; ++i) g[i] = fmaf(dl, __ldg(p.wh + k * p.C + ch + i), g[i]);
                    }
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
            *reinterpret_cast<uint4*>(p.gz + (size_t)pix[u] * p.C + ch) = make_uint4(o[0], o[1], o[2], o[3]);
        }
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
    if (p.tail.counter) bn_tail_run(p.tail, p.part, (int)gridDim.x, p.C, gridDim.x, threadIdx.x, 0, &red[0][0]);
}

// Specialised router for what the engine launches (no UP consumer: those go through up_transpose.cu first).  Same arithmetic
// for g_z; the second BatchNorm sum is accumulated as sum(g_z*y) and sum(g_z) and combined per thread at the end,
//   sum(g_z*yhat) = invstd * (sum(g_z*y) - mean*sum(g_z)),
// which keeps mean/invstd out of the loop: <= 64 registers, four CTAs per SM (twice the loads in flight of the generic kernel,
// which ncu showed latency-bound at 25 % occupancy).  NS = number of same-resolution consumers (compile time).
template <int NS, bool HAS_POOL, bool HAS_HEAD>
__global__ void __launch_bounds__(256, HAS_POOL ? 3 : 4) grad_router_fast_kernel(const RouterParams p) {
    const int cgs = p.C >> 3;
    const int cg_shift = 31 - __clz(cgs);
    const int cg = threadIdx.x & (cgs - 1);
    const int ch = cg * 8;
    const int HW = p.H * p.W;
    const int items = p.N * HW * cgs;
    float sc[8], sh[8], acc1[8], acc2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = p.scale[ch + i];
        sh[i] = p.shift[ch + i];
        acc1[i] = 0.f;
        acc2[i] = 0.f;
    }
    constexpr int U = (NS <= 2 && !HAS_POOL) ? 2 : 1;
    const int stride = gridDim.x * 256;
    for (int it0 = blockIdx.x * 256 + threadIdx.x; it0 < items; it0 += stride * U) {
        uint4 ry[U], rd[U][NS > 0 ? NS : 1];
        int pix[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int it = it0 + u * stride;
            pix[u] = (it < items) ? (it >> cg_shift) : -1;
            if (pix[u] >= 0) {
                ry[u] = __ldg(reinterpret_cast<const uint4*>(p.y + (size_t)pix[u] * p.C + ch));
#pragma unroll
                for (int ci = 0; ci < NS; ++ci)
                    rd[u][ci] = __ldg(reinterpret_cast<const uint4*>(p.cons[ci].dx + (size_t)pix[u] * p.cons[ci].pitch + p.cons[ci].coff + ch));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (pix[u] < 0) continue;
            float yv[8], g[8];
            unpack8f(ry[u], yv);
#pragma unroll
            for (int i = 0; i < 8; ++i) g[i] = 0.f;
#pragma unroll
            for (int ci = 0; ci < NS; ++ci) add8(g, rd[u][ci], 1.f);
            if constexpr (HAS_POOL || HAS_HEAD) {
                const int n = pix[u] / HW, rem = pix[u] - n * HW;
                if constexpr (HAS_POOL) {
                    // consumer at (H/2, W/2): route to the first maximum of the 2x2 window of a = relu(scale*y+shift)
                    const int yy = rem / p.W, xx = rem - yy * p.W;
                    const GradCons& c = p.cons[NS];
                    const int y0 = yy & ~1, x0 = xx & ~1;
                    const int me = ((yy & 1) << 1) | (xx & 1);
                    const size_t wbase = (size_t)n * HW + (size_t)y0 * p.W + x0;
                    float am[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) am[i] = fmaxf(fmaf(yv[i], sc[i], sh[i]), 0.f);
                    unsigned win = 0xFFu;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (k == me) continue;
                        float t[8];
                        unpack8f(__ldg(reinterpret_cast<const uint4*>(p.y + (wbase + (size_t)(k >> 1) * p.W + (k & 1)) * p.C + ch)), t);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float ak = fmaxf(fmaf(t[i], sc[i], sh[i]), 0.f);
                            const bool keep = (k < me) ? (am[i] > ak) : (am[i] >= ak);
                            if (!keep) win &= ~(1u << i);
                        }
                    }
                    const size_t cpix = ((size_t)n * (p.H >> 1) + (yy >> 1)) * (p.W >> 1) + (xx >> 1);
                    float d[8];
                    unpack8f(__ldg(reinterpret_cast<const uint4*>(c.dx + cpix * c.pitch + c.coff + ch)), d);
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        if (win & (1u << i)) g[i] += d[i];
                }
                if constexpr (HAS_HEAD) {
                    for (int k = 0; k < p.K; ++k) {
                        const float dl = __ldg(p.dlogits + ((size_t)n * p.K + k) * HW + rem);
#pragma unroll
                        for (int i = 0; i < 8; ++i) g[i] = fmaf(dl, __ldg(p.wh + k * p.C + ch + i), g[i]);
                    }
                }
            }
            uint32_t o[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float z = fmaf(yv[i], sc[i], sh[i]);
                const float gz = (z > 0.f) ? g[i] : 0.f;
                g[i] = gz;
                acc1[i] += gz;
                acc2[i] = fmaf(gz, yv[i], acc2[i]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = __floats2bfloat162_rn(g[2 * i], g[2 * i + 1]);
                o[i] = *reinterpret_cast<uint32_t*>(&h);
            }
            *reinterpret_cast<uint4*>(p.gz + (size_t)pix[u] * p.C + ch) = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = p.invstd[ch + i] * (acc2[i] - p.mean[ch + i] * acc1[i]);
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
    if (p.tail.counter) bn_tail_run(p.tail, p.part, (int)gridDim.x, p.C, gridDim.x, threadIdx.x, 0, &red[0][0]);
}

// ---- v2 (round 2): every consumer slice is now a DENSE tensor (the dgrads write one tensor per source), so
//  * POOL: a thread owns a whole 2x2 window (4 pixels x 8 channels): the window's y values are loaded once (round 1: every pixel
//    re-read its three neighbours), the arg-max is taken in registers, one pooled-gradient load serves four outputs;
//  * HEAD (K == 1, the reference's default num_classes): the head weights live in registers and the head's own parameter gradients
//    (dWh = sum dl*a, dbh = sum dl; round 1: a separate pass over y, nunet_head_bwd_params) are accumulated in the same pass and
//    finished by the last CTA together with the BatchNorm-backward coefficients.
template <int NS, bool POOL, bool HEAD>
__global__ void __launch_bounds__(256, POOL ? (NS >= 3 ? 1 : 2) : 3) grad_router_v2_kernel(const RouterParams p) {
    const int cgs = p.C >> 3;
    const int cg = threadIdx.x & (cgs - 1);      // C/8 is a power of two; grid strides are multiples of it
    const int ch = cg * 8;
    const int HW = p.H * p.W;
    float sc[8], sh[8], acc1[8], acc2[8], wh8[8], accw[8], accb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = p.scale[ch + i];
        sh[i] = p.shift[ch + i];
        acc1[i] = 0.f;
        acc2[i] = 0.f;
        wh8[i] = HEAD ? p.wh[ch + i] : 0.f;
        accw[i] = 0.f;
    }
    // one pixel: g (gathered gradient of the activation), yv (raw y) -> g_z store + BatchNorm sums (+ head parameter sums)
    auto finish_pixel = [&](size_t pix, const float (&yv)[8], float (&g)[8], float dl) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float z = fmaf(yv[i], sc[i], sh[i]);
            if (HEAD) {
                g[i] = fmaf(dl, wh8[i], g[i]);
                accw[i] = fmaf(dl, fmaxf(z, 0.f), accw[i]);
            }
            const float gz = (z > 0.f) ? g[i] : 0.f;
            g[i] = gz;
            acc1[i] += gz;
            acc2[i] = fmaf(gz, yv[i], acc2[i]);
        }
        if (HEAD && cg == 0) accb += dl;
        *reinterpret_cast<uint4*>(p.gz + pix * p.C + ch) = pack8(g);
    };
    const int stride = gridDim.x * 256;
    if constexpr (POOL) {
        const int h = p.H >> 1, w = p.W >> 1;
        const int items = p.N * h * w * cgs;
        const int cg_shift = 31 - __clz(cgs);
        const GradCons& pc = p.cons[NS];
        for (int it = blockIdx.x * 256 + threadIdx.x; it < items; it += stride) {
            const int win = it >> cg_shift;
            const int x = win % w, t2 = win / w;
            const int yy = t2 % h, n = t2 / h;
            const size_t p00 = ((size_t)n * p.H + 2 * yy) * p.W + 2 * x;
            const size_t off[4] = {0, 1, (size_t)p.W, (size_t)p.W + 1};
            uint4 ry[4], rd[4][NS > 0 ? NS : 1];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                ry[k] = __ldg(reinterpret_cast<const uint4*>(p.y + (p00 + off[k]) * p.C + ch));
#pragma unroll
                for (int ci = 0; ci < NS; ++ci)
                    rd[k][ci] = __ldg(reinterpret_cast<const uint4*>(p.cons[ci].dx + (p00 + off[k]) * p.cons[ci].pitch + p.cons[ci].coff + ch));
            }
            const uint4 rp = __ldg(reinterpret_cast<const uint4*>(pc.dx + (size_t)win * pc.pitch + pc.coff + ch));
            float yv[4][8], d[8];
            unpack8f(rp, d);
            int idx[8];
            float best[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unpack8f(ry[k], yv[k]);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float am = fmaxf(fmaf(yv[k][i], sc[i], sh[i]), 0.f);
                    if (k == 0 || am > best[i]) { best[i] = am; idx[i] = k; }   // first maximum in window scan order, as ATen
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float g[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = (idx[i] == k) ? d[i] : 0.f;
#pragma unroll
                for (int ci = 0; ci < NS; ++ci) add8(g, rd[k][ci], 1.f);
                float dl = 0.f;
                if (HEAD) dl = __ldg(p.dlogits + (size_t)n * HW + (size_t)(2 * yy + (k >> 1)) * p.W + 2 * x + (k & 1));
                finish_pixel(p00 + off[k], yv[k], g, dl);
            }
        }
    } else {
        const int items = p.N * HW * cgs;
        const int cg_shift = 31 - __clz(cgs);
        constexpr int U = (NS <= 2) ? 2 : 1;
        for (int it0 = blockIdx.x * 256 + threadIdx.x; it0 < items; it0 += stride * U) {
            uint4 ry[U], rd[U][NS > 0 ? NS : 1];
            int pix[U];
            float dl[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int it = it0 + u * stride;
                pix[u] = (it < items) ? (it >> cg_shift) : -1;
                dl[u] = 0.f;
                if (pix[u] >= 0) {
                    ry[u] = __ldg(reinterpret_cast<const uint4*>(p.y + (size_t)pix[u] * p.C + ch));
#pragma unroll
                    for (int ci = 0; ci < NS; ++ci)
                        rd[u][ci] = __ldg(reinterpret_cast<const uint4*>(p.cons[ci].dx + (size_t)pix[u] * p.cons[ci].pitch + p.cons[ci].coff + ch));
                    if (HEAD) dl[u] = __ldg(p.dlogits + pix[u]);     // K == 1: dlogits [N,1,H,W] is indexed by the pixel
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (pix[u] < 0) continue;
                float yv[8], g[8];
                unpack8f(ry[u], yv);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = 0.f;
#pragma unroll
                for (int ci = 0; ci < NS; ++ci) add8(g, rd[u][ci], 1.f);
                finish_pixel((size_t)pix[u], yv, g, dl[u]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = p.invstd[ch + i] * (acc2[i] - p.mean[ch + i] * acc1[i]);
    // block reduction: threads with equal cg (tid % cgs) hold the same 8 channels
    __shared__ float red[256][9];
    for (int pass = 0; pass < (HEAD ? 3 : 2); ++pass) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[threadIdx.x][i] = pass == 0 ? acc1[i] : (pass == 1 ? acc2[i] : accw[i]);
        if (pass == 2) red[threadIdx.x][8] = accb;
        __syncthreads();
        for (int c = threadIdx.x; c < p.C + (pass == 2 ? 1 : 0); c += 256) {
            float s = 0.f;
            if (c < p.C) {
                const int g8 = c >> 3, e = c & 7;
                for (int t = g8; t < 256; t += cgs) s += red[t][e];
            } else {
                for (int t = 0; t < 256; t += cgs) s += red[t][8];   // only the cg == 0 threads accumulated dbh
            }
            if (pass < 2) p.part[((size_t)blockIdx.x * 2 + pass) * p.C + c] = s;
            else p.hpart[(size_t)blockIdx.x * 33 + c] = s;
        }
        __syncthreads();
    }
    if (p.tail.counter) {
        const bool last = bn_tail_run(p.tail, p.part, (int)gridDim.x, p.C, gridDim.x, threadIdx.x, 0, &red[0][0]);
        if (HEAD && last) {
            // head parameter gradients: 33 columns x gridDim.x rows; 7 row lanes per column, eight loads in flight, fixed order
            __syncthreads();
            double* shd = reinterpret_cast<double*>(&red[0][0]);   // [7][33]
            const int col = threadIdx.x % 33, lane = threadIdx.x / 33;
            if (lane < 7) {
                double a = 0.0;
                for (int r0 = lane; r0 < (int)gridDim.x; r0 += 7 * 8) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = r0 + u * 7;
                        v[u] = (r < (int)gridDim.x) ? __ldcg(p.hpart + (size_t)r * 33 + col) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) a += (double)v[u];
                }
                shd[lane * 33 + col] = a;
            }
            __syncthreads();
            if (threadIdx.x < 33) {
                double a = 0.0;
                for (int l = 0; l < 7; ++l) a += shd[l * 33 + threadIdx.x];
                if (threadIdx.x < 32) p.dwh[threadIdx.x] = (float)a;
                else if (p.dbh) p.dbh[0] = (float)a;
            }
        }
    }
}

// ---- bulk-copy variant for same-resolution consumers with dense rows (every router of the level-0 heads and the decoder nodes).
// The register kernels above are latency-bound at few consumers: 768 threads x (NS + 1) x U 16-byte loads is all an SM has in
// flight (x0_4's router: 24 KB, 2.6 TB/s).  Here a producer warp streams 4 KB chunks of y, of every consumer's dX and of dlogits
// into a ring of shared-memory stages with cp.async.bulk (up to 96 KB in flight per CTA, two CTAs per SM); a consumer thread owns
// one 16-byte item of every stream per chunk — its channel group is fixed (256 % (C/8) == 0), so the sums stay in registers.
constexpr uint32_t kRChunk = 4096;

template <int NS, bool HEAD>
__global__ void __launch_bounds__(288, 2) grad_router_bulk_kernel(const RouterParams p, int S, int rows_total) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ float red[256][9];
    const int cgs = p.C >> 3;
    const int cg_shift = 31 - __clz(cgs);
    const int PIX = 256 >> cg_shift;                       // pixels per chunk
    const int npix = p.N * p.H * p.W;
    const int nchunks = (npix + PIX - 1) / PIX;
    constexpr uint32_t kStreams = NS + 1;
    constexpr uint32_t stage_bytes = kStreams * kRChunk + (HEAD ? 256u : 0u);   // + dlogits: <= 64 floats
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
    uint64_t* empty = full + S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 8);
        }
        mbar_init_fence();
    }
    __syncthreads();
    if (warp == 8) {
        if (elect_one()) {
            int s = 0;
            uint32_t ph = 1;
            for (int ck = blockIdx.x; ck < nchunks; ck += gridDim.x) {
                const int p0 = ck * PIX;
                const int np = min(PIX, npix - p0);
                const uint32_t bytes = (uint32_t)np * (uint32_t)p.C * 2u;
                const size_t goff = (size_t)p0 * p.C * 2;
                uint8_t* st = smem + (size_t)s * stage_bytes;
                mbar_wait(&empty[s], ph);
                mbar_arrive_expect_tx(&full[s], bytes * kStreams + (HEAD ? (uint32_t)np * 4u : 0u));
                bulk_g2s(st, reinterpret_cast<const uint8_t*>(p.y) + goff, bytes, &full[s]);
#pragma unroll
                for (int ci = 0; ci < NS; ++ci)
                    bulk_g2s(st + (ci + 1) * kRChunk, reinterpret_cast<const uint8_t*>(p.cons[ci].dx) + goff, bytes, &full[s]);
                if (HEAD) bulk_g2s(st + kStreams * kRChunk, p.dlogits + p0, (uint32_t)np * 4u, &full[s]);
                if (++s == S) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    const int cg = tid & (cgs - 1), lp = tid >> cg_shift;
    const int ch = cg * 8;
    float sc[8], sh[8], acc1[8], acc2[8], wh8[8], accw[8], accb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = p.scale[ch + i];
        sh[i] = p.shift[ch + i];
        acc1[i] = 0.f;
        acc2[i] = 0.f;
        wh8[i] = HEAD ? p.wh[ch + i] : 0.f;
        accw[i] = 0.f;
    }
    {
        int s = 0;
        uint32_t ph = 0;
        const uint8_t* st = smem;
        for (int ck = blockIdx.x; ck < nchunks; ck += gridDim.x) {
            const int p0 = ck * PIX;
            const int np = min(PIX, npix - p0);
            mbar_wait(&full[s], ph);
            const uint4 ry = *reinterpret_cast<const uint4*>(st + tid * 16);
            uint4 rd[NS > 0 ? NS : 1];
#pragma unroll
            for (int ci = 0; ci < NS; ++ci) rd[ci] = *reinterpret_cast<const uint4*>(st + (ci + 1) * kRChunk + tid * 16);
            float dl = 0.f;
            if (HEAD) dl = *reinterpret_cast<const float*>(st + kStreams * kRChunk + lp * 4);
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);          // this warp holds its items in registers
            st += stage_bytes;
            if (++s == S) { s = 0; ph ^= 1u; st = smem; }
            if (lp < np) {
                float yv[8], g[8];
                unpack8f(ry, yv);
#pragma unroll
                for (int i = 0; i < 8; ++i) g[i] = 0.f;
#pragma unroll
                for (int ci = 0; ci < NS; ++ci) add8(g, rd[ci], 1.f);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float z = fmaf(yv[i], sc[i], sh[i]);
                    if (HEAD) {
                        g[i] = fmaf(dl, wh8[i], g[i]);
                        accw[i] = fmaf(dl, fmaxf(z, 0.f), accw[i]);
                    }
                    const float gz = (z > 0.f) ? g[i] : 0.f;
                    g[i] = gz;
                    acc1[i] += gz;
                    acc2[i] = fmaf(gz, yv[i], acc2[i]);
                }
                if (HEAD && cg == 0) accb += dl;
                *reinterpret_cast<uint4*>(p.gz + (size_t)(p0 + lp) * p.C + ch) = pack8(g);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc2[i] = p.invstd[ch + i] * (acc2[i] - p.mean[ch + i] * acc1[i]);
    // block reduction over the 256 consumer threads (named barrier 1: the producer warp has left)
    for (int pass = 0; pass < (HEAD ? 3 : 2); ++pass) {
#pragma unroll
        for (int i = 0; i < 8; ++i) red[tid][i] = pass == 0 ? acc1[i] : (pass == 1 ? acc2[i] : accw[i]);
        if (pass == 2) red[tid][8] = accb;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int c = tid; c < p.C + (pass == 2 ? 1 : 0); c += 256) {
            float sum = 0.f;
            if (c < p.C) {
                const int g8 = c >> 3, e = c & 7;
                for (int t = g8; t < 256; t += cgs) sum += red[t][e];
            } else {
                for (int t = 0; t < 256; t += cgs) sum += red[t][8];
            }
            if (pass < 2) p.part[((size_t)blockIdx.x * 2 + pass) * p.C + c] = sum;
            else p.hpart[(size_t)blockIdx.x * 33 + c] = sum;
            // the caller's buffers hold nunet_grad_router_rows() rows (sized for the register kernels' grids): the rest is zero
            for (int r = blockIdx.x + gridDim.x; r < rows_total; r += gridDim.x) {
                if (pass < 2) p.part[((size_t)r * 2 + pass) * p.C + c] = 0.f;
                else p.hpart[(size_t)r * 33 + c] = 0.f;
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (p.tail.counter) {
        const bool last = bn_tail_run(p.tail, p.part, (int)gridDim.x, p.C, gridDim.x, tid, 1, &red[0][0]);
        if (HEAD && last) {
            asm volatile("bar.sync 1, 256;" ::: "memory");
            double* shd = reinterpret_cast<double*>(&red[0][0]);   // [7][33]
            const int col = tid % 33, rl = tid / 33;
            if (rl < 7) {
                double a = 0.0;
                for (int r0 = rl; r0 < (int)gridDim.x; r0 += 7 * 8) {
                    float v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int r = r0 + u * 7;
                        v[u] = (r < (int)gridDim.x) ? __ldcg(p.hpart + (size_t)r * 33 + col) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) a += (double)v[u];
                }
                shd[rl * 33 + col] = a;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (tid < 33) {
                double a = 0.0;
                for (int l = 0; l < 7; ++l) a += shd[l * 33 + tid];
                if (tid < 32) p.dwh[tid] = (float)a;
                else if (p.dbh) p.dbh[0] = (float)a;
            }
        }
    }
}

template <int NS, bool HEAD>
static int launch_router_bulk(const RouterParams& p, int rows, cudaStream_t st) {
    constexpr uint32_t stage_bytes = (NS + 1) * kRChunk + (HEAD ? 256u : 0u);
    int S = (int)(96u * 1024u / stage_bytes);
    if (S > 8) S = 8;
    const size_t smem = (size_t)S * stage_bytes + 2 * S * 8;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(grad_router_bulk_kernel<NS, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
        attr_done = true;
    }
    const int cgs = p.C >> 3, PIX = 256 / cgs;
    const long long nchunks = ((long long)p.N * p.H * p.W + PIX - 1) / PIX;
    int grid = rows < 2 * 148 ? rows : 2 * 148;
    if (nchunks < grid) grid = (int)nchunks;
    grad_router_bulk_kernel<NS, HEAD><<<grid, 288, smem, st>>>(p, S, rows);
    cudaError_t e0 = cudaGetLastError();
    return e0 == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e0;
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_grad_router_rows(int32_t N, int32_t H, int32_t W, int32_t C) {
    long long items = (long long)N * H * W * (C / 8);
    long long b = (items + 255) / 256;
    // the last CTA reduces rows x 2C partial sums in its tail: fewer rows for the wide (and small) deep layers
    const long long cap = C <= 64 ? 148 * 4 : (C <= 128 ? 148 * 2 : 148);
    if (b > cap) b = cap;
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
    int n_same = 0, n_pool = 0, n_up = 0;
    for (int pass = 0; pass < 3; ++pass) {  // stable reorder: SAME..., POOL, UP
        const int want = pass == 0 ? NUNET_SRC_SAME : (pass == 1 ? NUNET_SRC_POOL : NUNET_SRC_UP);
        for (int i = 0; i < d->num_cons; ++i) {
            const nunet_grad_cons_t& c = d->cons[i];
            if (!c.dx || c.pitch % 8 != 0 || c.coff % 8 != 0) return NUNET_E_BADARG;
            if (c.kind != NUNET_SRC_SAME && c.kind != NUNET_SRC_POOL && c.kind != NUNET_SRC_UP) return NUNET_E_BADARG;
            if (c.kind != want) continue;
            if (c.kind == NUNET_SRC_POOL && ((d->H | d->W) & 1)) return NUNET_E_UNSUPPORTED;
            const int j = n_same + n_pool + n_up;
            p.cons[j].dx = (const __nv_bfloat16*)c.dx; p.cons[j].pitch = c.pitch; p.cons[j].coff = c.coff; p.cons[j].kind = c.kind;
            if (pass == 0) ++n_same; else if (pass == 1) ++n_pool; else ++n_up;
        }
    }
    if (n_same > 4 || n_pool > 1 || n_up > 1) return NUNET_E_UNSUPPORTED;
    if ((long long)d->N * d->H * d->W * (d->C / 8) >= (1LL << 31) - (1LL << 22)) return NUNET_E_UNSUPPORTED;
    p.dlogits = d->dlogits; p.wh = d->wh; p.K = d->K;
    p.gz = (__nv_bfloat16*)d->gz; p.part = d->part;
    if (d->bn.counter) {
        const nunet_bn_bwd_tail_t& b = d->bn;
        if (b.count <= 0 || (b.gamma ? (!b.save_mean || !b.save_invstd || !b.dgamma || !b.dbeta || !b.k1 || !b.k2 || !b.k3) : !b.sums_out)) return NUNET_E_BADARG;
        p.tail.counter = b.counter; p.tail.sums_out = b.sums_out;
        p.tail.fin.mode = 1; p.tail.fin.count = b.count; p.tail.fin.gamma = b.gamma;
        p.tail.fin.save_mean = const_cast<float*>(b.save_mean); p.tail.fin.save_invstd = const_cast<float*>(b.save_invstd);
        p.tail.fin.dgamma = b.dgamma; p.tail.fin.dbeta = b.dbeta; p.tail.fin.k1 = b.k1; p.tail.fin.k2 = b.k2; p.tail.fin.k3 = b.k3;
        p.tail.fin.grad_scale = b.grad_scale;
        if (!exchange_ok(b.xch)) return NUNET_E_BADARG;
        copy_exchange(p.tail.xch, b.xch);
    }
    const int rows = nunet_grad_router_rows(d->N, d->H, d->W, d->C);
    const bool hp = n_pool > 0, hu = n_up > 0, hh = d->dlogits != nullptr;
    cudaStream_t st = (cudaStream_t)stream;
    static const bool no_fast = getenv("NUNET_ROUTER_GENERIC") != nullptr;
    static const bool no_v2 = getenv("NUNET_ROUTER_V1") != nullptr;
    const bool fused_head = d->head_part != nullptr;
    if (fused_head && (!hh || d->K != 1 || d->C != 32 || !d->dwh || !d->bn.counter || hu || hp)) return NUNET_E_BADARG;
    p.hpart = d->head_part; p.dwh = d->dwh; p.dbh = d->dbh;
    // bulk-copy streaming kernel: same-resolution consumers only, dense rows, no head or the fused K == 1 head
    static const bool no_bulk = getenv("NUNET_ROUTER_NO_BULK") != nullptr;
    if (!no_bulk && !no_fast && !hu && !hp && (!hh || fused_head) && ((uintptr_t)d->y & 15) == 0) {
        bool dense = true;
        for (int i = 0; i < n_same; ++i) dense = dense && p.cons[i].pitch == d->C && p.cons[i].coff == 0 && ((uintptr_t)p.cons[i].dx & 15) == 0;
        if (hh) dense = dense && ((long long)d->N * d->H * d->W) % 4 == 0 && ((uintptr_t)d->dlogits & 15) == 0;
        if (dense) {
#define NUNET_BULK_CASE(NS_) if (n_same == NS_) return hh ? launch_router_bulk<NS_, true>(p, rows, st) : launch_router_bulk<NS_, false>(p, rows, st)
            NUNET_BULK_CASE(0); NUNET_BULK_CASE(1); NUNET_BULK_CASE(2); NUNET_BULK_CASE(3); NUNET_BULK_CASE(4);
#undef NUNET_BULK_CASE
        }
    }
    // v2 where it wins: a POOL consumer (window per thread) or the fused head; plain same-resolution sums stay on the 64-register kernel
    if (!hu && !no_fast && !no_v2 && (hp || fused_head) && !(hp && hh)) {
#define NUNET_V2_CASE(NS_, P_, H_) if (n_same == NS_ && hp == P_ && hh == H_) grad_router_v2_kernel<NS_, P_, H_><<<rows, 256, 0, st>>>(p)
        NUNET_V2_CASE(0, false, true); NUNET_V2_CASE(1, false, false); NUNET_V2_CASE(1, false, true); NUNET_V2_CASE(2, false, false);
        NUNET_V2_CASE(2, false, true); NUNET_V2_CASE(3, false, false); NUNET_V2_CASE(3, false, true); NUNET_V2_CASE(4, false, false);
        NUNET_V2_CASE(0, true, false); NUNET_V2_CASE(1, true, false); NUNET_V2_CASE(2, true, false); NUNET_V2_CASE(3, true, false);
        NUNET_V2_CASE(4, true, false); NUNET_V2_CASE(0, false, false);
#undef NUNET_V2_CASE
        cudaError_t e0 = cudaGetLastError();
        return e0 == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e0;
    }
    if (fused_head) return NUNET_E_UNSUPPORTED;
    if (!hu && !no_fast) {
#define NUNET_FAST_CASE(NS_, P_, H_) if (n_same == NS_ && hp == P_ && hh == H_) grad_router_fast_kernel<NS_, P_, H_><<<rows, 256, 0, st>>>(p)
#define NUNET_FAST_NS(NS_) NUNET_FAST_CASE(NS_, false, false); NUNET_FAST_CASE(NS_, false, true); NUNET_FAST_CASE(NS_, true, false); NUNET_FAST_CASE(NS_, true, true)
        NUNET_FAST_NS(0); NUNET_FAST_NS(1); NUNET_FAST_NS(2); NUNET_FAST_NS(3); NUNET_FAST_NS(4);
#undef NUNET_FAST_NS
#undef NUNET_FAST_CASE
        cudaError_t e0 = cudaGetLastError();
        return e0 == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e0;
    }
#define NUNET_ROUTER_CASE(P_, U_, H_) if (hp == P_ && hu == U_ && hh == H_) grad_router_kernel<P_, U_, H_><<<rows, 256, 0, st>>>(p)
    NUNET_ROUTER_CASE(false, false, false);
    NUNET_ROUTER_CASE(false, false, true);
    NUNET_ROUTER_CASE(false, true, false);
    NUNET_ROUTER_CASE(false, true, true);
    NUNET_ROUTER_CASE(true, false, false);
    NUNET_ROUTER_CASE(true, false, true);
    NUNET_ROUTER_CASE(true, true, false);
    NUNET_ROUTER_CASE(true, true, true);
#undef NUNET_ROUTER_CASE
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
