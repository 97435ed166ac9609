// fp32 CHECK MODE (forward and backward): the same network graph evaluated with plain CUDA-core fp32 arithmetic, fp64 BatchNorm
// statistics and no tensor cores, so that the host logic (graph walk, concat order, BatchNorm semantics incl. running
// statistics, pooling, align_corners bilinear upsampling, heads) can be held against the reference's fp32 path to
// max|d|/max|ref| <= 1e-4 (north_star; SURVEY.md 8c finding 4: bf16 operands cannot meet that bound, two fp32
// implementations differ by ~5e-6 from summation order alone).  It is a verification path selected explicitly
// (engine.CHECK_FP32 / NUNET_FP32_CHECK=1), never a fallback: the product path still fails loudly without the tcgen05 kernels.
//
// Layouts: activations fp32 NHWC with an explicit channel pitch (a concat buffer is written slice by slice by its
// producers' copy kernels), weights repacked to [tap][ci][co] (coalesced over co), logits fp32 NCHW.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/nunet_b200.h"

namespace nunet {

// y[n,y,x,co] = bias[co] + sum_{tap,ci} x[n,y+dy,x+dx,ci] * wt[tap][ci][co]   (nn.Conv2d(3, padding=1), finished/archs1.py:18,20)
// one thread per (pixel, co); a warp = 32 output channels of one pixel (x is a broadcast load, wt is coalesced).
__global__ void __launch_bounds__(256) check_conv3x3_kernel(const float* __restrict__ x, int N, int H, int W, int Cin, int xpitch,
                                                            const float* __restrict__ wt, const float* __restrict__ bias, int Cout,
                                                            float* __restrict__ y) {
    const long long total = (long long)N * H * W * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long pix = i / Cout;
        const int px = (int)(pix % W); pix /= W;
        const int py = (int)(pix % H);
        const int n = (int)(pix / H);
        float acc = bias ? bias[co] : 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float* xr = x + (((size_t)n * H + yy) * W + xx) * xpitch;
            const float* wr = wt + (size_t)tap * Cin * Cout + co;
            for (int ci = 0; ci < Cin; ++ci) acc = fmaf(xr[ci], wr[(size_t)ci * Cout], acc);
        }
        y[i] = acc;
    }
}

// w OIHW [Cout][Cin][3][3] -> wt [tap][ci][co]
__global__ void check_repack_w_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ wt) {
    const int total = Cout * Cin * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int co = i % Cout, ci = (i / Cout) % Cin, tap = i / (Cout * Cin);
        wt[i] = w[((size_t)co * Cin + ci) * 9 + tap];
    }
}

// Train-mode BatchNorm statistics of y [M][C] in fp64: one block per channel.  Emits scale/shift of
// a = relu(scale*y + shift) and updates running_mean / running_var (unbiased) like nn.BatchNorm2d (finished/archs1.py:19,21).
__global__ void __launch_bounds__(256) check_bn_stats_kernel(const float* __restrict__ y, long long M, int C, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* running_mean, float* running_var,
                                                             float momentum, float eps, float* __restrict__ scale, float* __restrict__ shift,
                                                             float* save_mean, float* save_invstd) {
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0;
    for (long long i = threadIdx.x; i < M; i += blockDim.x) {
        const double v = (double)y[i * C + c];
        s += v;
        q += v * v;
    }
    __shared__ double rs[256], rq[256];
    rs[threadIdx.x] = s; rq[threadIdx.x] = q;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { rs[threadIdx.x] += rs[threadIdx.x + off]; rq[threadIdx.x] += rq[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double mean = rs[0] / (double)M;
        double var = rq[0] / (double)M - mean * mean;
        if (var < 0.0) var = 0.0;
        const double invstd = 1.0 / sqrt(var + (double)eps);
        const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
        scale[c] = (float)(g * invstd);
        shift[c] = (float)(b - mean * g * invstd);
        if (save_mean) save_mean[c] = (float)mean;
        if (save_invstd) save_invstd[c] = (float)invstd;
        if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
        if (running_var) {
            const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
            running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unb);
        }
    }
}

// eval mode: scale/shift from the running statistics
__global__ void check_bn_eval_kernel(int C, const float* gamma, const float* beta, const float* rm, const float* rv, float eps, float* scale,
                                     float* shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double invstd = 1.0 / sqrt((double)rv[c] + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    scale[c] = (float)(g * invstd);
    shift[c] = (float)(b - (double)rm[c] * g * invstd);
}

// dst[m][off + c] = relu(scale[c]*y[m][c] + shift[c])   (dst pitch may exceed C: slice of a concat buffer)
__global__ void check_bn_relu_kernel(const float* __restrict__ y, long long M, int C, const float* __restrict__ scale,
                                     const float* __restrict__ shift, float* __restrict__ dst, int pitch, int off) {
    const long long total = M * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long m = i / C;
        dst[m * pitch + off + c] = fmaxf(fmaf(y[i], scale[c], shift[c]), 0.f);
    }
}

// mode 0: copy, 1: 2x2 max-pool (nn.MaxPool2d(2,2)), 2: bilinear 2x upsample, align_corners=True (nn.Upsample, archs1.py:41-42)
// src [N,h,w,C] dense -> dst [N,H,W,pitch] slice at `off` (H,W = h,w | h/2,w/2 | 2h,2w)
__global__ void check_route_kernel(const float* __restrict__ src, int N, int h, int w, int C, int mode, float* __restrict__ dst, int pitch,
                                   int off) {
    const int H = mode == 1 ? h / 2 : (mode == 2 ? 2 * h : h), W = mode == 1 ? w / 2 : (mode == 2 ? 2 * w : w);
    const long long total = (long long)N * H * W * C;
    const float ry = (mode == 2 && H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = (mode == 2 && W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long pix = i / C;
        const int X = (int)(pix % W); pix /= W;
        const int Y = (int)(pix % H);
        const int n = (int)(pix / H);
        const float* s = src + (size_t)n * h * w * C + c;
        float v;
        if (mode == 0) {
            v = s[((size_t)Y * w + X) * C];
        } else if (mode == 1) {
            const float a = s[((size_t)(2 * Y) * w + 2 * X) * C], b = s[((size_t)(2 * Y) * w + 2 * X + 1) * C];
            const float d = s[((size_t)(2 * Y + 1) * w + 2 * X) * C], e = s[((size_t)(2 * Y + 1) * w + 2 * X + 1) * C];
            v = fmaxf(fmaxf(a, b), fmaxf(d, e));
        } else {
            const float sy = ry * (float)Y, sx = rx * (float)X;
            const int y0 = (int)sy, x0 = (int)sx;
            const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
            const float ly = sy - (float)y0, lx = sx - (float)x0;
            const float a = s[((size_t)y0 * w + x0) * C], b = s[((size_t)y0 * w + x1) * C];
            const float d = s[((size_t)y1 * w + x0) * C], e = s[((size_t)y1 * w + x1) * C];
            v = (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * d + lx * e);
        }
        dst[(((size_t)n * H + Y) * W + X) * pitch + off + c] = v;
    }
}

// network input fp32 NCHW -> fp32 NHWC
__global__ void check_nchw_to_nhwc_kernel(const float* __restrict__ x, int N, int C, int HW, float* __restrict__ out) {
    const long long total = (long long)N * C * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW, hw = pix % HW;
        out[i] = x[((size_t)n * C + c) * HW + hw];
    }
}

// 1x1 head (finished/archs1.py:106-111): logits[n,k,hw] = bh[k] + sum_c a[n,hw,c] * wh[k][c]
__global__ void check_head_kernel(const float* __restrict__ a, int N, int HW, int C, const float* __restrict__ wh, const float* __restrict__ bh,
                                  int K, float* __restrict__ logits) {
    const long long total = (long long)N * K * HW;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int hw = (int)(i % HW);
        const int k = (int)((i / HW) % K);
        const long long n = i / ((long long)HW * K);
        const float* ar = a + ((size_t)n * HW + hw) * C;
        float acc = bh ? bh[k] : 0.f;
        for (int c = 0; c < C; ++c) acc = fmaf(ar[c], wh[(size_t)k * C + c], acc);
        logits[i] = acc;
    }
}


// ------------------------------------------------------------------------------------------------ backward (autograd of the above)
// ReLU + train-mode BatchNorm backward (autograd of finished/archs1.py:25-26,29-30), pass 1: one block per channel, fp64.
//   g_z = g_a * [scale*y+shift > 0];  dbeta = S1 = sum g_z;  dgamma = S2 = sum g_z * yhat,  yhat = (y - mean) * invstd
__global__ void __launch_bounds__(256) check_bn_bwd_stats_kernel(const float* __restrict__ ga, const float* __restrict__ y, long long M, int C,
                                                                 const float* __restrict__ scale, const float* __restrict__ shift,
                                                                 const float* __restrict__ mean, const float* __restrict__ invstd,
                                                                 float* __restrict__ dgamma, float* __restrict__ dbeta) {
    const int c = blockIdx.x;
    const float sc = scale[c], sh = shift[c];
    const double mu = mean[c], r = invstd[c];
    double s1 = 0.0, s2 = 0.0;
    for (long long i = threadIdx.x; i < M; i += blockDim.x) {
        const float yv = y[i * C + c];
        const double gz = (fmaf(yv, sc, sh) > 0.f) ? (double)ga[i * C + c] : 0.0;
        s1 += gz;
        s2 += gz * (((double)yv - mu) * r);
    }
    __shared__ double r1[256], r2[256];
    r1[threadIdx.x] = s1; r2[threadIdx.x] = s2;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) { r1[threadIdx.x] += r1[threadIdx.x + off]; r2[threadIdx.x] += r2[threadIdx.x + off]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { dbeta[c] = (float)r1[0]; dgamma[c] = (float)r2[0]; }
}
// pass 2: g_y = gamma*invstd*(g_z - S1/M - yhat*S2/M)
__global__ void check_bn_bwd_apply_kernel(const float* __restrict__ ga, const float* __restrict__ y, long long M, int C,
                                          const float* __restrict__ scale, const float* __restrict__ shift, const float* __restrict__ gamma,
                                          const float* __restrict__ mean, const float* __restrict__ invstd, const float* __restrict__ dgamma,
                                          const float* __restrict__ dbeta, float* __restrict__ gy) {
    const long long total = M * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const float yv = y[i];
        const double gz = (fmaf(yv, scale[c], shift[c]) > 0.f) ? (double)ga[i] : 0.0;
        const double r = invstd[c], yhat = ((double)yv - (double)mean[c]) * r;
        const double g = gamma ? (double)gamma[c] : 1.0;
        gy[i] = (float)(g * r * (gz - (double)dbeta[c] / (double)M - yhat * (double)dgamma[c] / (double)M));
    }
}

// dx[n,y,x,ci] = sum_{tap,co} gy[n,y-dy,x-dx,co] * w[co][ci][tap]   (autograd of nn.Conv2d w.r.t. its input); one thread per (pixel, ci)
// wd: [tap][co][ci] (coalesced over ci)
__global__ void __launch_bounds__(256) check_dgrad_kernel(const float* __restrict__ gy, int N, int H, int W, int Cout, const float* __restrict__ wd,
                                                          int Cin, float* __restrict__ dx) {
    const long long total = (long long)N * H * W * Cin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        long long pix = i / Cin;
        const int px = (int)(pix % W); pix /= W;
        const int py = (int)(pix % H);
        const int n = (int)(pix / H);
        float acc = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
            const int yy = py - (tap / 3 - 1), xx = px - (tap % 3 - 1);   // the output pixel that read us through this tap
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float* gr = gy + (((size_t)n * H + yy) * W + xx) * Cout;
            const float* wr = wd + (size_t)tap * Cout * Cin + ci;
            for (int co = 0; co < Cout; ++co) acc = fmaf(gr[co], wr[(size_t)co * Cin], acc);
        }
        dx[i] = acc;
    }
}
// w OIHW -> wd [tap][co][ci]
__global__ void check_repack_wd_kernel(const float* __restrict__ w, int Cout, int Cin, float* __restrict__ wd) {
    const int total = Cout * Cin * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int ci = i % Cin, co = (i / Cin) % Cout, tap = i / (Cout * Cin);
        wd[i] = w[((size_t)co * Cin + ci) * 9 + tap];
    }
}

// dw[co][ci][tap] = sum_pixels x[n,y+dy,x+dx,ci] * gy[n,y,x,co]   (autograd of nn.Conv2d w.r.t. the weight); fp64 accumulation.
// One block per (ci, tap) and 32 output channels; threads = 32 co x 8 pixel lanes.
__global__ void __launch_bounds__(256) check_wgrad_kernel(const float* __restrict__ x, int N, int H, int W, int Cin, int xpitch,
                                                          const float* __restrict__ gy, int Cout, float* __restrict__ dw) {
    const int tap = blockIdx.x % 9, ci = blockIdx.x / 9;
    const int co = blockIdx.y * 32 + (threadIdx.x & 31), pl = threadIdx.x >> 5;
    const int dy = tap / 3 - 1, dxx = tap % 3 - 1;
    const long long M = (long long)N * H * W;
    double acc = 0.0;
    if (co < Cout) {
        for (long long m = pl; m < M; m += 8) {
            const int px = (int)(m % W), py = (int)((m / W) % H);
            const long long n = m / ((long long)W * H);
            const int yy = py + dy, xx = px + dxx;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            acc += (double)x[(((size_t)n * H + yy) * W + xx) * xpitch + ci] * (double)gy[(size_t)m * Cout + co];
        }
    }
    __shared__ double red[8][33];
    red[pl][threadIdx.x & 31] = acc;
    __syncthreads();
    if (pl == 0 && co < Cout) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x & 31];
        dw[((size_t)co * Cin + ci) * 9 + tap] = (float)s;
    }
}
// dbias[co] = sum_pixels gy (analytically zero under train-mode BatchNorm; the reference's autograd still produces its round-off)
__global__ void __launch_bounds__(256) check_colsum_kernel(const float* __restrict__ g, long long M, int C, float* __restrict__ out) {
    const int c = blockIdx.x;
    double s = 0.0;
    for (long long i = threadIdx.x; i < M; i += blockDim.x) s += (double)g[i * C + c];
    __shared__ double rs[256];
    rs[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) rs[threadIdx.x] += rs[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[c] = (float)rs[0];
}

// Backward of check_route_kernel: ga[N,h,w,C] (dense, accumulated) += route^T(dx[N,H,W,pitch] slice at off).
// mode 0 copy | 1 max-pool (first maximum of the 2x2 window of `a` in scan order, as ATen) | 2 bilinear 2x, align_corners=True
__global__ void check_route_bwd_kernel(const float* __restrict__ dx, int N, int h, int w, int C, int mode, int pitch, int off,
                                       const float* __restrict__ a, float* __restrict__ ga) {
    const int H = mode == 1 ? h / 2 : (mode == 2 ? 2 * h : h), W = mode == 1 ? w / 2 : (mode == 2 ? 2 * w : w);
    const long long total = (long long)N * h * w * C;
    const float ry = (mode == 2 && H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = (mode == 2 && W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long pix = i / C;
        const int x = (int)(pix % w); pix /= w;
        const int y = (int)(pix % h);
        const int n = (int)(pix / h);
        const float* d = dx + (size_t)n * H * W * pitch + off + c;
        float g = 0.f;
        if (mode == 0) {
            g = d[((size_t)y * W + x) * pitch];
        } else if (mode == 1) {
            const int y0 = y & ~1, x0 = x & ~1, me = ((y & 1) << 1) | (x & 1);
            const float* ab = a + (size_t)n * h * w * C + c;
            const float am = ab[((size_t)y * w + x) * C];
            bool win = true;
            for (int k = 0; k < 4; ++k) {
                if (k == me) continue;
                const float ak = ab[((size_t)(y0 + (k >> 1)) * w + x0 + (k & 1)) * C];
                win = win && ((k < me) ? (am > ak) : (am >= ak));
            }
            if (win) g = d[((size_t)(y >> 1) * W + (x >> 1)) * pitch];
        } else {
            for (int Y = 2 * y - 2; Y <= 2 * y + 3; ++Y) {
                if (Y < 0 || Y >= H) continue;
                const float sy = ry * (float)Y;
                const int i0 = (int)sy, i1 = i0 + (i0 < h - 1);
                const float l1 = sy - (float)i0;
                const float wy = (i0 == y ? 1.f - l1 : 0.f) + (i1 == y ? l1 : 0.f);
                if (wy == 0.f) continue;
                for (int X = 2 * x - 2; X <= 2 * x + 3; ++X) {
                    if (X < 0 || X >= W) continue;
                    const float sx = rx * (float)X;
                    const int j0 = (int)sx, j1 = j0 + (j0 < w - 1);
                    const float m1 = sx - (float)j0;
                    const float wx = (j0 == x ? 1.f - m1 : 0.f) + (j1 == x ? m1 : 0.f);
                    if (wx != 0.f) g = fmaf(wy * wx, d[((size_t)Y * W + X) * pitch], g);
                }
            }
        }
        ga[i] += g;
    }
}

// Backward of the 1x1 head: ga[n,hw,c] += sum_k dl[n,k,hw]*wh[k][c]
__global__ void check_head_bwd_act_kernel(const float* __restrict__ dl, int N, int HW, int C, const float* __restrict__ wh, int K,
                                          float* __restrict__ ga) {
    const long long total = (long long)N * HW * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long pix = i / C;
        const long long n = pix / HW, hw = pix % HW;
        float g = 0.f;
        for (int k = 0; k < K; ++k) g = fmaf(dl[((size_t)n * K + k) * HW + hw], wh[(size_t)k * C + c], g);
        ga[i] += g;
    }
}
// dwh[k][c] = sum dl*a ; dbh[k] = sum dl : one block per (k, c) with c == C meaning the bias, fp64
__global__ void __launch_bounds__(256) check_head_bwd_params_kernel(const float* __restrict__ dl, const float* __restrict__ a, int N, int HW, int C,
                                                                    int K, float* __restrict__ dwh, float* __restrict__ dbh) {
    const int c = blockIdx.x % (C + 1), k = blockIdx.x / (C + 1);
    double s = 0.0;
    const long long M = (long long)N * HW;
    for (long long m = threadIdx.x; m < M; m += blockDim.x) {
        const long long n = m / HW, hw = m % HW;
        const double d = dl[((size_t)n * K + k) * HW + hw];
        s += (c < C) ? d * (double)a[(size_t)m * C + c] : d;
    }
    __shared__ double rs[256];
    rs[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) rs[threadIdx.x] += rs[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (c < C) dwh[(size_t)k * C + c] = (float)rs[0];
        else if (dbh) dbh[k] = (float)rs[0];
    }
}

}  // namespace nunet

using namespace nunet;

static inline int crc() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
static inline int nblocks(long long total) {
    long long b = (total + 255) / 256;
    return (int)(b > 148 * 32 ? 148 * 32 : (b < 1 ? 1 : b));
}

extern "C" int nunet_check_conv3x3(const float* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t xpitch, const float* w_oihw,
                                   const float* bias, int32_t Cout, float* wt_scratch, float* y, void* stream) {
    if (!x || !w_oihw || !wt_scratch || !y || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || xpitch < Cin) return NUNET_E_BADARG;
    check_repack_w_kernel<<<nblocks((long long)Cout * Cin * 9), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, wt_scratch);
    check_conv3x3_kernel<<<nblocks((long long)N * H * W * Cout), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, Cin, xpitch, wt_scratch, bias, Cout, y);
    return crc();
}

extern "C" int nunet_check_bn(const float* y, long long M, int32_t C, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, float momentum, float eps, int32_t train, float* scale, float* shift, float* save_mean,
                              float* save_invstd, void* stream) {
    if (!y || !scale || !shift || M < 1 || C < 1 || (!train && (!running_mean || !running_var))) return NUNET_E_BADARG;
    if (train)
        check_bn_stats_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(y, M, C, gamma, beta, running_mean, running_var, momentum, eps, scale, shift, save_mean,
                                                                  save_invstd);
    else
        check_bn_eval_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(C, gamma, beta, running_mean, running_var, eps, scale, shift);
    return crc();
}

extern "C" int nunet_check_bn_relu(const float* y, long long M, int32_t C, const float* scale, const float* shift, float* dst, int32_t pitch,
                                   int32_t off, void* stream) {
    if (!y || !scale || !shift || !dst || M < 1 || C < 1 || pitch < off + C) return NUNET_E_BADARG;
    check_bn_relu_kernel<<<nblocks(M * C), 256, 0, (cudaStream_t)stream>>>(y, M, C, scale, shift, dst, pitch, off);
    return crc();
}

extern "C" int nunet_check_route(const float* src, int32_t N, int32_t h, int32_t w, int32_t C, int32_t mode, float* dst, int32_t pitch,
                                 int32_t off, void* stream) {
    if (!src || !dst || N < 1 || h < 1 || w < 1 || C < 1 || mode < 0 || mode > 2 || pitch < off + C || (mode == 1 && ((h | w) & 1))) return NUNET_E_BADARG;
    const long long out_pix = mode == 1 ? (long long)N * (h / 2) * (w / 2) : (mode == 2 ? (long long)N * 4 * h * w : (long long)N * h * w);
    check_route_kernel<<<nblocks(out_pix * C), 256, 0, (cudaStream_t)stream>>>(src, N, h, w, C, mode, dst, pitch, off);
    return crc();
}

extern "C" int nunet_check_nchw_to_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream) {
    if (!x || !out || N < 1 || C < 1 || H < 1 || W < 1) return NUNET_E_BADARG;
    check_nchw_to_nhwc_kernel<<<nblocks((long long)N * C * H * W), 256, 0, (cudaStream_t)stream>>>(x, N, C, H * W, out);
    return crc();
}

extern "C" int nunet_check_head(const float* a, int32_t N, int32_t H, int32_t W, int32_t C, const float* wh, const float* bh, int32_t K,
                                float* logits, void* stream) {
    if (!a || !wh || !logits || N < 1 || C < 1 || K < 1) return NUNET_E_BADARG;
    check_head_kernel<<<nblocks((long long)N * K * H * W), 256, 0, (cudaStream_t)stream>>>(a, N, H * W, C, wh, bh, K, logits);
    return crc();
}

extern "C" int nunet_check_relu_bn_bwd(const float* ga, const float* y, long long M, int32_t C, const float* scale, const float* shift,
                                       const float* gamma, const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta,
                                       float* gy, void* stream) {
    if (!ga || !y || !scale || !shift || !save_mean || !save_invstd || !dgamma || !dbeta || !gy || M < 1 || C < 1) return NUNET_E_BADARG;
    check_bn_bwd_stats_kernel<<<C, 256, 0, (cudaStream_t)stream>>>(ga, y, M, C, scale, shift, save_mean, save_invstd, dgamma, dbeta);
    check_bn_bwd_apply_kernel<<<nblocks(M * C), 256, 0, (cudaStream_t)stream>>>(ga, y, M, C, scale, shift, gamma, save_mean, save_invstd, dgamma,
                                                                               dbeta, gy);
    return crc();
}

extern "C" int nunet_check_conv3x3_dgrad(const float* gy, int32_t N, int32_t H, int32_t W, int32_t Cout, const float* w_oihw, int32_t Cin,
                                         float* wt_scratch, float* dx, void* stream) {
    if (!gy || !w_oihw || !wt_scratch || !dx || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1) return NUNET_E_BADARG;
    check_repack_wd_kernel<<<nblocks((long long)Cout * Cin * 9), 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, wt_scratch);
    check_dgrad_kernel<<<nblocks((long long)N * H * W * Cin), 256, 0, (cudaStream_t)stream>>>(gy, N, H, W, Cout, wt_scratch, Cin, dx);
    return crc();
}

extern "C" int nunet_check_conv3x3_wgrad(const float* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t xpitch, const float* gy,
                                         int32_t Cout, float* dw_oihw, float* dbias, void* stream) {
    if (!x || !gy || !dw_oihw || N < 1 || H < 1 || W < 1 || Cin < 1 || Cout < 1 || xpitch < Cin) return NUNET_E_BADARG;
    check_wgrad_kernel<<<dim3(Cin * 9, (Cout + 31) / 32), 256, 0, (cudaStream_t)stream>>>(x, N, H, W, Cin, xpitch, gy, Cout, dw_oihw);
    if (dbias) check_colsum_kernel<<<Cout, 256, 0, (cudaStream_t)stream>>>(gy, (long long)N * H * W, Cout, dbias);
    return crc();
}

extern "C" int nunet_check_route_bwd(const float* dx, int32_t N, int32_t h, int32_t w, int32_t C, int32_t mode, int32_t pitch, int32_t off,
                                     const float* a_src, float* ga, void* stream) {
    if (!dx || !ga || N < 1 || h < 1 || w < 1 || C < 1 || mode < 0 || mode > 2 || pitch < off + C || (mode == 1 && (!a_src || ((h | w) & 1))))
        return NUNET_E_BADARG;
    check_route_bwd_kernel<<<nblocks((long long)N * h * w * C), 256, 0, (cudaStream_t)stream>>>(dx, N, h, w, C, mode, pitch, off, a_src, ga);
    return crc();
}

extern "C" int nunet_check_head_bwd(const float* dlogits, const float* a, int32_t N, int32_t H, int32_t W, int32_t C, const float* wh, int32_t K,
                                    float* ga, float* dwh, float* dbh, void* stream) {
    if (!dlogits || !a || !wh || !ga || !dwh || N < 1 || C < 1 || K < 1) return NUNET_E_BADARG;
    check_head_bwd_act_kernel<<<nblocks((long long)N * H * W * C), 256, 0, (cudaStream_t)stream>>>(dlogits, N, H * W, C, wh, K, ga);
    check_head_bwd_params_kernel<<<K * (C + 1), 256, 0, (cudaStream_t)stream>>>(dlogits, a, N, H * W, C, K, dwh, dbh);
    return crc();
}
