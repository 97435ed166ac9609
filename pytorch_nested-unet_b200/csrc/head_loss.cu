// 1x1 output heads, BCE+Dice loss, IoU counts and their backward — the HBM-trivial tail of the network.
//   head_fwd      : logits[n,k,y,x] = sum_c relu(scale*y+shift)[n,y,x,c] * Wh[k,c] + bh[k]   (finished/archs1.py:106-111,135-142)
//   loss_iou_fwd  : one pass over (logits, target): BCE-with-logits sum, per-sample sum(p*t), sum(p), sum(t)
//                   (losses.py:107-117) and the integer intersection/union counts of metrics.py:13-16,
//                   warp-shuffle + block reduction into per-block partials, finished deterministically in fp64.
//   loss_bwd      : dlogits of 0.5*bce + (1 - mean_n dice_n), scaled by the upstream gradient (device scalar).
//   head_bwd_params: dWh, dbh.
#include "conv_common.cuh"

namespace nunet {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float sigmoid_f32(float x) { return 1.0f / (1.0f + expf(-x)); }

constexpr int kMaxClasses = 8;

// ---------------------------------------------------------------------------------------------- head forward
// y: bf16 NHWC [NP, 32]; logits: fp32 NCHW [N, K, HW]
__global__ void __launch_bounds__(256) head_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const float* __restrict__ wh,
                                                       const float* __restrict__ bh, int K, int NP, int HW,
                                                       float* __restrict__ logits) {
    __shared__ float s_w[kMaxClasses * 32], s_b[kMaxClasses], s_sc[32], s_sh[32];
    for (int i = threadIdx.x; i < K * 32; i += blockDim.x) s_w[i] = wh[i];
    if (threadIdx.x < K) s_b[threadIdx.x] = bh[threadIdx.x];
    if (threadIdx.x < 32) { s_sc[threadIdx.x] = scale[threadIdx.x]; s_sh[threadIdx.x] = shift[threadIdx.x]; }
    __syncthreads();
    for (int pix = blockIdx.x * blockDim.x + threadIdx.x; pix < NP; pix += gridDim.x * blockDim.x) {
        float a[32];
        const uint4* src = reinterpret_cast<const uint4*>(y + (size_t)pix * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 r = __ldg(src + j);
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = __bfloat1622float2(h[i]);
                a[8 * j + 2 * i] = fmaxf(fmaf(t.x, s_sc[8 * j + 2 * i], s_sh[8 * j + 2 * i]), 0.f);
                a[8 * j + 2 * i + 1] = fmaxf(fmaf(t.y, s_sc[8 * j + 2 * i + 1], s_sh[8 * j + 2 * i + 1]), 0.f);
            }
        }
        const int n = pix / HW, rem = pix - n * HW;
        for (int k = 0; k < K; ++k) {
            float acc = s_b[k];
#pragma unroll
            for (int c = 0; c < 32; ++c) acc = fmaf(a[c], s_w[k * 32 + c], acc);
            logits[((size_t)n * K + k) * HW + rem] = acc;
        }
    }
}

// ---------------------------------------------------------------------------------------------- loss + IoU forward
// grid (BPS, N): block b of sample n strides over the sample's E = K*H*W elements.
// part: [N][BPS][4] float (bce, p*t, p, t); cnt: [N][BPS][2] int (I, U)
__global__ void __launch_bounds__(256) loss_iou_partial_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                               int E, float* __restrict__ part, int* __restrict__ cnt) {
    const int n = blockIdx.y;
    const float* x = logits + (size_t)n * E;
    const float* t = target ? target + (size_t)n * E : nullptr;
    float bce = 0.f, pt = 0.f, ps = 0.f, ts = 0.f;
    int I = 0, U = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
        const float xv = x[i], tv = t[i];
        const float p = sigmoid_f32(xv);
        bce += fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)));
        pt = fmaf(p, tv, pt);
        ps += p;
        ts += tv;
        const bool o_ = p > 0.5f, t_ = tv > 0.5f;
        I += (o_ && t_);
        U += (o_ || t_);
    }
    __shared__ float sf[8][4];
    __shared__ int si[8][2];
    bce = warp_sum(bce); pt = warp_sum(pt); ps = warp_sum(ps); ts = warp_sum(ts);
    I = warp_sum_i(I); U = warp_sum_i(U);
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sf[w][0] = bce; sf[w][1] = pt; sf[w][2] = ps; sf[w][3] = ts; si[w][0] = I; si[w][1] = U; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float a = 0.f;
        for (int i = 0; i < 8; ++i) a += sf[i][threadIdx.x];
        part[((size_t)n * gridDim.x + blockIdx.x) * 4 + threadIdx.x] = a;
    } else if (threadIdx.x < 6) {
        int a = 0;
        for (int i = 0; i < 8; ++i) a += si[i][threadIdx.x - 4];
        cnt[((size_t)n * gridDim.x + blockIdx.x) * 2 + threadIdx.x - 4] = a;
    }
}

// single block. sums: [N][3] double (I=sum p*t, P, T) kept for backward; out: loss (float), counts: [2] int64
__global__ void __launch_bounds__(256) loss_iou_finalize_kernel(const float* __restrict__ part, const int* __restrict__ cnt, int N,
                                                                int BPS, int E, double* __restrict__ sums, float* __restrict__ loss,
                                                                long long* __restrict__ counts) {
    __shared__ double s_bce[256], s_dice[256];
    __shared__ long long s_I[256], s_U[256];
    double bce = 0.0, dice = 0.0;
    long long I = 0, U = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double b = 0, pt = 0, ps = 0, ts = 0;
        for (int k = 0; k < BPS; ++k) {
            const float* q = part + ((size_t)n * BPS + k) * 4;
            b += q[0]; pt += q[1]; ps += q[2]; ts += q[3];
            I += cnt[((size_t)n * BPS + k) * 2];
            U += cnt[((size_t)n * BPS + k) * 2 + 1];
        }
        bce += b;
        const double smooth = 1e-5;
        dice += (2.0 * pt + smooth) / (ps + ts + smooth);
        sums[n * 3 + 0] = pt; sums[n * 3 + 1] = ps; sums[n * 3 + 2] = ts;
    }
    s_bce[threadIdx.x] = bce; s_dice[threadIdx.x] = dice; s_I[threadIdx.x] = I; s_U[threadIdx.x] = U;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if (threadIdx.x < o) {
            s_bce[threadIdx.x] += s_bce[threadIdx.x + o]; s_dice[threadIdx.x] += s_dice[threadIdx.x + o];
            s_I[threadIdx.x] += s_I[threadIdx.x + o]; s_U[threadIdx.x] += s_U[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (loss) *loss = (float)(0.5 * s_bce[0] / ((double)N * E) + (1.0 - s_dice[0] / N));
        if (counts) { counts[0] = s_I[0]; counts[1] = s_U[0]; }
    }
}

// IoU counts only (metrics.iou_score on device logits; target may be any float tensor)
__global__ void __launch_bounds__(256) iou_count_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                        long long total, unsigned long long* __restrict__ counts) {
    int I = 0, U = 0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const bool o_ = sigmoid_f32(logits[i]) > 0.5f, t_ = target[i] > 0.5f;
        I += (o_ && t_);
        U += (o_ || t_);
    }
    I = warp_sum_i(I); U = warp_sum_i(U);
    if ((threadIdx.x & 31) == 0) {  // integer atomics: order-independent, exact
        atomicAdd(&counts[0], (unsigned long long)I);
        atomicAdd(&counts[1], (unsigned long long)U);
    }
}

// ---------------------------------------------------------------------------------------------- loss backward
__global__ void __launch_bounds__(256) loss_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                       const double* __restrict__ sums, const float* __restrict__ gscale, float gmul,
                                                       int N, int E, float* __restrict__ dlogits) {
    const int n = blockIdx.y;
    const double smooth = 1e-5;
    const double pt = sums[n * 3], ps = sums[n * 3 + 1], ts = sums[n * 3 + 2];
    const double den = ps + ts + smooth, num = 2.0 * pt + smooth;
    const float g = (gscale ? *gscale : 1.f) * gmul;
    const float c_bce = 0.5f / ((float)N * (float)E);
    const float a_t = (float)(-2.0 / (den * N));       // coefficient of t_i in d(1 - mean dice)/dp_i
    const float a_0 = (float)(num / (den * den * N));  // constant term
    const float* x = logits + (size_t)n * E;
    const float* t = target + (size_t)n * E;
    float* d = dlogits + (size_t)n * E;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
        const float p = sigmoid_f32(x[i]), tv = t[i];
        d[i] = g * (c_bce * (p - tv) + (a_t * tv + a_0) * p * (1.f - p));
    }
}

// ---------------------------------------------------------------------------------------------- all deep-supervision heads at once
// trains.py:120-123 calls the criterion once per head ON THE SAME TARGET: here the NH heads share one pass (the target is read once,
// one launch instead of NH) with exactly the per-head thread -> element mapping and reduction order of the kernels above, so every
// head's partial sums, loss, counts and gradient are bitwise those of the single-head launches.
constexpr int kMaxHeads = 4;
struct LossHeads {
    const float* logits[kMaxHeads];
    const float* gscale[kMaxHeads];   // backward: upstream gradient of head h's loss (device scalar) or nullptr (= 1)
    float* dlogits[kMaxHeads];        // backward: nullptr = this head's loss received no gradient
};

// part: [NH][N][BPS][4], cnt: [NH][N][BPS][2]
template <int NH>
__global__ void __launch_bounds__(256) loss_iou_partial_multi_kernel(const LossHeads h, const float* __restrict__ target, int E,
                                                                     float* __restrict__ part, int* __restrict__ cnt) {
    const int n = blockIdx.y, N = gridDim.y;
    const float* t = target + (size_t)n * E;
    float bce[NH], pt[NH], ps[NH], ts[NH];
    int I[NH], U[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) { bce[k] = 0.f; pt[k] = 0.f; ps[k] = 0.f; ts[k] = 0.f; I[k] = 0; U[k] = 0; }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
        const float tv = t[i];
        const bool t_ = tv > 0.5f;
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            const float xv = h.logits[k][(size_t)n * E + i];
            const float p = sigmoid_f32(xv);
            bce[k] += fmaxf(xv, 0.f) - xv * tv + log1pf(expf(-fabsf(xv)));
            pt[k] = fmaf(p, tv, pt[k]);
            ps[k] += p;
            ts[k] += tv;
            const bool o_ = p > 0.5f;
            I[k] += (o_ && t_);
            U[k] += (o_ || t_);
        }
    }
    __shared__ float sf[NH][8][4];
    __shared__ int si[NH][8][2];
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const float a = warp_sum(bce[k]), b = warp_sum(pt[k]), c = warp_sum(ps[k]), d = warp_sum(ts[k]);
        const int ii = warp_sum_i(I[k]), uu = warp_sum_i(U[k]);
        if ((threadIdx.x & 31) == 0) { sf[k][w][0] = a; sf[k][w][1] = b; sf[k][w][2] = c; sf[k][w][3] = d; si[k][w][0] = ii; si[k][w][1] = uu; }
    }
    __syncthreads();
    if (threadIdx.x < 6 * NH) {
        const int k = threadIdx.x / 6, j = threadIdx.x - 6 * k;
        const size_t blk = ((size_t)k * N + n) * gridDim.x + blockIdx.x;
        if (j < 4) {
            float a = 0.f;
            for (int i = 0; i < 8; ++i) a += sf[k][i][j];
            part[blk * 4 + j] = a;
        } else {
            int a = 0;
            for (int i = 0; i < 8; ++i) a += si[k][i][j - 4];
            cnt[blk * 2 + j - 4] = a;
        }
    }
}

// block k finishes head k: sums [NH][N][3], loss [NH], counts [NH][2]
__global__ void __launch_bounds__(256) loss_iou_finalize_multi_kernel(const float* __restrict__ part, const int* __restrict__ cnt, int N,
                                                                      int BPS, int E, double* __restrict__ sums, float* __restrict__ loss,
                                                                      long long* __restrict__ counts) {
    const int k = blockIdx.x;
    part += (size_t)k * N * BPS * 4;
    cnt += (size_t)k * N * BPS * 2;
    sums += (size_t)k * N * 3;
    __shared__ double s_bce[256], s_dice[256];
    __shared__ long long s_I[256], s_U[256];
    double bce = 0.0, dice = 0.0;
    long long I = 0, U = 0;
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        double b = 0, pt = 0, ps = 0, ts = 0;
        for (int q_ = 0; q_ < BPS; ++q_) {
            const float* q = part + ((size_t)n * BPS + q_) * 4;
            b += q[0]; pt += q[1]; ps += q[2]; ts += q[3];
            I += cnt[((size_t)n * BPS + q_) * 2];
            U += cnt[((size_t)n * BPS + q_) * 2 + 1];
        }
        bce += b;
        const double smooth = 1e-5;
        dice += (2.0 * pt + smooth) / (ps + ts + smooth);
        sums[n * 3 + 0] = pt; sums[n * 3 + 1] = ps; sums[n * 3 + 2] = ts;
    }
    s_bce[threadIdx.x] = bce; s_dice[threadIdx.x] = dice; s_I[threadIdx.x] = I; s_U[threadIdx.x] = U;
    __syncthreads();
    for (int o = 128; o >= 1; o >>= 1) {
        if (threadIdx.x < o) {
            s_bce[threadIdx.x] += s_bce[threadIdx.x + o]; s_dice[threadIdx.x] += s_dice[threadIdx.x + o];
            s_I[threadIdx.x] += s_I[threadIdx.x + o]; s_U[threadIdx.x] += s_U[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        loss[k] = (float)(0.5 * s_bce[0] / ((double)N * E) + (1.0 - s_dice[0] / N));
        counts[2 * k] = s_I[0];
        counts[2 * k + 1] = s_U[0];
    }
}

template <int NH>
__global__ void __launch_bounds__(256) loss_bwd_multi_kernel(const LossHeads h, const float* __restrict__ target,
                                                             const double* __restrict__ sums, float gmul, int N, int E) {
    const int n = blockIdx.y;
    const double smooth = 1e-5;
    const float c_bce = 0.5f / ((float)N * (float)E);
    float g[NH], a_t[NH], a_0[NH];
#pragma unroll
    for (int k = 0; k < NH; ++k) {
        const double* sm = sums + ((size_t)k * N + n) * 3;
        const double pt = sm[0], ps = sm[1], ts = sm[2];
        const double den = ps + ts + smooth, num = 2.0 * pt + smooth;
        g[k] = (h.gscale[k] ? *h.gscale[k] : 1.f) * gmul;
        a_t[k] = (float)(-2.0 / (den * N));
        a_0[k] = (float)(num / (den * den * N));
    }
    const float* t = target + (size_t)n * E;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < E; i += gridDim.x * blockDim.x) {
        const float tv = t[i];
#pragma unroll
        for (int k = 0; k < NH; ++k) {
            if (h.dlogits[k] == nullptr) continue;
            const float p = sigmoid_f32(h.logits[k][(size_t)n * E + i]);
            h.dlogits[k][(size_t)n * E + i] = g[k] * (c_bce * (p - tv) + (a_t[k] * tv + a_0[k]) * p * (1.f - p));
        }
    }
}

// ---------------------------------------------------------------------------------------------- head parameter grads
// dWh[k][c] = sum_pix dlogit[n,k,pix] * a[pix,c] ; dbh[k] = sum dlogit.  part: [gridDim.x][K][33]
template <int KT>
__global__ void __launch_bounds__(256) head_bwd_params_partial_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, const float* __restrict__ dlogits,
                                                                      int K, int NP, int HW, float* __restrict__ part) {
    // thread = (pixel lane pl = tid/4, channel group cg = tid%4); each thread owns 8 channels
    const int cg = threadIdx.x & 3;
    float sc[8], sh[8], acc[KT][8], accb[KT];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i] = scale[cg * 8 + i]; sh[i] = shift[cg * 8 + i]; }
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        accb[k] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[k][i] = 0.f;
    }
    constexpr int U = 4;   // pixels in flight per thread (the loop is latency-bound otherwise: one 16-byte load per iteration)
    const int pstride = gridDim.x * 64;
    for (int pix0 = blockIdx.x * 64 + (threadIdx.x >> 2); pix0 < NP; pix0 += U * pstride) {
        uint4 r[U];
        float dl[U][KT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pix = pix0 + u * pstride;
            if (pix < NP) {
                r[u] = __ldg(reinterpret_cast<const uint4*>(y + (size_t)pix * 32 + cg * 8));
                const int n = pix / HW, rem = pix - n * HW;
#pragma unroll
                for (int k = 0; k < KT; ++k)
                    if (k < K) dl[u][k] = __ldg(dlogits + ((size_t)n * K + k) * HW + rem);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (pix0 + u * pstride >= NP) continue;
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r[u]);
            float a[8];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = __bfloat1622float2(h[i]);
                a[2 * i] = fmaxf(fmaf(t.x, sc[2 * i], sh[2 * i]), 0.f);
                a[2 * i + 1] = fmaxf(fmaf(t.y, sc[2 * i + 1], sh[2 * i + 1]), 0.f);
            }
#pragma unroll
            for (int k = 0; k < KT; ++k) {
                if (k < K) {
                    accb[k] += dl[u][k];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[k][i] = fmaf(dl[u][k], a[i], acc[k][i]);
                }
            }
        }
    }
    __shared__ float red[8][KT][33];
    for (int i = threadIdx.x; i < 8 * KT * 33; i += blockDim.x) (&red[0][0][0])[i] = 0.f;
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KT; ++k) {
        if (k < K) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float v = acc[k][i];  // reduce over the 8 pixel lanes of this warp (lanes with equal cg: stride 4)
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                v += __shfl_xor_sync(0xffffffffu, v, 16);
                if (lane < 4) red[w][k][cg * 8 + i] = v;
            }
            float b = (cg == 0) ? accb[k] : 0.f;
            b = warp_sum(b);
            if (lane == 0) red[w][k][32] = b;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * 33; i += blockDim.x) {
        const int k = i / 33, c = i - k * 33;
        float s = 0.f;
        for (int ww = 0; ww < 8; ++ww) s += red[ww][k][c];
        part[((size_t)blockIdx.x * K + k) * 33 + c] = s;
    }
}
// grid K, block 64: sums the block partials in fp64 (fixed order)
__global__ void head_bwd_params_finalize_kernel(const float* __restrict__ part, int rows, int K, float* __restrict__ dwh,
                                                float* __restrict__ dbh, int accumulate) {
    const int k = blockIdx.x, c = threadIdx.x;
    if (c >= 33) return;
    double s = 0.0;
    for (int r = 0; r < rows; ++r) s += part[((size_t)r * K + k) * 33 + c];
    if (c < 32) dwh[k * 32 + c] = (accumulate ? dwh[k * 32 + c] : 0.f) + (float)s;
    else dbh[k] = (accumulate ? dbh[k] : 0.f) + (float)s;
}

// val.py:100-105: torch.sigmoid(output).cpu().numpy(), then (p * 255).astype('uint8') per image and class — the sigmoid, the scaling and
// the truncating cast run on the device, only the uint8 planes cross to the host (4x fewer bytes, no second fp32 pass on the CPU)
__global__ void __launch_bounds__(256) sigmoid_u8_kernel(const float* __restrict__ logits, long long total, uint8_t* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const float v = sigmoid_f32(logits[i]) * 255.0f;      // fp32 product like numpy's float32 array * 255
        out[i] = (uint8_t)(int)v;                             // astype('uint8'): truncation toward zero (values are in [0, 255])
    }
}

}  // namespace nunet

using namespace nunet;
static inline int cuda_rc() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
static inline int clampi(long long v, int lo, int hi) { return (int)(v < lo ? lo : (v > hi ? hi : v)); }

extern "C" int nunet_head_fwd(const void* y, const float* scale, const float* shift, const float* wh, const float* bh, int32_t N,
                              int32_t H, int32_t W, int32_t K, float* logits, void* stream) {
    if (!y || !scale || !shift || !wh || !bh || !logits || K < 1 || K > kMaxClasses || N < 1) return NUNET_E_BADARG;
    const long long NP = (long long)N * H * W;
    if (NP >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    const int blocks = clampi((NP + 255) / 256, 1, 148 * 16);
    head_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, scale, shift, wh, bh, K, (int)NP, H * W, logits);
    return cuda_rc();
}

extern "C" int nunet_loss_blocks_per_sample(int32_t N, int32_t E) {
    long long want = (148LL * 8 + N - 1) / N;  // ~8 blocks per SM over the batch
    long long maxb = ((long long)E + 1023) / 1024;
    return clampi(want < maxb ? want : maxb, 1, 4096);
}

extern "C" int nunet_loss_iou_fwd(const float* logits, const float* target, int32_t N, int32_t E, float* part, int32_t* cnt,
                                  double* sums, float* loss, long long* counts, void* stream) {
    if (!logits || !target || !part || !cnt || !sums || N < 1 || E < 1) return NUNET_E_BADARG;
    const int bps = nunet_loss_blocks_per_sample(N, E);
    loss_iou_partial_kernel<<<dim3(bps, N), 256, 0, (cudaStream_t)stream>>>(logits, target, E, part, cnt);
    loss_iou_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(part, cnt, N, bps, E, sums, loss, counts);
    return cuda_rc();
}

extern "C" int nunet_iou_counts(const float* logits, const float* target, long long total, unsigned long long* counts, void* stream) {
    if (!logits || !target || !counts || total < 1) return NUNET_E_BADARG;
    cudaError_t e = cudaMemsetAsync(counts, 0, 16, (cudaStream_t)stream);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    const int blocks = clampi((total + 1023) / 1024, 1, 148 * 8);
    iou_count_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(logits, target, total, counts);
    return cuda_rc();
}

extern "C" int nunet_loss_bwd(const float* logits, const float* target, const double* sums, const float* gscale, float gmul, int32_t N,
                              int32_t E, float* dlogits, void* stream) {
    if (!logits || !target || !sums || !dlogits || N < 1 || E < 1) return NUNET_E_BADARG;
    const int bps = nunet_loss_blocks_per_sample(N, E);
    loss_bwd_kernel<<<dim3(bps, N), 256, 0, (cudaStream_t)stream>>>(logits, target, sums, gscale, gmul, N, E, dlogits);
    return cuda_rc();
}

extern "C" int nunet_loss_iou_fwd_multi(const float* const* logits, int32_t nheads, const float* target, int32_t N, int32_t E, float* part,
                                        int32_t* cnt, double* sums, float* loss, long long* counts, void* stream) {
    if (!logits || nheads < 1 || nheads > kMaxHeads || !target || !part || !cnt || !sums || !loss || !counts || N < 1 || E < 1) return NUNET_E_BADARG;
    LossHeads h{};
    for (int k = 0; k < nheads; ++k) {
        if (!logits[k]) return NUNET_E_BADARG;
        h.logits[k] = logits[k];
    }
    const int bps = nunet_loss_blocks_per_sample(N, E);
    const dim3 grid(bps, N);
    cudaStream_t st = (cudaStream_t)stream;
    switch (nheads) {
        case 1: loss_iou_partial_multi_kernel<1><<<grid, 256, 0, st>>>(h, target, E, part, cnt); break;
        case 2: loss_iou_partial_multi_kernel<2><<<grid, 256, 0, st>>>(h, target, E, part, cnt); break;
        case 3: loss_iou_partial_multi_kernel<3><<<grid, 256, 0, st>>>(h, target, E, part, cnt); break;
        default: loss_iou_partial_multi_kernel<4><<<grid, 256, 0, st>>>(h, target, E, part, cnt); break;
    }
    loss_iou_finalize_multi_kernel<<<nheads, 256, 0, st>>>(part, cnt, N, bps, E, sums, loss, counts);
    return cuda_rc();
}

extern "C" int nunet_loss_bwd_multi(const float* const* logits, int32_t nheads, const float* target, const double* sums,
                                    const float* const* gscale, float gmul, int32_t N, int32_t E, float* const* dlogits, void* stream) {
    if (!logits || nheads < 1 || nheads > kMaxHeads || !target || !sums || !dlogits || N < 1 || E < 1) return NUNET_E_BADARG;
    LossHeads h{};
    for (int k = 0; k < nheads; ++k) {
        if (!logits[k]) return NUNET_E_BADARG;
        h.logits[k] = logits[k];
        h.gscale[k] = gscale ? gscale[k] : nullptr;
        h.dlogits[k] = dlogits[k];
    }
    const int bps = nunet_loss_blocks_per_sample(N, E);
    const dim3 grid(bps, N);
    cudaStream_t st = (cudaStream_t)stream;
    switch (nheads) {
        case 1: loss_bwd_multi_kernel<1><<<grid, 256, 0, st>>>(h, target, sums, gmul, N, E); break;
        case 2: loss_bwd_multi_kernel<2><<<grid, 256, 0, st>>>(h, target, sums, gmul, N, E); break;
        case 3: loss_bwd_multi_kernel<3><<<grid, 256, 0, st>>>(h, target, sums, gmul, N, E); break;
        default: loss_bwd_multi_kernel<4><<<grid, 256, 0, st>>>(h, target, sums, gmul, N, E); break;
    }
    return cuda_rc();
}

extern "C" int nunet_head_bwd_params_rows(void) { return 148 * 4; }

extern "C" int nunet_head_bwd_params(const void* y, const float* scale, const float* shift, const float* dlogits, int32_t N, int32_t H,
                                     int32_t W, int32_t K, float* part, float* dwh, float* dbh, int32_t accumulate, void* stream) {
    if (!y || !scale || !shift || !dlogits || !part || !dwh || !dbh || K < 1 || K > kMaxClasses) return NUNET_E_BADARG;
    const long long NP = (long long)N * H * W;
    if (NP >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    const int rows = clampi((NP + 63) / 64, 1, nunet_head_bwd_params_rows());
    if (K == 1)
        head_bwd_params_partial_kernel<1><<<rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, scale, shift, dlogits, K, (int)NP, H * W, part);
    else if (K <= 4)
        head_bwd_params_partial_kernel<4><<<rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, scale, shift, dlogits, K, (int)NP, H * W, part);
    else
        head_bwd_params_partial_kernel<8><<<rows, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)y, scale, shift, dlogits, K, (int)NP, H * W, part);
    head_bwd_params_finalize_kernel<<<K, 64, 0, (cudaStream_t)stream>>>(part, rows, K, dwh, dbh, accumulate);
    return cuda_rc();
}

extern "C" int nunet_sigmoid_u8(const float* logits, long long total, uint8_t* out, void* stream) {
    if (!logits || !out || total < 1) return NUNET_E_BADARG;
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    sigmoid_u8_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, total, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
