// Transposed align_corners bilinear 2x upsample (backward of nn.Upsample(scale_factor=2, mode='bilinear',
// align_corners=True), finished/archs1.py:42,83): routes a consumer's dgrad slice at (2h, 2w) back to the (h, w)
// resolution of the node that was upsampled.  out[n,y,x,c] = sum_{Y,X} wy(Y,y) * wx(X,x) * dx[n,Y,X,coff+c].
//
// One CTA = an 8x8 tile of source pixels x 32 channels.  The (<= 22 x 22) window of contributing output pixels is
// brought into shared memory once with cp.async; the gather is separable: x-reduction into an fp32 tile, then
// y-reduction, so every dX element is read from HBM/L2 ~(22/16)^2 times instead of ~4 gathers per element from 16
// scattered loads per thread.  The result is a dense bf16 NHWC tensor the gradient router reads as a SAME consumer.
#include "conv_common.cuh"
#include <cstdlib>

namespace nunet {

constexpr int kUT = 8;             // source tile edge
constexpr int kUW = 2 * kUT + 6;   // output window edge (candidates [2s-3, 2s+4])
constexpr int kUC = 32;            // channels per CTA (4 CTAs per SM)
constexpr int kTaps = 5;           // contributing outputs of one source index are consecutive and at most 5 (2/r = 4.0x, r < 1/2)

__device__ __forceinline__ float upw(int D, int s, int h, float r) {
    const float src = r * (float)D;
    const int i0 = (int)src;
    const int i1 = i0 + (i0 < h - 1);
    const float l1 = src - (float)i0;
    return (i0 == s ? 1.f - l1 : 0.f) + (i1 == s ? l1 : 0.f);
}

struct UpTParams {
    const __nv_bfloat16* dx;
    int pitch, coff, N, h, w, C;
    __nv_bfloat16* out;
};

__global__ void __launch_bounds__(256, 4) up_transpose_kernel(const UpTParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kCg = kUC / 8;
    uint4* s_in = reinterpret_cast<uint4*>(smem);                                    // [kUW][kUW][kCg] x 16 B
    float* s_tmp = reinterpret_cast<float*>(smem + (size_t)kUW * kUW * kCg * 16);     // [kUW][kUT][kUC] fp32
    float* s_wy = s_tmp + kUW * kUT * kUC;                                            // [kUT][kTaps]
    float* s_wx = s_wy + kUT * kTaps;                                                 // [kUT][kTaps]
    int* s_k0 = reinterpret_cast<int*>(s_wx + kUT * kTaps);                           // [2][kUT] first contributing candidate (window coords 2s + k0)
    const int tiles_x = (p.w + kUT - 1) / kUT, tiles_y = (p.h + kUT - 1) / kUT;
    int t = blockIdx.x;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    const int n = t;
    const int c0 = blockIdx.y * kUC;
    const int ncg = min(kCg, (p.C - c0) >> 3);
    const int y0 = ty * kUT, x0 = tx * kUT;
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    const int Ymin = 2 * y0 - 3, Xmin = 2 * x0 - 3;
    const float ry = (float)(p.h - 1) / (float)(H2 - 1), rx = (float)(p.w - 1) / (float)(W2 - 1);
    // per tile row / column s: first candidate k0 in [0,8) with a non-zero weight and the kTaps weights from there
    if (threadIdx.x < 2 * kUT) {
        const int which = threadIdx.x / kUT, s = threadIdx.x % kUT;
        const int src = (which == 0 ? y0 : x0) + s, hh = which == 0 ? p.h : p.w, D2 = which == 0 ? H2 : W2;
        const float r = which == 0 ? ry : rx;
        float wk[8];
        int k0 = 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int D = 2 * src - 3 + k;
            wk[k] = (src < hh && D >= 0 && D < D2) ? upw(D, src, hh, r) : 0.f;
            if (wk[k] != 0.f && k0 == 8) k0 = k;
        }
        if (k0 > 8 - kTaps) k0 = 8 - kTaps;
        float* dstw = (which == 0 ? s_wy : s_wx) + s * kTaps;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k == k0 + j) v = wk[k];
            dstw[j] = v;
        }
        s_k0[which * kUT + s] = k0;
    }
    // window load: kUW x kUW output pixels x ncg 16-byte groups (zero outside the image)
    for (int i = threadIdx.x; i < kUW * kUW * kCg; i += 256) {
        const int cg = i % kCg, px = (i / kCg) % kUW, py = i / (kCg * kUW);
        const int Y = Ymin + py, X = Xmin + px;
        const bool ok = cg < ncg && Y >= 0 && Y < H2 && X >= 0 && X < W2;
        const size_t pix = ((size_t)n * H2 + (ok ? Y : 0)) * W2 + (ok ? X : 0);
        cp_async16(&s_in[i], p.dx + pix * p.pitch + p.coff + c0 + (ok ? cg : 0) * 8, ok ? 16u : 0u);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    // x-reduction: tmp[py][xs][c] = sum_j wx[xs][j] * in[py][2*xs + k0 + j (window coords)][c]
    for (int i = threadIdx.x; i < kUW * kUT * kCg; i += 256) {
        const int cg = i % kCg, xs = (i / kCg) % kUT, py = i / (kCg * kUT);
        const int kb = 2 * xs + s_k0[kUT + xs];
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const float wgt = s_wx[xs * kTaps + j];
            const uint4 r = s_in[(py * kUW + kb + j) * kCg + cg];
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hh[e]);
                acc[2 * e] = fmaf(wgt, f.x, acc[2 * e]);
                acc[2 * e + 1] = fmaf(wgt, f.y, acc[2 * e + 1]);
            }
        }
        float4* d = reinterpret_cast<float4*>(s_tmp + ((size_t)py * kUT + xs) * kUC + cg * 8);
        d[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        d[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    __syncthreads();
    // y-reduction and store
    for (int i = threadIdx.x; i < kUT * kUT * kCg; i += 256) {
        const int cg = i % kCg, xs = (i / kCg) % kUT, ys = i / (kCg * kUT);
        const int y = y0 + ys, x = x0 + xs;
        if (y >= p.h || x >= p.w || cg >= ncg) continue;
        const int kb = 2 * ys + s_k0[ys];
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const float wgt = s_wy[ys * kTaps + j];
            const float4* s4 = reinterpret_cast<const float4*>(s_tmp + ((size_t)(kb + j) * kUT + xs) * kUC + cg * 8);
            const float4 a = s4[0], b = s4[1];
            acc[0] = fmaf(wgt, a.x, acc[0]); acc[1] = fmaf(wgt, a.y, acc[1]); acc[2] = fmaf(wgt, a.z, acc[2]); acc[3] = fmaf(wgt, a.w, acc[3]);
            acc[4] = fmaf(wgt, b.x, acc[4]); acc[5] = fmaf(wgt, b.y, acc[5]); acc[6] = fmaf(wgt, b.z, acc[6]); acc[7] = fmaf(wgt, b.w, acc[7]);
        }
        uint32_t o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 q = __floats2bfloat162_rn(acc[2 * e], acc[2 * e + 1]);
            o[e] = *reinterpret_cast<uint32_t*>(&q);
        }
        *reinterpret_cast<uint4*>(p.out + (((size_t)n * p.h + y) * p.w + x) * p.C + c0 + cg * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// ---- streaming variant (default).  A thread owns one (source column x, 8-channel group) and walks the output rows Y that
// touch its block of source rows.  Per output row it gathers the <= 5 contributing output columns (weights precomputed once),
// and adds the x-reduced value to the two source rows Y blends (y0 with 1-ly, y1 with ly): y0 is non-decreasing in Y, so two
// running row accumulators suffice and a row is written out the moment y0 moves past it.  10 16-byte loads per output item
// (L1 serves the overlap between neighbouring threads), no shared memory, no barriers: HBM-bound instead of issue-bound.
constexpr int kRB = 16;   // source rows per CTA (3 extra output rows are re-read per CTA)
constexpr int kRBBulk = 24; // bulk-copy variant: issue-bound, so fewer re-read rows win (0.091 ms at 16, 0.086 ms at 24 for 48x48x64; profiles/r02_summary.md)

__global__ void __launch_bounds__(128) up_transpose_stream_kernel(const UpTParams p) {
    const int cgs = p.C >> 3;
    const int item = blockIdx.y * 128 + threadIdx.x;
    if (item >= p.w * cgs) return;
    const int cg = item % cgs, x = item / cgs;
    const int rblocks = (p.h + kRB - 1) / kRB;
    const int n = blockIdx.x / rblocks, ys = (blockIdx.x - n * rblocks) * kRB;
    const int ye = min(p.h, ys + kRB);
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    const float ry = (float)(p.h - 1) / (float)(H2 - 1), rx = (float)(p.w - 1) / (float)(W2 - 1);
    // the <= kTaps consecutive output columns that read source column x
    float wx[kTaps];
    int X0 = 0;
    {
        float wk[8];
        int k0 = 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int D = 2 * x - 3 + k;
            wk[k] = (D >= 0 && D < W2) ? upw(D, x, p.w, rx) : 0.f;
            if (wk[k] != 0.f && k0 == 8) k0 = k;
        }
        if (k0 > 8 - kTaps) k0 = 8 - kTaps;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k == k0 + j) v = wk[k];
            wx[j] = v;
        }
        X0 = 2 * x - 3 + k0;
        // clamp the window into the image (weights of clamped-away columns are zero by construction)
        if (X0 < 0) {
#pragma unroll
            for (int sft = 0; sft < 3; ++sft)
                if (X0 < 0) {
#pragma unroll
                    for (int j = 0; j < kTaps - 1; ++j) wx[j] = wx[j + 1];
                    wx[kTaps - 1] = 0.f;
                    ++X0;
                }
        }
    }
    const __nv_bfloat16* src = p.dx + (size_t)n * H2 * W2 * p.pitch + p.coff + cg * 8;
    __nv_bfloat16* dst = p.out + (size_t)n * p.h * p.w * p.C + (size_t)x * p.C + cg * 8;
    float acc_a[8], acc_b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc_a[e] = 0.f; acc_b[e] = 0.f; }
    const int Ystart = max(0, 2 * ys - 4);
    int cur = (int)(ry * (float)Ystart);
    for (int Y = Ystart; Y < H2; ++Y) {
        const float sy = ry * (float)Y;
        const int y0 = (int)sy;
        if (y0 >= ye) break;
        const int y1 = y0 + (y0 < p.h - 1);
        const float ly = sy - (float)y0;
        if (y0 > cur) {   // row `cur` is complete
            if (cur >= ys) {
                uint4 o;
                __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) q[e] = __floats2bfloat162_rn(acc_a[2 * e], acc_a[2 * e + 1]);
                *reinterpret_cast<uint4*>(dst + (size_t)cur * p.w * p.C) = o;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc_a[e] = acc_b[e]; acc_b[e] = 0.f; }
            cur = y0;
        }
        const __nv_bfloat16* rowp = src + ((size_t)Y * W2 + X0) * p.pitch;
        uint4 r[kTaps];
#pragma unroll
        for (int j = 0; j < kTaps; ++j) r[j] = (X0 + j < W2) ? __ldg(reinterpret_cast<const uint4*>(rowp + (size_t)j * p.pitch)) : make_uint4(0, 0, 0, 0);
        float xr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) xr[e] = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r[j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hh[e]);
                xr[2 * e] = fmaf(wx[j], f.x, xr[2 * e]);
                xr[2 * e + 1] = fmaf(wx[j], f.y, xr[2 * e + 1]);
            }
        }
        const float w0 = 1.f - ly;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc_a[e] = fmaf(w0, xr[e], acc_a[e]);
        if (y1 == y0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc_a[e] = fmaf(ly, xr[e], acc_a[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc_b[e] = fmaf(ly, xr[e], acc_b[e]);
        }
    }
    // rows still held: `cur` (acc_a) and `cur + 1` (acc_b)
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int row = cur + t;
        if (row >= ys && row < ye) {
            uint4 o;
            __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) q[e] = t == 0 ? __floats2bfloat162_rn(acc_a[2 * e], acc_a[2 * e + 1]) : __floats2bfloat162_rn(acc_b[2 * e], acc_b[2 * e + 1]);
            *reinterpret_cast<uint4*>(dst + (size_t)row * p.w * p.C) = o;
        }
    }
}

// ---- bulk-copy variant (default when a dX row is contiguous and w * C / 8 <= 992).  The streaming kernel above is latency-bound:
// five 16-byte loads in flight per thread, 88 % of its stall samples on the load scoreboard, 2.6 TB/s (profiles/r02_summary.md).
// Here one elected thread of a producer warp streams whole dX rows (W2 * C bf16, 12 KB at every level of the 96x96 step) into a
// ring of shared-memory stages with cp.async.bulk (up to 72 KB in flight per CTA, no registers spent on it); the consumer threads
// run the same per-thread walk as the streaming kernel, reading their five taps from the stage.
struct UpTBulkParams {
    const __nv_bfloat16* dx;
    int N, h, w, C, rb, S;
    __nv_bfloat16* out;
};

__global__ void __launch_bounds__(1024) up_transpose_bulk_kernel(const UpTBulkParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int cgs = p.C >> 3, items = p.w * cgs;
    const int cwarps = (items + 31) >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int H2 = 2 * p.h, W2 = 2 * p.w;
    const uint32_t row_bytes = (uint32_t)W2 * (uint32_t)p.C * 2u;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)p.S * row_bytes);
    uint64_t* empty = full + p.S;
    const int rblocks = (p.h + p.rb - 1) / p.rb;
    const int n = blockIdx.x / rblocks, ys = (blockIdx.x - n * rblocks) * p.rb;
    const int ye = min(p.h, ys + p.rb);
    const float ry = (float)(p.h - 1) / (float)(H2 - 1), rx = (float)(p.w - 1) / (float)(W2 - 1);
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], (uint32_t)cwarps);
        }
        mbar_init_fence();
    }
    __syncthreads();
    const int Ystart = max(0, 2 * ys - 4);
    if (warp == cwarps) {
        // ---- producer: rows Ystart .. (last row whose upper source row is below ye), one bulk copy each
        if (elect_one()) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.dx) + (size_t)n * H2 * row_bytes;
            int s = 0;
            uint32_t ph = 1;
            for (int Y = Ystart; Y < H2; ++Y) {
                if ((int)(ry * (float)Y) >= ye) break;
                mbar_wait(&empty[s], ph);
                mbar_arrive_expect_tx(&full[s], row_bytes);
                bulk_g2s(smem + (size_t)s * row_bytes, src + (size_t)Y * row_bytes, row_bytes, &full[s]);
                if (++s == p.S) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }
    const int item = threadIdx.x;
    const bool active = item < items;
    const int cg = active ? item % cgs : 0, x = active ? item / cgs : 0;
    // the <= kTaps consecutive output columns that read source column x (as in the streaming kernel)
    float wx[kTaps];
    int X0 = 0;
    {
        float wk[8];
        int k0 = 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int D = 2 * x - 3 + k;
            wk[k] = (D >= 0 && D < W2) ? upw(D, x, p.w, rx) : 0.f;
            if (wk[k] != 0.f && k0 == 8) k0 = k;
        }
        if (k0 > 8 - kTaps) k0 = 8 - kTaps;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (k == k0 + j) v = wk[k];
            wx[j] = v;
        }
        X0 = 2 * x - 3 + k0;
#pragma unroll
        for (int sft = 0; sft < 3; ++sft)
            if (X0 < 0) {
#pragma unroll
                for (int j = 0; j < kTaps - 1; ++j) wx[j] = wx[j + 1];
                wx[kTaps - 1] = 0.f;
                ++X0;
            }
        // taps past the right edge carry zero weight by construction: point them at the last column instead of predicating the loads
    }
    uint32_t toff[kTaps];
#pragma unroll
    for (int j = 0; j < kTaps; ++j) toff[j] = (uint32_t)(min(X0 + j, W2 - 1) * p.C + cg * 8) * 2u;
    __nv_bfloat16* dst = p.out + (size_t)n * p.h * p.w * p.C + (size_t)x * p.C + cg * 8;
    float acc_a[8], acc_b[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { acc_a[e] = 0.f; acc_b[e] = 0.f; }
    int cur = (int)(ry * (float)Ystart);
    int s = 0;
    uint32_t ph = 0;
    const uint8_t* stage = smem;
    for (int Y = Ystart; Y < H2; ++Y) {
        const float sy = ry * (float)Y;
        const int y0 = (int)sy;
        if (y0 >= ye) break;
        const int y1 = y0 + (y0 < p.h - 1);
        const float ly = sy - (float)y0;
        mbar_wait(&full[s], ph);
        uint4 r[kTaps];
#pragma unroll
        for (int j = 0; j < kTaps; ++j) r[j] = *reinterpret_cast<const uint4*>(stage + toff[j]);
        float xr[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) xr[e] = 0.f;
#pragma unroll
        for (int j = 0; j < kTaps; ++j) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&r[j]);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(hh[e]);
                xr[2 * e] = fmaf(wx[j], f.x, xr[2 * e]);
                xr[2 * e + 1] = fmaf(wx[j], f.y, xr[2 * e + 1]);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);      // the stage has been read into registers by this warp
        stage += row_bytes;
        if (++s == p.S) { s = 0; ph ^= 1u; stage = smem; }
        if (y0 > cur) {   // row `cur` is complete
            if (cur >= ys && active) {
                uint4 o;
                __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
                for (int e = 0; e < 4; ++e) q[e] = __floats2bfloat162_rn(acc_a[2 * e], acc_a[2 * e + 1]);
                *reinterpret_cast<uint4*>(dst + (size_t)cur * p.w * p.C) = o;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) { acc_a[e] = acc_b[e]; acc_b[e] = 0.f; }
            cur = y0;
        }
        const float w0 = 1.f - ly;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc_a[e] = fmaf(w0, xr[e], acc_a[e]);
        if (y1 == y0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc_a[e] = fmaf(ly, xr[e], acc_a[e]);
        } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) acc_b[e] = fmaf(ly, xr[e], acc_b[e]);
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int row = cur + t;
        if (row >= ys && row < ye && active) {
            uint4 o;
            __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
            for (int e = 0; e < 4; ++e) q[e] = t == 0 ? __floats2bfloat162_rn(acc_a[2 * e], acc_a[2 * e + 1]) : __floats2bfloat162_rn(acc_b[2 * e], acc_b[2 * e + 1]);
            *reinterpret_cast<uint4*>(dst + (size_t)row * p.w * p.C) = o;
        }
    }
}

}  // namespace nunet

using namespace nunet;

extern "C" int nunet_up_transpose(const void* dx, int32_t pitch, int32_t coff, int32_t N, int32_t h, int32_t w, int32_t C, void* out,
                                  void* stream) {
    if (!dx || !out || N < 1 || h < 1 || w < 1 || C < 8 || C % 8 != 0 || pitch % 8 != 0 || coff % 8 != 0) return NUNET_E_BADARG;
    const long long tiles = (long long)N * ((h + kUT - 1) / kUT) * ((w + kUT - 1) / kUT);
    if (tiles >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    UpTParams p{(const __nv_bfloat16*)dx, pitch, coff, N, h, w, C, (__nv_bfloat16*)out};
    static const bool tiled = getenv("NUNET_UPT_TILED") != nullptr;
    static const bool no_bulk = getenv("NUNET_UPT_NO_BULK") != nullptr;
    static const int rb_env = getenv("NUNET_UPT_RB") ? atoi(getenv("NUNET_UPT_RB")) : 0;
    const long long row_bytes = 2LL * w * C * 2;
    if (!tiled && !no_bulk && pitch == C && coff == 0 && w * (C / 8) <= 992 && row_bytes <= 24 * 1024 && ((uintptr_t)dx & 15) == 0) {
        UpTBulkParams b{(const __nv_bfloat16*)dx, N, h, w, C, rb_env > 0 ? rb_env : kRBBulk, 0, (__nv_bfloat16*)out};
        b.S = (int)(72 * 1024 / row_bytes);
        if (b.S > 8) b.S = 8;
        const long long bx = (long long)N * ((h + b.rb - 1) / b.rb);
        if (bx >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
        const size_t smem = (size_t)b.S * row_bytes + 2 * b.S * 8;
        static bool attr_done = false;
        if (!attr_done) {
            cudaError_t e = cudaFuncSetAttribute(up_transpose_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
            attr_done = true;
        }
        const int threads = ((w * (C / 8) + 31) / 32) * 32 + 32;
        up_transpose_bulk_kernel<<<(unsigned)bx, threads, smem, (cudaStream_t)stream>>>(b);
        cudaError_t e0 = cudaGetLastError();
        return e0 == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e0;
    }
    if (!tiled) {
        const long long bx = (long long)N * ((h + kRB - 1) / kRB);
        if (bx >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
        up_transpose_stream_kernel<<<dim3((unsigned)bx, (w * (C / 8) + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
        cudaError_t e0 = cudaGetLastError();
        return e0 == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e0;
    }
    const size_t smem = (size_t)kUW * kUW * (kUC / 8) * 16 + (size_t)kUW * kUT * kUC * 4 + 2 * kUT * kTaps * 4 + 2 * kUT * 4;
    cudaError_t e = cudaFuncSetAttribute(up_transpose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    up_transpose_kernel<<<dim3((unsigned)tiles, (C + kUC - 1) / kUC), 256, smem, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
