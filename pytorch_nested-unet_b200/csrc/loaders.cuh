// Operand loaders shared by the forward/dgrad and wgrad kernels: one call produces the 8 bf16 values (one 16-byte
// core-matrix row) of a concatenated source at one output pixel, applying the producer's BatchNorm+ReLU and the skip
// pathway's pool / upsample on the fly (or, for backward operands, the folded BatchNorm-backward affine).
#pragma once
#include "conv_common.cuh"

namespace nunet {

// ------------------------------------------------------------------------------------------------ loaders
struct PixInfo {
    int pix;  // (n*H + y)*W + x at OUTPUT resolution, or -1
    int yx;   // (y << 16) | x
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// relu(scale*y+shift) on 8 channels
__device__ __forceinline__ void bn_relu8(const uint4& raw, const float (&sc)[8], const float (&sh)[8], float (&a)[8]) {
    float y[8];
    unpack8(raw, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(fmaf(y[i], sc[i], sh[i]), 0.f);
}

// Produces the 8 bf16 values (one 16-byte core-matrix row) of source `s`, channels [ch, ch+8), at output pixel `pi`.
__device__ __forceinline__ uint4 load_transform(const ConvSrc& s, int H, int W, PixInfo pi, int ch, const float (&sc)[8],
                                                const float (&sh)[8], const float (&k2)[8]) {
    float a[8];
    const int C = s.C;
    if (s.kind == NUNET_SRC_SAME) {
        uint4 raw = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data) + (size_t)pi.pix * C + ch);
        bn_relu8(raw, sc, sh, a);
    } else if (s.kind == NUNET_SRC_POOL) {
        const int x = pi.yx & 0xffff;
        const size_t base = (size_t)4 * pi.pix - 2 * x;  // (n*2H + 2y)*2W + 2x
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s.data) + base * C + ch;
        uint4 r0 = ldg16(p), r1 = ldg16(p + C), r2 = ldg16(p + (size_t)2 * W * C), r3 = ldg16(p + (size_t)(2 * W + 1) * C);
        float b[8];
        bn_relu8(r0, sc, sh, a);
        bn_relu8(r1, sc, sh, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
        bn_relu8(r2, sc, sh, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
        bn_relu8(r3, sc, sh, b);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
    } else if (s.kind == NUNET_SRC_UP) {
        const int y = pi.yx >> 16, x = pi.yx & 0xffff;
        const int h = H >> 1, w = W >> 1;
        // ATen area_pixel_compute_scale (align_corners=True): scale = (in-1)/(out-1), src = scale*dst
        const float ry = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
        const float rx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
        const float sy = ry * (float)y, sx = rx * (float)x;
        const int y0 = (int)sy, x0 = (int)sx;
        const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
        const float ly1 = sy - (float)y0, lx1 = sx - (float)x0;
        const float ly0 = 1.f - ly1, lx0 = 1.f - lx1;
        const size_t nb = (size_t)((pi.pix - y * W - x) >> 2);  // n*h*w
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s.data) + ch;
        uint4 r00 = ldg16(p + (nb + (size_t)y0 * w + x0) * C), r01 = ldg16(p + (nb + (size_t)y0 * w + x1) * C);
        uint4 r10 = ldg16(p + (nb + (size_t)y1 * w + x0) * C), r11 = ldg16(p + (nb + (size_t)y1 * w + x1) * C);
        float b[8], c[8], d[8];
        bn_relu8(r00, sc, sh, a);
        bn_relu8(r01, sc, sh, b);
        bn_relu8(r10, sc, sh, c);
        bn_relu8(r11, sc, sh, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ly0 * (lx0 * a[i] + lx1 * b[i]) + ly1 * (lx0 * c[i] + lx1 * d[i]);
    } else if (s.kind == NUNET_SRC_GRAD) {
        uint4 rg = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data) + (size_t)pi.pix * C + ch);
        uint4 ry = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data2) + (size_t)pi.pix * C + ch);
        float g[8], y[8];
        unpack8(rg, g);
        unpack8(ry, y);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(g[i], sc[i], fmaf(y[i], k2[i], sh[i]));
    } else {  // NUNET_SRC_INPUT: fp32 NCHW, C real channels (<= 16)
        const int HW = H * W;
        const int n = pi.pix / HW;
        const int rem = pi.pix - n * HW;
        const float* p = reinterpret_cast<const float*>(s.data) + (size_t)n * C * HW + rem;
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = (ch + i < C) ? __ldg(p + (size_t)(ch + i) * HW) : 0.f;
    }
    uint4 o;
    o.x = pack_bf16x2(a[0], a[1]);
    o.y = pack_bf16x2(a[2], a[3]);
    o.z = pack_bf16x2(a[4], a[5]);
    o.w = pack_bf16x2(a[6], a[7]);
    return o;
}


// Decodes a flattened padded position q = R*Wp + c into the output pixel it denotes (pix = -1 on padding).
__device__ __forceinline__ PixInfo decode_position(long long q, int N, int H, int W, int Wp, int Hp1, long long Q) {
    PixInfo pi;
    pi.pix = -1;
    pi.yx = 0;
    if (q >= 0 && q < Q) {
        const int R = (int)(q / Wp), c = (int)(q - (long long)R * Wp);
        const int n = R / Hp1, rr = R - n * Hp1;
        if (c >= 1 && c <= W && rr >= 1 && n < N) {
            pi.pix = (n * H + (rr - 1)) * W + (c - 1);
            pi.yx = ((rr - 1) << 16) | (c - 1);
        }
    }
    return pi;
}

}  // namespace nunet
