// Operand loaders shared by the forward/dgrad and wgrad kernels.  One "item" is the 8 bf16 values (one 16-byte
// core-matrix row) of a concatenated source at one output pixel; producing it applies the producer's BatchNorm+ReLU and
// the skip pathway's 2x2 max-pool / align_corners bilinear 2x upsample on the fly (or, for backward operands, the
// folded BatchNorm-backward affine).  Loads are split into issue() / finish() so a thread keeps several items' global
// loads in flight before it transforms any of them (the loaders are latency-bound otherwise).
#pragma once
#include "conv_common.cuh"

namespace nunet {

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
__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
    uint4 o;
    o.x = pack_bf16x2(a[0], a[1]);
    o.y = pack_bf16x2(a[2], a[3]);
    o.z = pack_bf16x2(a[4], a[5]);
    o.w = pack_bf16x2(a[6], a[7]);
    return o;
}
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// relu(scale*y+shift) on 8 channels
__device__ __forceinline__ void bn_relu8(const uint4& raw, const float (&sc)[8], const float (&sh)[8], float (&a)[8]) {
    float y[8];
    unpack8(raw, y);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaxf(fmaf(y[i], sc[i], sh[i]), 0.f);
}

// Number of 16-byte global loads one item of this kind needs.
template <int KIND>
struct LoadsOf {
    static constexpr int n = (KIND == NUNET_SRC_SAME || KIND == NUNET_SRC_ACT) ? 1 : (KIND == NUNET_SRC_GRAD || KIND == NUNET_SRC_INPUT) ? 2 : 4;
};

struct UpTaps {  // bilinear taps of one output pixel (ATen area_pixel_compute_scale, align_corners=True)
    int o00, o01, o10, o11;  // source pixel indices
    float ly0, ly1, lx0, lx1;
};
__device__ __forceinline__ UpTaps up_taps(int H, int W, PixInfo pi) {
    const int y = pi.yx >> 16, x = pi.yx & 0xffff;
    const int h = H >> 1, w = W >> 1;
    const float ry = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.f;
    const float rx = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const float sy = ry * (float)y, sx = rx * (float)x;
    const int y0 = (int)sy, x0 = (int)sx;
    const int y1 = y0 + (y0 < h - 1), x1 = x0 + (x0 < w - 1);
    UpTaps t;
    t.ly1 = sy - (float)y0;
    t.lx1 = sx - (float)x0;
    t.ly0 = 1.f - t.ly1;
    t.lx0 = 1.f - t.lx1;
    const int nb = (pi.pix - y * W - x) >> 2;  // n*h*w
    t.o00 = nb + y0 * w + x0;
    t.o01 = nb + y0 * w + x1;
    t.o10 = nb + y1 * w + x0;
    t.o11 = nb + y1 * w + x1;
    return t;
}

// Issues the global loads of one item (pi.pix >= 0 required).
template <int KIND>
__device__ __forceinline__ void issue_loads(const ConvSrc& s, int H, int W, PixInfo pi, int ch, uint4 (&r)[LoadsOf<KIND>::n]) {
    const int C = s.C;
    if constexpr (KIND == NUNET_SRC_SAME || KIND == NUNET_SRC_ACT) {
        r[0] = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data) + (size_t)pi.pix * C + ch);
    } else if constexpr (KIND == NUNET_SRC_POOL || KIND == NUNET_SRC_POOL_ACT) {
        const int x = pi.yx & 0xffff;
        const size_t base = (size_t)4 * pi.pix - 2 * x;  // (n*2H + 2y)*2W + 2x
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s.data) + base * C + ch;
        r[0] = ldg16(p);
        r[1] = ldg16(p + C);
        r[2] = ldg16(p + (size_t)2 * W * C);
        r[3] = ldg16(p + (size_t)(2 * W + 1) * C);
    } else if constexpr (KIND == NUNET_SRC_UP || KIND == NUNET_SRC_UP_ACT) {
        const UpTaps t = up_taps(H, W, pi);
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s.data) + ch;
        r[0] = ldg16(p + (size_t)t.o00 * C);
        r[1] = ldg16(p + (size_t)t.o01 * C);
        r[2] = ldg16(p + (size_t)t.o10 * C);
        r[3] = ldg16(p + (size_t)t.o11 * C);
    } else if constexpr (KIND == NUNET_SRC_GRAD) {
        r[0] = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data) + (size_t)pi.pix * C + ch);
        r[1] = ldg16(reinterpret_cast<const __nv_bfloat16*>(s.data2) + (size_t)pi.pix * C + ch);
    } else {  // NUNET_SRC_INPUT: fp32 NCHW, C real channels (<= 16); r holds 8 floats
        const int HW = H * W;
        const int n = pi.pix / HW;
        const int rem = pi.pix - n * HW;
        const float* p = reinterpret_cast<const float*>(s.data) + (size_t)n * C * HW + rem;
        float f[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = (ch + i < C) ? __ldg(p + (size_t)(ch + i) * HW) : 0.f;
        r[0] = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
        r[1] = make_uint4(__float_as_uint(f[4]), __float_as_uint(f[5]), __float_as_uint(f[6]), __float_as_uint(f[7]));
    }
}

// Transforms the loaded values into the 8 bf16 operand values.
template <int KIND>
__device__ __forceinline__ uint4 finish_item(int H, int W, PixInfo pi, const uint4 (&r)[LoadsOf<KIND>::n], const float (&sc)[8],
                                             const float (&sh)[8], const float (&k2)[8]) {
    float a[8];
    if constexpr (KIND == NUNET_SRC_SAME) {
        bn_relu8(r[0], sc, sh, a);
    } else if constexpr (KIND == NUNET_SRC_POOL) {
        float b[8];
        bn_relu8(r[0], sc, sh, a);
#pragma unroll
        for (int k = 1; k < 4; ++k) {
            bn_relu8(r[k], sc, sh, b);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaxf(a[i], b[i]);
        }
    } else if constexpr (KIND == NUNET_SRC_UP) {
        const UpTaps t = up_taps(H, W, pi);
        float b[8], c[8], d[8];
        bn_relu8(r[0], sc, sh, a);
        bn_relu8(r[1], sc, sh, b);
        bn_relu8(r[2], sc, sh, c);
        bn_relu8(r[3], sc, sh, d);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = t.ly0 * (t.lx0 * a[i] + t.lx1 * b[i]) + t.ly1 * (t.lx0 * c[i] + t.lx1 * d[i]);
    } else if constexpr (KIND == NUNET_SRC_ACT) {
        return r[0];
    } else if constexpr (KIND == NUNET_SRC_POOL_ACT) {
        // max of four already-activated bf16 values: exact in bf16x2 arithmetic
        const __nv_bfloat162* h0 = reinterpret_cast<const __nv_bfloat162*>(&r[0]);
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 m = h0[i];
#pragma unroll
            for (int k = 1; k < 4; ++k) m = __hmax2(m, reinterpret_cast<const __nv_bfloat162*>(&r[k])[i]);
            ow[i] = *reinterpret_cast<uint32_t*>(&m);
        }
        return o;
    } else if constexpr (KIND == NUNET_SRC_UP_ACT) {
        const UpTaps t = up_taps(H, W, pi);
        float b[8], c[8], d[8];
        unpack8(r[0], a);
        unpack8(r[1], b);
        unpack8(r[2], c);
        unpack8(r[3], d);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = t.ly0 * (t.lx0 * a[i] + t.lx1 * b[i]) + t.ly1 * (t.lx0 * c[i] + t.lx1 * d[i]);
    } else if constexpr (KIND == NUNET_SRC_GRAD) {
        float g[8], y[8];
        unpack8(r[0], g);
        unpack8(r[1], y);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(g[i], sc[i], fmaf(y[i], k2[i], sh[i]));
    } else {
        a[0] = __uint_as_float(r[0].x); a[1] = __uint_as_float(r[0].y); a[2] = __uint_as_float(r[0].z); a[3] = __uint_as_float(r[0].w);
        a[4] = __uint_as_float(r[1].x); a[5] = __uint_as_float(r[1].y); a[6] = __uint_as_float(r[1].z); a[7] = __uint_as_float(r[1].w);
    }
    return pack8(a);
}

// Stages `nblk` blocks of 8 positions for one 8-channel group: position index of block b, lane-local pl is
// pos0 + b*8 + pl (must be < plimit to be staged); `tab` is indexed by (position + tab_off); the 16-byte result goes to
// dst + position*16.  Blocks are distributed round-robin over `nwarps` warps; U items are in flight per thread.
template <int KIND, int U>
__device__ __forceinline__ void stage_blocks(const ConvSrc& s, int H, int W, const PixInfo* __restrict__ tab, int tab_off, int plimit,
                                             int first_blk, int nblk, int blk_stride, int pl, int ch, const float* __restrict__ psc,
                                             const float* __restrict__ psh, const float* __restrict__ pk2, uint8_t* dst) {
    constexpr int NL = LoadsOf<KIND>::n;
    constexpr bool kCoef = (KIND == NUNET_SRC_SAME || KIND == NUNET_SRC_POOL || KIND == NUNET_SRC_UP || KIND == NUNET_SRC_GRAD);
    float sc[8], sh[8], k2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        sc[i] = kCoef ? psc[i] : 1.f;
        sh[i] = kCoef ? psh[i] : 0.f;
        k2[i] = (KIND == NUNET_SRC_GRAD) ? pk2[i] : 0.f;
    }
    for (int b0 = first_blk; b0 < nblk; b0 += blk_stride * U) {
        PixInfo pi[U];
        uint4 r[U][NL];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = (b0 + u * blk_stride) * 8 + pl;
            pi[u].pix = -1;
            pi[u].yx = 0;
            if (b0 + u * blk_stride < nblk && pp < plimit) pi[u] = tab[pp + tab_off];
            if (pi[u].pix >= 0) issue_loads<KIND>(s, H, W, pi[u], ch, r[u]);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int pp = (b0 + u * blk_stride) * 8 + pl;
            if (b0 + u * blk_stride < nblk && pp < plimit) {
                uint4 v = make_uint4(0, 0, 0, 0);
                if (pi[u].pix >= 0) v = finish_item<KIND>(H, W, pi[u], r[u], sc, sh, k2);
                *reinterpret_cast<uint4*>(dst + (size_t)pp * 16) = v;
            }
        }
    }
}

// Same-resolution pre-activated source: pure copy, issued as cp.async (no registers, many copies in flight per thread;
// padding positions are zero-filled by a zero-byte source).  The caller commits / waits the cp.async group.
__device__ __forceinline__ void stage_blocks_cpasync(const ConvSrc& s, const PixInfo* __restrict__ tab, int tab_off, int plimit,
                                                     int first_blk, int nblk, int blk_stride, int pl, int ch, uint8_t* dst) {
    const __nv_bfloat16* base = reinterpret_cast<const __nv_bfloat16*>(s.data) + ch;
    const int C = s.C;
    for (int b = first_blk; b < nblk; b += blk_stride * 4) {
        int pix[4];  // -2: nothing to stage, -1: padding (zero fill)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int bb = b + u * blk_stride, pp = bb * 8 + pl;
            pix[u] = (bb < nblk && pp < plimit) ? tab[pp + tab_off].pix : -2;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int pp = (b + u * blk_stride) * 8 + pl;
            if (pix[u] != -2) cp_async16(dst + (size_t)pp * 16, base + (size_t)(pix[u] >= 0 ? pix[u] : 0) * C, pix[u] >= 0 ? 16u : 0u);
        }
    }
}

// Source-kind combinations that occur in the network; the conv kernels are specialised on them so that only the loader
// variants a launch can need are compiled into it (keeps the producer loop in registers).
enum { MODE_INPUT = 0, MODE_SAME = 1, MODE_POOL = 2, MODE_SAME_UP = 3, MODE_GRAD = 4, MODE_ACT = 5, MODE_POOL_ACT = 6, MODE_ACT_UP = 7,
       MODE_ACT_TMA = 8 /* all sources pre-activated, same resolution, staged by TMA tensor loads (conv_igemm only) */ };

template <int MODE>
__device__ __forceinline__ void stage_blocks_mode(const ConvSrc& s, int H, int W, const PixInfo* __restrict__ tab, int tab_off, int plimit,
                                                  int first_blk, int nblk, int blk_stride, int pl, int ch, const float* __restrict__ sc,
                                                  const float* __restrict__ sh, const float* __restrict__ k2, uint8_t* dst) {
    if constexpr (MODE == MODE_INPUT) {
        stage_blocks<NUNET_SRC_INPUT, 2>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    } else if constexpr (MODE == MODE_SAME) {
        stage_blocks<NUNET_SRC_SAME, 4>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    } else if constexpr (MODE == MODE_POOL) {
        stage_blocks<NUNET_SRC_POOL, 2>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    } else if constexpr (MODE == MODE_GRAD) {
        stage_blocks<NUNET_SRC_GRAD, 4>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    } else if constexpr (MODE == MODE_ACT || MODE == MODE_ACT_TMA) {
        stage_blocks_cpasync(s, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, dst);
    } else if constexpr (MODE == MODE_POOL_ACT) {
        stage_blocks<NUNET_SRC_POOL_ACT, 2>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    } else if constexpr (MODE == MODE_ACT_UP) {
        if (s.kind == NUNET_SRC_UP_ACT)
            stage_blocks<NUNET_SRC_UP_ACT, 2>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
        else
            stage_blocks_cpasync(s, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, dst);
    } else {
        if (s.kind == NUNET_SRC_UP)
            stage_blocks<NUNET_SRC_UP, 2>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
        else
            stage_blocks<NUNET_SRC_SAME, 4>(s, H, W, tab, tab_off, plimit, first_blk, nblk, blk_stride, pl, ch, sc, sh, k2, dst);
    }
}

// Host side: the mode of a source list, or -1 if the combination is not supported.
static inline int sources_mode(const int* kinds, int n) {
    if (n == 1) {
        switch (kinds[0]) {
            case NUNET_SRC_INPUT: return MODE_INPUT;
            case NUNET_SRC_SAME: return MODE_SAME;
            case NUNET_SRC_POOL: return MODE_POOL;
            case NUNET_SRC_GRAD: return MODE_GRAD;
            case NUNET_SRC_UP: return MODE_SAME_UP;
            case NUNET_SRC_ACT: return MODE_ACT;
            case NUNET_SRC_POOL_ACT: return MODE_POOL_ACT;
            case NUNET_SRC_UP_ACT: return MODE_ACT_UP;
        }
        return -1;
    }
    if (kinds[0] == NUNET_SRC_ACT || kinds[0] == NUNET_SRC_UP_ACT) {
        for (int i = 0; i < n; ++i) {
            if (kinds[i] == NUNET_SRC_ACT) continue;
            if (kinds[i] == NUNET_SRC_UP_ACT && i == n - 1) continue;
            return -1;
        }
        return (kinds[n - 1] == NUNET_SRC_UP_ACT) ? MODE_ACT_UP : MODE_ACT;
    }
    for (int i = 0; i < n; ++i) {
        const bool last = (i == n - 1);
        if (kinds[i] == NUNET_SRC_SAME) continue;
        if (kinds[i] == NUNET_SRC_UP && last) continue;
        return -1;
    }
    return (kinds[n - 1] == NUNET_SRC_UP) ? MODE_SAME_UP : MODE_SAME;
}

// Decodes a flattened padded position q = R*Wp + c into the output pixel it denotes (pix = -1 on padding).
__device__ __forceinline__ PixInfo decode_position(int q, int N, int H, int W, int Wp, int Hp1, int Q) {
    PixInfo pi;
    pi.pix = -1;
    pi.yx = 0;
    if (q >= 0 && q < Q) {
        const int R = q / Wp, c = q - R * Wp;
        const int n = R / Hp1, rr = R - n * Hp1;
        if (c >= 1 && c <= W && rr >= 1 && n < N) {
            pi.pix = (n * H + (rr - 1)) * W + (c - 1);
            pi.yx = ((rr - 1) << 16) | (c - 1);
        }
    }
    return pi;
}

__device__ __forceinline__ int decode_pix_fast(int q, int N, int H, int W, int Wp, int Hp1, int Q, const FastDiv& dWp, const FastDiv& dHp1) {
    if (q < 0 || q >= Q) return -1;
    const int R = (int)fastdiv((uint32_t)q, dWp), c = q - R * Wp;
    const int n = (int)fastdiv((uint32_t)R, dHp1), rr = R - n * Hp1;
    return (c >= 1 && c <= W && rr >= 1 && n < N) ? (n * H + (rr - 1)) * W + (c - 1) : -1;
}
// Column strips (rows wider than a tensor-load box): "image" v of the flattened geometry is strip v % strips of image v / strips.
__device__ __forceinline__ int decode_pix_strips(int q, int Nv, int H, int Ws, int Wp, int Hp1, int Q, const FastDiv& dWp, const FastDiv& dHp1,
                                                 int strips, int Wreal, const FastDiv& dStrips) {
    if (q < 0 || q >= Q) return -1;
    const int R = (int)fastdiv((uint32_t)q, dWp), c = q - R * Wp;
    const int v = (int)fastdiv((uint32_t)R, dHp1), rr = R - v * Hp1;
    const int n = (int)fastdiv((uint32_t)v, dStrips), s = v - n * strips;
    const int x = s * Ws + c - 1;
    return (c >= 1 && c <= Ws && x < Wreal && rr >= 1 && v < Nv) ? (n * H + (rr - 1)) * Wreal + x : -1;
}

}  // namespace nunet
