// Shared declarations of the conv kernels (kernel parameter blocks mirror include/nunet_b200.h descriptors).
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "bn_tail.cuh"
#include "../../include/nunet_b200.h"

namespace nunet {

struct ConvSrc {
    const void* data;
    const void* data2;
    const float* scale;
    const float* shift;
    const float* k2;
    int C;
    int kind;
};

// Unsigned division by a launch-invariant divisor d >= 2 (Granlund-Montgomery / Hacker's Delight 10-10): exact for every 32-bit n.
// The epilogue decodes one position per lane and tile; the two `/` of decode_position were ~15 % of its instructions.
struct FastDiv {
    uint32_t m, s;   // s = ceil(log2 d) >= 1, m = floor(2^32 * (2^s - d) / d) + 1
};
__host__ inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t s = 1;
    while ((1ull << s) < d) ++s;
    f.s = s;
    f.m = (uint32_t)((((1ull << s) - d) << 32) / d + 1);
    return f;
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t n, const FastDiv& f) {
    const uint32_t t = __umulhi(f.m, n);
    return (t + ((n - t) >> 1)) >> (f.s - 1);
}

struct ConvParams {
    CUtensorMap tmap[NUNET_MAX_SRC];  // TMA mode only: one tiled map per pre-activated source (box = 32 ch x (W+2) px, swizzle 64B)
    int N, H, W;
    int Wp, Hp1;  // W+2, H+1
    FastDiv div_wp, div_hp1;   // position decode in the TMA-mode epilogue
    int strips, Wreal;         // strips > 1: N / W above describe N*strips column strips of the Wreal-pixel rows (TMA mode only)
    FastDiv div_strips;
    int Q;        // padded positions: (N*(H+1)+1)*(W+2)
    int T, NT, CK, S, P, nchunks, cin_pad, Cout, num_src, tmem_cols, ntg;
    int res_b;          // 1: single K chunk, the weight block is fetched once per CTA and stays resident
    int a_stage_bytes;  // TMA mode: bytes of one A stage ((P + Wp) rows of 64 B, rounded to 1 KB)
    int epi_warps;      // TMA mode: warps that drain TMEM (16, or 8 where the transpose scratch of 16 does not fit)
    ConvSrc src[NUNET_MAX_SRC];
    const void* wpack;
    void* out;
    float* stat_part;
    BnTail tail;        // BatchNorm finalisation by the last CTA (tail.counter != nullptr)
    // fused block-internal gradient router (dgrad of a block's second conv): out = dX * [mask_scale*mask_y + mask_shift > 0]
    const void* mask_y;
    const float *mask_scale, *mask_shift, *mask_mean, *mask_invstd;
    // plain epilogue: output channels scattered to one dense tensor per concatenated source of the forward conv
    void* out_seg[NUNET_MAX_SRC];
    int out_seg_c[NUNET_MAX_SRC];
    int num_out_seg;
};

// host: C-ABI descriptor -> kernel parameter block
static inline void copy_exchange(BnExchange& x, const nunet_bn_exchange_t& e) {
    x.world = e.world; x.rank = e.rank; x.seq = e.seq;
    for (int i = 0; i < 8; ++i) { x.data[i] = e.data[i]; x.flags[i] = e.flags[i]; }
}
static inline bool exchange_ok(const nunet_bn_exchange_t& e) {
    if (e.world <= 1) return true;
    if (e.world > 8 || e.rank < 0 || e.rank >= e.world || !e.seq) return false;
    for (int i = 0; i < e.world; ++i)
        if (!e.data[i] || !e.flags[i]) return false;
    return true;
}
static inline BnTail make_bn_tail(const nunet_bn_tail_t& b) {
    BnTail t{};
    copy_exchange(t.xch, b.xch);
    t.counter = b.counter;
    t.sums_out = b.sums_out;
    t.fin.mode = 0; t.fin.count = b.count; t.fin.gamma = b.gamma; t.fin.beta = b.beta; t.fin.conv_bias = b.conv_bias;
    t.fin.running_mean = b.running_mean; t.fin.running_var = b.running_var; t.fin.momentum = b.momentum; t.fin.eps = b.eps;
    t.fin.scale = b.scale; t.fin.shift = b.shift; t.fin.save_mean = b.save_mean; t.fin.save_invstd = b.save_invstd;
    return t;
}
static inline bool bn_tail_ok(const nunet_bn_tail_t& b, const float* stat_part) {
    if (!b.counter) return true;
    if (!stat_part || b.count <= 0 || !exchange_ok(b.xch)) return false;
    if (b.gamma) return b.beta && b.scale && b.shift;
    return b.sums_out != nullptr;
}

}  // namespace nunet
