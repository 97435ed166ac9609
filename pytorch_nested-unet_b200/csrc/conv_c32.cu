// 3x3 convolution with 32 OUTPUT channels over pre-activated bf16 NHWC sources (the level-0 layers of the network: forward
// convs of the full-resolution nodes, and the dgrad of their second convs), TMA-fed, kernel columns stacked into N = 96.
//
// Why a second conv kernel.  With Cout = 32 the generic implicit GEMM (conv_igemm.cu) issues nine 128x32x16 MMAs per 16 input
// channels and tile; each costs ~45 cycles because the 4 KB A operand is re-read from shared memory for every tap
// (profiles/r01_umma_issue_rate.log): 36 % of the tensor peak at best, and those layers are the largest of the step.  Here the
// three kernel COLUMNS share one A read:
//     z_dc[q][co] = sum_{dr,ci} X[q + (dr-1)*PW][ci] * W[dr][dc][ci][co]          (3 MMAs 128x96x16 per 16 channels, ~56 cycles each)
//     y[p][co]    = z_0[p-1][co] + z_1[p][co] + z_2[p+1][co]                      (epilogue: neighbour lanes)
// i.e. N = 96 = (dc, co) on the weight side, and the column shift is applied to the accumulators instead of the operand.
//
// Geometry.  Rows are padded to PW >= W+2 positions; the tensor-load box starts at pixel -1, so position index u of a row
// holds pixel u-1 (index 0 and indices > W are zero: the conv padding, filled by the TMA unit, like the rows above / below the
// image).  A group is `ry` rows of one image = ry*PW consecutive positions, covered by tiles of 128 accumulator rows that
// overlap by two (tile t = positions [126t, 126t+128) of the group; its lanes 1..126 produce outputs), so every output finds
// both neighbours inside its tile.  The A stage of a 32-channel chunk is ONE tensor load: (32 ch, PW px, ry+3 rows), 64B
// swizzle = the K-major SWIZZLE_64B operand with 64-byte rows; tap row dr and tile t are start-address offsets.
// Weights: [chunk][dr][k-group][dc*32+co][8] (NUNET_PACK_*_C32): all chunks fetched once per CTA and kept resident when three
// activation stages still fit beside them, else one 18 KB chunk travels with every stage.
// Warps: 0-15 epilogue (TMEM lane quadrant x channel quarter), 16 MMA issuer + TMEM owner, 17 TMA / weight fetcher.  TMEM holds
// the up to five tiles of a group (480 columns) with one full/empty barrier pair PER TILE: the issuer starts tile t of the next
// group as soon as the epilogue has read tile t of the current one, so the epilogue overlaps the following group's MMAs.
// Epilogue (round 2: no CTA-wide barrier per tile): tcgen05.ld of the three column blocks, the accumulator is RELEASED as soon as
// its values sit in registers, neighbour exchange by warp shuffles; the two rows a quadrant needs from its neighbours travel
// through an eight-deep shared-memory ring with one sequence flag per warp (a warp only ever waits for its two neighbours, and only
// if they are behind), bf16 NHWC store, per-thread BatchNorm sums reduced once at the end (fixed order: deterministic).
// Optional fused block-internal gradient router (dgrad of a block's second conv): out = z * [mask_scale*y1 + mask_shift > 0] and
// the BatchNorm-backward sums, exactly as in conv_igemm.cu.
#include "conv_common.cuh"
#include "loaders.cuh"
#include "tma.cuh"
#include <cstdlib>

namespace nunet {

struct C32Params {
    CUtensorMap xmap[NUNET_MAX_SRC];
    int N, H, W, PW;
    uint32_t pw_magic;           // ceil(2^32 / PW): u / PW == __umulhi(u, pw_magic) for the u < 2^16 of a group
    int strips, Wreal;           // strips > 1: N / W describe N*strips column strips of the Wreal-pixel rows (see nunet_conv_plan_t.strips)
    FastDiv div_strips;
    int ry, gpi, groups_total;   // rows per group, groups per image
    int box_rows, S, nch, num_src;
#ifdef NUNET_EXPERIMENT
    int dbg;                     // experiment builds only (NUNET_C32_DBG): 1 = skip the activation tensor loads, 2 = skip the MMAs (wrong results)
#endif
    int res_b;                   // 1: all weight chunks resident; 0: one chunk per pipeline stage (keeps >= 3 activation stages in flight)
    int srcC[NUNET_MAX_SRC];
    int a_stage_bytes;
    const void* wpack;
    __nv_bfloat16* out;
    float* stat_part;            // [gridDim.x][2][32] or null
    BnTail tail;                 // BatchNorm finalisation by the last CTA
    const __nv_bfloat16* mask_y; // fused block-internal gradient router (see conv_igemm.cu) or null
    const float *mask_scale, *mask_shift, *mask_mean, *mask_invstd;
};

#ifdef NUNET_EXPERIMENT
#define C32_DBG(bit) ((p.dbg & (bit)) != 0)
#else
#define C32_DBG(bit) false
#endif
constexpr int kC32Threads = 576;
constexpr int kC32EpiWarps = 16, kC32MmaWarp = 16, kC32TmaWarp = 17;
constexpr int kC32Ring = 8;                              // exchange slots per warp (> the five tiles a warp can run ahead of a neighbour)
constexpr uint32_t kC32BBytes = 3u * 4u * 96u * 16u;   // one chunk of stacked weights
constexpr int kC32MaxTiles = 5;

__global__ void __launch_bounds__(kC32Threads, 1) conv3x3_c32_kernel(const __grid_constant__ C32Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)p.S * p.a_stage_bytes;                         // [nch] resident weight chunks
    float* s_ex = reinterpret_cast<float*>(sB + (size_t)(p.res_b ? p.nch : p.S) * kC32BBytes);   // [ring][16 warps][2 sides][8]
    float* s_red = s_ex + kC32Ring * kC32EpiWarps * 2 * 8;                     // [16 warps][2][8]
    unsigned* s_flag = reinterpret_cast<unsigned*>(s_red + kC32EpiWarps * 2 * 8);   // [16 warps]: last tile sequence number published
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_flag + kC32EpiWarps);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.S;
    uint64_t* acc_full = bars + 2 * p.S;      // [kC32MaxTiles] MMA -> epilogue, per tile
    uint64_t* acc_empty = acc_full + kC32MaxTiles;
    uint64_t* b_full = acc_empty + kC32MaxTiles;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(b_full + 1);

    if (threadIdx.x < kC32EpiWarps) s_flag[threadIdx.x] = 0u;
    if (warp == kC32MmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < p.S; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            for (int t = 0; t < kC32MaxTiles; ++t) {
                mbar_init(&acc_full[t], 1);
                mbar_init(&acc_empty[t], kC32EpiWarps);
            }
            mbar_init(b_full, 1);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_ptr);
    }
    // the last tile of a group reads a few positions past the loaded rows: keep them finite
    for (uint32_t i = threadIdx.x * 16u; i < (uint32_t)p.S * (uint32_t)p.a_stage_bytes; i += kC32Threads * 16u) *reinterpret_cast<uint4*>(sA + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == kC32TmaWarp) {
        // =============================== tensor loads + weights
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            for (int si = 0; si < p.num_src; ++si) tma_prefetch_desc(&p.xmap[si]);
            const uint32_t a_box_bytes = (uint32_t)p.box_rows * (uint32_t)p.PW * 64u;
            if ((int)blockIdx.x >= p.groups_total) {
                // a CTA without work (tiny problems: the grid is the generic plan's) must not leave bulk copies in flight at exit
            } else if (p.res_b) {
                mbar_arrive_expect_tx(b_full, (uint32_t)p.nch * kC32BBytes);
                for (int c = 0; c < p.nch; ++c)
                    bulk_g2s(sB + (size_t)c * kC32BBytes, reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)c * kC32BBytes, kC32BBytes, b_full);
            } else {
                mbar_arrive(b_full);
            }
            int cc = 0;
            for (int g = blockIdx.x; g < p.groups_total; g += gridDim.x) {
                const int n = g / p.gpi, y0 = (g - n * p.gpi) * p.ry;
                int si = 0, c0 = 0;
                for (int c = 0; c < p.nch; ++c, ++cc) {
                    const int s = cc % p.S;
                    const uint32_t ph = (uint32_t)(cc / p.S) & 1;
                    const int cglob = c * 32;
                    while (cglob >= c0 + p.srcC[si]) { c0 += p.srcC[si]; ++si; }
                    mbar_wait(&empty[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full[s], (C32_DBG(1) ? 0u : a_box_bytes) + (p.res_b ? 0u : kC32BBytes));
                    if (!C32_DBG(1)) {
                        if (p.strips > 1) {   // "image" n = strip n % strips of image n / strips; the box starts one pixel left of the strip
                            const int img = (int)fastdiv((uint32_t)n, p.div_strips), st = n - img * p.strips;
                            tma_load_4d(sA + (size_t)s * p.a_stage_bytes, &p.xmap[si], cglob - c0, st * p.W - 1, y0 - 1, img, &full[s]);
                        } else {
                            tma_load_4d(sA + (size_t)s * p.a_stage_bytes, &p.xmap[si], cglob - c0, -1, y0 - 1, n, &full[s]);
                        }
                    }
                    if (!p.res_b) bulk_g2s(sB + (size_t)s * kC32BBytes, reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)c * kC32BBytes, kC32BBytes, &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == kC32MmaWarp) {
        // =============================== MMA issuer
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            const uint32_t idesc = umma_idesc_bf16(128, 96, kMajorK, kMajorK);
            const uint32_t a_hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B (8 positions), version 1, SWIZZLE_64B
            const uint32_t b_hi = umma_desc_hi(128u);                      // no swizzle: 8-row groups 128 B apart
            const uint32_t pw4 = (uint32_t)p.PW * 4u;                      // one padded row in 16-byte units
            int cc = 0, it = 0;
            uint32_t tile_phase = 0;   // bit t: parity of tile t's barriers (groups may use fewer tiles than others)
            for (int g = blockIdx.x; g < p.groups_total; g += gridDim.x, ++it) {
                const int n = g / p.gpi, y0 = (g - n * p.gpi) * p.ry;
                const int rows = min(p.ry, p.H - y0);
                const int ntile = (rows * p.PW + 125) / 126;
                if (it == 0) mbar_wait(b_full, 0);
                for (int c = 0; c < p.nch; ++c, ++cc) {
                    const int s = cc % p.S;
                    const uint32_t ph = (uint32_t)(cc / p.S) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t a_base = ((smem_u32(sA + (size_t)s * p.a_stage_bytes) >> 4) & 0x3FFFu) | (1u << 16);
                    const uint32_t b_base = umma_desc_lo(smem_u32(sB + (size_t)(p.res_b ? c : s) * kC32BBytes), 96u * 16u);   // LBO = next 8 input channels
                    for (int t = 0; t < ntile; ++t) {
                        if (c == 0) {   // the epilogue has drained this tile's accumulator of the previous group
                            mbar_wait(&acc_empty[t], ((tile_phase >> t) & 1u) ^ 1u);
                            tc_fence_after();
                        }
                        const uint32_t d_addr = tmem_base + (uint32_t)(t * 96);
                        const uint32_t a_t = a_base + (uint32_t)(t * 126) * 4u;    // 64 B per position
#pragma unroll
                        for (int dr = 0; dr < 3; ++dr) {
                            if (C32_DBG(2) && (c | dr) != 0) continue;
                            const uint32_t a_lo = a_t + (uint32_t)dr * pw4;
                            const uint32_t b_lo = b_base + (uint32_t)dr * 384u;     // 4 k-groups x 96 rows x 16 B
                            umma_bf16_lohi(d_addr, a_lo, a_hi, b_lo, b_hi, idesc, (uint32_t)((c | dr) != 0));
                            umma_bf16_lohi(d_addr, a_lo + 2u, a_hi, b_lo + 192u, b_hi, idesc, 1u);   // channels 16..31: +32 B / +2 k-groups
                        }
                        if (c == p.nch - 1) {
                            umma_commit(&acc_full[t]);
                            tile_phase ^= 1u << t;
                        }
                    }
                    umma_commit(&empty[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp < kC32EpiWarps) {
        // =============================== epilogue: warp = cq*4 + lq (TMEM lane quadrant lq, channels cq*8 .. cq*8+7)
        const int lq = warp & 3, cq = warp >> 2;
        const int m = lq * 32 + lane;
        const bool want_stats = p.stat_part != nullptr;
        const bool masked = p.mask_y != nullptr;
        float msc[8], msh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            msc[i] = masked ? p.mask_scale[cq * 8 + i] : 0.f;
            msh[i] = masked ? p.mask_shift[cq * 8 + i] : 0.f;
        }
        float ssum[8], ssq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        unsigned seq = 0;
        uint32_t tile_phase = 0;
        volatile unsigned* vflag = s_flag;
        for (int g = blockIdx.x; g < p.groups_total; g += gridDim.x) {
            const int n = g / p.gpi, y0 = (g - n * p.gpi) * p.ry;
            const int rows = min(p.ry, p.H - y0);
            const int ntile = (rows * p.PW + 125) / 126;
            int img_row0 = (n * p.H + y0) * p.W, row_pitch = p.W, col_lim = p.W;
            if (p.strips > 1) {
                const int img = (int)fastdiv((uint32_t)n, p.div_strips), st = n - img * p.strips;
                img_row0 = (img * p.H + y0) * p.Wreal + st * p.W;
                row_pitch = p.Wreal;
                col_lim = min(p.W, p.Wreal - st * p.W);      // the last strip may be narrower
            }
            for (int t = 0; t < ntile; ++t) {
                ++seq;
                mbar_wait(&acc_full[t], (tile_phase >> t) & 1u);
                tile_phase ^= 1u << t;
                tc_fence_after();
                uint32_t z0[8], z1[8], z2[8];
                const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * 96 + cq * 8);
                tmem_ld8(taddr, z0);
                tmem_ld8(taddr + 32u, z1);
                tmem_ld8(taddr + 64u, z2);
                tmem_ld_wait();
                // the accumulator's values are in registers: hand it back to the MMA issuer before the arithmetic and the stores
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[t]);
                // publish the two rows the neighbouring quadrants need (ring slot seq % 8, then this warp's sequence flag)
                float* ex = s_ex + ((size_t)((seq & (kC32Ring - 1)) * kC32EpiWarps + warp) * 2) * 8;
                if (lane == 31) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) ex[i] = __uint_as_float(z0[i]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) ex[8 + i] = __uint_as_float(z2[i]);
                }
                __syncwarp();
                if (lane == 0) {
                    __threadfence_block();
                    vflag[warp] = seq;
                }
                // this lane's output position inside the group (index u holds pixel u % PW - 1 of row u / PW)
                const int u = t * 126 + m;
                const int row = (int)__umulhi((uint32_t)u, p.pw_magic), px = u - row * p.PW - 1;
                const bool valid = (m >= 1) && (m <= 126) && (row < rows) && (px >= 0) && (px < col_lim);
                const size_t pix = (size_t)(uint32_t)(img_row0 + row * row_pitch + px);       // N*H*W < 2^31 (plan check)
                uint4 ym = make_uint4(0, 0, 0, 0);
                if (masked && valid) ym = __ldg(reinterpret_cast<const uint4*>(p.mask_y + pix * 32 + cq * 8));
                float y[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float l = __shfl_up_sync(0xffffffffu, __uint_as_float(z0[i]), 1);
                    const float r = __shfl_down_sync(0xffffffffu, __uint_as_float(z2[i]), 1);
                    y[i] = l + __uint_as_float(z1[i]);     // lane 0: l is its own z0 (replaced below); lane 31: r likewise
                    z2[i] = __float_as_uint(r);
                }
                if (lane == 0 && lq > 0) {                 // z0 of the previous quadrant's last row
                    unsigned spins = 0;
                    while ((int)(vflag[warp - 1] - seq) < 0) { if (++spins > (1u << 26)) __trap(); }
                    __threadfence_block();
                    const volatile float* src = s_ex + ((size_t)((seq & (kC32Ring - 1)) * kC32EpiWarps + warp - 1) * 2) * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) y[i] = src[i] + __uint_as_float(z1[i]);
                }
                if (lane == 31 && lq < 3) {                // z2 of the next quadrant's first row
                    unsigned spins = 0;
                    while ((int)(vflag[warp + 1] - seq) < 0) { if (++spins > (1u << 26)) __trap(); }
                    __threadfence_block();
                    const volatile float* src = s_ex + ((size_t)((seq & (kC32Ring - 1)) * kC32EpiWarps + warp + 1) * 2) * 8 + 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) z2[i] = __float_as_uint(src[i]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 8; ++i) y[i] += __uint_as_float(z2[i]);
                if (valid) {
                    if (masked) {
                        float yv[8];
                        unpack8(ym, yv);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float zz = fmaf(yv[i], msc[i], msh[i]);
                            y[i] = (zz > 0.f) ? y[i] : 0.f;
                            ssum[i] += y[i];
                            ssq[i] = fmaf(y[i], yv[i], ssq[i]);
                        }
                    } else if (want_stats) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) { ssum[i] += y[i]; ssq[i] = fmaf(y[i], y[i], ssq[i]); }
                    }
                    *reinterpret_cast<uint4*>(p.out + pix * 32 + cq * 8) = pack8(y);
                }
            }
        }
        if (want_stats) {
            // warp butterfly (fixed order), then the four quadrant warps of a channel quarter are summed in a fixed order
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    ssum[i] += __shfl_xor_sync(0xffffffffu, ssum[i], off);
                    ssq[i] += __shfl_xor_sync(0xffffffffu, ssq[i], off);
                }
            }
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { s_red[(warp * 2 + 0) * 8 + i] = ssum[i]; s_red[(warp * 2 + 1) * 8 + i] = ssq[i]; }
            }
            asm volatile("bar.sync 2, 512;" ::: "memory");
            if (threadIdx.x < 32) {
                const int c = threadIdx.x, q = c >> 3, i = c & 7;
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int l4 = 0; l4 < 4; ++l4) { a += s_red[((q * 4 + l4) * 2 + 0) * 8 + i]; b += s_red[((q * 4 + l4) * 2 + 1) * 8 + i]; }
                if (masked) b = p.mask_invstd[c] * (b - p.mask_mean[c] * a);     // sum g_z * yhat
                p.stat_part[((size_t)blockIdx.x * 2 + 0) * 32 + c] = a;
                p.stat_part[((size_t)blockIdx.x * 2 + 1) * 32 + c] = b;
            }
            if (p.tail.counter && threadIdx.x < 256) bn_tail_run(p.tail, p.stat_part, (int)gridDim.x, 32, gridDim.x, threadIdx.x, 3, sA);   // operand stages are idle: free scratch
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kC32MmaWarp) tmem_dealloc<512>(tmem_base);
}

}  // namespace nunet

using namespace nunet;

struct C32Tiling {
    int PW, ry, gpi, box_rows, S, res_b, a_stage_bytes, smem_bytes;
};

// Rows wider than a tensor-load box: N*strips column strips of Ws (even) pixels, exactly as nunet_conv3x3_plan describes them.
static int c32_strips(int W, int* Ws) {
    static const bool no_strips = getenv("NUNET_NO_STRIPS") != nullptr;
    int strips = (W + 2 > 256 && !no_strips) ? (W + 253) / 254 : 1;
    int w = (W + strips - 1) / strips;
    if (strips > 1) w += w & 1;
    *Ws = w;
    return strips;
}

static bool c32_tiling(int H, int W, int nch, C32Tiling* t) {
    { int Ws; c32_strips(W, &Ws); W = Ws; }
    const int PW = (W + 2 + 1) & ~1;
    if (PW > 256) return false;
    const int budget = 226 * 1024;
    double best = 1e30;
    for (int ry = 1; ry <= H && ry <= 64; ++ry) {
        const int ntile = (ry * PW + 125) / 126;
        if (ntile > 5) break;
        // rows read: tap rows 0..2 above/below + the overhang of the last tile
        const int box_rows = (2 * PW + 126 * (ntile - 1) + 128 + PW - 1) / PW;
        if (box_rows > 256) break;
        const int a = ((box_rows * PW * 64) + 1023) & ~1023;
        // weights resident if that still leaves three activation stages in flight (the 56 KB tensor loads need ~2 stages of
        // latency cover), otherwise one weight chunk travels with every stage
        const int misc = kC32Ring * kC32EpiWarps * 2 * 8 * 4 + kC32EpiWarps * 2 * 8 * 4 + kC32EpiWarps * 4 + 256;   // exchange ring, stat slots, flags, barriers
        int res_b = 1, S = 4;
        while (S >= 3 && S * a + nch * (int)kC32BBytes + misc > budget) --S;
        if (S < 3) {
            res_b = 0; S = 4;
            while (S >= 2 && S * (a + (int)kC32BBytes) + misc > budget) --S;
            if (S < 2) continue;
        }
        const int fixed = (res_b ? nch : S) * (int)kC32BBytes + misc;
        const int gpi = (H + ry - 1) / ry;
        const int last_rows = H - (gpi - 1) * ry;
        const double tiles = (double)(gpi - 1) * ntile + (last_rows * PW + 125) / 126;
        double cost = tiles + 0.15 * gpi;           // + the per-group epilogue hand-off (single-buffered accumulators)
        if (S == 2) cost *= 1.08;
        if (cost < best) {
            best = cost;
            t->PW = PW; t->ry = ry; t->gpi = gpi; t->box_rows = box_rows; t->S = S; t->res_b = res_b; t->a_stage_bytes = a;
            t->smem_bytes = S * a + fixed;
        }
    }
    return best < 1e30;
}

// 1 if nunet_conv3x3_fwd will route this conv (all sources NUNET_SRC_ACT) to the stacked kernel, i.e. expects weights packed
// with NUNET_PACK_FWD_C32 / NUNET_PACK_DGRAD_C32.
extern "C" int nunet_conv3x3_c32_eligible(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout) {
    static const bool off = getenv("NUNET_NO_C32") != nullptr;
    C32Tiling t;
    // Single-chunk layers (Cin = 32) stay on the generic kernel: 0.097 ms there against 0.166 ms here at 96x96, batch 256, also with the
    // barrier-free sixteen-warp epilogue of round 2 (tools/ab_epilogue.py; NUNET_C32_MIN_CIN=32 to compare) — with one chunk per group
    // the single accumulator set serialises tensor loads, MMAs and the drain.  Cin >= 96: 0.193 vs 0.226, 0.318 vs 0.37 ms.
    static const int min_cin = getenv("NUNET_C32_MIN_CIN") ? atoi(getenv("NUNET_C32_MIN_CIN")) : 64;
    if (off || Cout != 32 || cin_total < min_cin || cin_total % 32 != 0 || N < 1 || H < 1 || W < 1) return 0;
    return c32_tiling(H, W, cin_total / 32, &t) ? 1 : 0;
}

int nunet_conv_c32_launch(const nunet_conv_desc_t* d, int cin, int grid, const BnTail& tail, cudaStream_t stream) {
    C32Tiling t;
    if (!c32_tiling(d->H, d->W, cin / 32, &t)) return NUNET_E_UNSUPPORTED;
    C32Params p{};
    int Ws;
    const int strips = c32_strips(d->W, &Ws);
    p.strips = strips; p.Wreal = d->W; p.div_strips = make_fastdiv((uint32_t)(strips > 1 ? strips : 2));
    p.N = d->N * strips; p.H = d->H; p.W = Ws; p.PW = t.PW; p.pw_magic = (uint32_t)(((1ULL << 32) + (uint64_t)t.PW - 1) / (uint64_t)t.PW); p.ry = t.ry; p.gpi = t.gpi; p.groups_total = d->N * strips * t.gpi;
    p.box_rows = t.box_rows; p.S = t.S; p.res_b = t.res_b; p.nch = cin / 32; p.num_src = d->num_src; p.a_stage_bytes = t.a_stage_bytes;
    p.wpack = d->wpack; p.out = (__nv_bfloat16*)d->out; p.stat_part = d->stat_part;
    p.tail = tail;
    p.mask_y = (const __nv_bfloat16*)d->mask_y; p.mask_scale = d->mask_scale; p.mask_shift = d->mask_shift; p.mask_mean = d->mask_mean;
    p.mask_invstd = d->mask_invstd;
    for (int i = 0; i < d->num_src; ++i) {
        p.srcC[i] = d->src[i].C;
        if (encode_nhwc_map(&p.xmap[i], d->src[i].data, d->N, d->H, d->W, d->src[i].C, 32, t.PW, 64, t.box_rows, 1) != 0) return NUNET_E_UNSUPPORTED;
    }
    cudaError_t e = cudaFuncSetAttribute(conv3x3_c32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, t.smem_bytes);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
#ifdef NUNET_EXPERIMENT   // experiment builds only (tools/dbg_c32.py); never compiled into the shipped library
    if (const char* dg = getenv("NUNET_C32_DBG")) p.dbg = (p.stat_part == nullptr) ? atoi(dg) : 0;
    if (const char* g = getenv("NUNET_C32_GRID")) {   // fewer CTAs -> is the L2->SM path per-SM or chip-wide bound?
        const int v = atoi(g);
        if (v >= 1 && v < grid && p.stat_part == nullptr) grid = v;
    }
#endif
    conv3x3_c32_kernel<<<grid, kC32Threads, t.smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
