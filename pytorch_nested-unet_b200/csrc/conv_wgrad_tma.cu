// Weight gradient with TMA-fed MN-major operands (sm_100a): the variant used when every source of the conv is a
// pre-activated same-resolution bf16 NHWC tensor (NUNET_SRC_ACT).
//
//   dW[co][ci][dr][dc] = sum over output pixels (n,y,x) of  X[n][y+dr-1][x+dc-1][ci] * gy[n][y][x][co]
//
// Geometry.  Every image is H+1 padded rows (row y = -1 is a zero separator) of PW >= W+2 positions.  A K chunk is either
// `ry` consecutive padded rows of ONE image (large images) or `kimg` whole images (small ones), so that each operand block
// of a chunk is ONE tensor load: a box of (32 channels x PW pixels x rows [x images]), 64B-swizzled; out-of-bounds
// pixels / rows / images are zero-filled by the TMA unit (= the conv's zero padding and the separator rows).  In shared memory a
// block is [row][PW px][32 ch] = one 64-byte line per position, which IS the canonical MN-major SWIZZLE_64B operand
// (K = position, MN = channel, SBO = 8 positions = 512 B).  The X box starts at pixel 0 and the gy box at pixel -1, so tap
// (dr,dc) pairs X position k with gy position k + (2-dr)*PW + (2-dc): a constant start-address offset of the descriptor
// (row-shifted starts of swizzled operands: tools/umma_probe_sw.cu).
//
// Tap stacking.  The MN stride between 32-channel blocks of an operand (the descriptor's LBO) may be ANY multiple of 16 B,
// so "block b" of the gy operand is simply the same buffer one position (64 B) further: LBO = 64 B makes
// N = 96 = 3 kernel columns x 32 output channels out of one copy of gy.  One K step (16 positions) is then
//   regular (G=3): 3 MMAs 128x96x16 (one per kernel row) for a (128 ci) x (32 co) x (9 taps) block = 288 TMEM columns,
//                  ~56 cycles each instead of nine ~45-cycle 128x32x16 MMAs (the tensor pipe is operand-fetch bound);
//   stacked (G=1): cin_pad == 32 only: the kernel ROWS are stacked on the A side the same way (LBO = PW*64 B, the X
//                  buffer carries the two halo rows), so ONE 128x96x16 MMA per K step produces all nine taps
//                  (TMEM lanes dr*32+ci, 96 columns); the level-0 conv2 layers become HBM-bound.
// The chunk's positions are rounded up to a multiple of 16; the tail of the operand that carries no halo is never written
// by TMA and stays zero from the prologue, so the rounding (and the zero padding columns) contribute nothing.
// No producer warps: one thread issues the tensor loads, one the MMAs, eight warps drain TMEM at the end.
#include "conv_common.cuh"
#include "tma.cuh"
#include <cstdlib>

namespace nunet {

struct WgradTmaParams {
    CUtensorMap xmap[NUNET_MAX_SRC];
    CUtensorMap gmap;
    int Hp1;
    int PW, ry, cpi, kimg;   // rows mode (kimg == 0): ry rows per chunk, cpi chunks per image; image mode: kimg images per chunk
    int x_box_rows, g_box_rows;
    int chunks_total, cpc, S, ksteps;
    int stacked;  // 1: cin_pad == 32, kernel rows stacked on the A operand
    int strips, Ws;   // strips > 1: rows wider than a tensor-load box — "image" v is column strip v % strips (Ws pixels) of image v / strips
    FastDiv div_strips;
    int cin_pad, Cout, num_src;
    int srcC[NUNET_MAX_SRC];
    int x_block_bytes, x_stage_bytes, g_stage_bytes;
    float* part;  // [splits][Cout/32][9][cin_pad][32]
};

constexpr int kTThreads = 320;  // warps 0-7: epilogue, warp 8: MMA issuer + TMEM owner, warp 9: TMA issuer

__global__ void __launch_bounds__(kTThreads, 1) conv3x3_wgrad_tma_kernel(const __grid_constant__ WgradTmaParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t x_bytes = (uint32_t)p.x_stage_bytes, g_bytes = (uint32_t)p.g_stage_bytes;
    uint8_t* sX = smem;
    uint8_t* sG = sX + (size_t)p.S * x_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sG + (size_t)p.S * g_bytes);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.S;
    uint64_t* accum = bars + 2 * p.S;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * p.S + 1);

    const int split = blockIdx.x, m_tile = blockIdx.y, n_tile = blockIdx.z;
    const int chunk0 = split * p.cpc;
    const int nchunks = min(p.cpc, p.chunks_total - chunk0);
    const int ci0 = m_tile * 128;
    const int xblocks = min(4, (p.cin_pad - ci0) >> 5);

    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < p.S; ++s) {
                mbar_init(&full[s], 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum, 1);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_ptr);
    }
    // zero tails (see header) and no NaN garbage anywhere the shifted reads may land
    for (uint32_t i = threadIdx.x * 16u; i < (uint32_t)p.S * (x_bytes + g_bytes); i += kTThreads * 16u) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t row_bytes = (uint32_t)p.PW * 64u;

    if (warp == 9) {
        // =============================== TMA issuer
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            for (int si = 0; si < p.num_src; ++si) tma_prefetch_desc(&p.xmap[si]);
            tma_prefetch_desc(&p.gmap);
            int bsrc[4], bch[4];  // (source, channel) of each 32-channel block of this M tile
            {
                int si = 0, c0 = 0;
                for (int b = 0; b < 4; ++b) {
                    const int cglob = ci0 + b * 32;
                    bsrc[b] = 0; bch[b] = 0;
                    if (b < xblocks) {
                        while (cglob >= c0 + p.srcC[si]) { c0 += p.srcC[si]; ++si; }
                        bsrc[b] = si; bch[b] = cglob - c0;
                    }
                }
            }
            // rows mode : regular X rows y0 .. y0+ry-1, gy rows y0-1 .. y0+ry ; stacked X rows y0-1 .. y0+ry, gy rows y0 .. y0+ry-1
            // image mode: the operand that carries the halo is loaded one row down; its first and last buffer rows are never
            //             written (zeros): they only ever meet the other operand's separator rows
            const uint32_t tx = (uint32_t)(p.x_box_rows * xblocks + p.g_box_rows) * row_bytes;
            const uint32_t x_off = (p.kimg && p.stacked) ? row_bytes : 0u, g_off = (p.kimg && !p.stacked) ? row_bytes : 0u;
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                const int s = chunk % p.S;
                const uint32_t ph = (chunk / p.S) & 1;
                const int c = chunk0 + chunk;
                int n, y0;
                if (p.kimg) { n = c * p.kimg; y0 = -1; }
                else { n = c / p.cpi; y0 = (c - n * p.cpi) * p.ry - 1; }
                const int xy = (!p.kimg && p.stacked) ? y0 - 1 : y0, gyy = (!p.kimg && !p.stacked) ? y0 - 1 : y0;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* dx = sX + (size_t)s * x_bytes + x_off;
                if (p.strips > 1) {
                    // Column strips (rows mode only).  The sum over a row splits into strips if X is RESTRICTED to the strip's own
                    // pixels while gy keeps one real halo pixel on either side: X goes row by row through a (32 ch x Ws px) box into
                    // the PW-pitch rows — positions Ws .. PW-1 of every row are never written and stay zero from the prologue — and
                    // the gy box simply starts one pixel to the left of the strip.
                    const int img = (int)fastdiv((uint32_t)n, p.div_strips), st = n - img * p.strips;
                    mbar_arrive_expect_tx(&full[s], (uint32_t)(p.x_box_rows * xblocks) * (uint32_t)p.Ws * 64u + (uint32_t)p.g_box_rows * row_bytes);
                    for (int b = 0; b < xblocks; ++b)
                        for (int r = 0; r < p.x_box_rows; ++r)
                            tma_load_4d(dx + (size_t)b * p.x_block_bytes + (size_t)r * row_bytes, &p.xmap[bsrc[b]], bch[b], st * p.Ws, xy + r, img, &full[s]);
                    tma_load_4d(sG + (size_t)s * g_bytes + g_off, &p.gmap, n_tile * 32, st * p.Ws - 1, gyy, img, &full[s]);
                } else {
                    mbar_arrive_expect_tx(&full[s], tx);
                    for (int b = 0; b < xblocks; ++b) tma_load_4d(dx + (size_t)b * p.x_block_bytes, &p.xmap[bsrc[b]], bch[b], 0, xy, n, &full[s]);
                    tma_load_4d(sG + (size_t)s * g_bytes + g_off, &p.gmap, n_tile * 32, -1, gyy, n, &full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 8) {
        // =============================== MMA issuer
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            const uint32_t idesc = umma_idesc_bf16(128, 96, kMajorMN, kMajorMN);
            const uint32_t hi = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B (8 positions), version 1, SWIZZLE_64B
            const uint32_t a_lbo = p.stacked ? (row_bytes >> 4) : ((uint32_t)p.x_block_bytes >> 4);
            const uint32_t pw4 = (uint32_t)p.PW * 4u;                    // one padded row in 16-byte address units
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                const int s = chunk % p.S;
                const uint32_t ph = (chunk / p.S) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                uint32_t a_lo = ((smem_u32(sX + (size_t)s * x_bytes) >> 4) & 0x3FFFu) | (a_lbo << 16);
                uint32_t b_lo = ((smem_u32(sG + (size_t)s * g_bytes) >> 4) & 0x3FFFu) | (4u << 16);   // LBO = 64 B: block b = +b positions = kernel column 2-b
                if (p.stacked) {
                    for (int ks = 0; ks < p.ksteps; ++ks) {   // 16 positions = 1024 B = 64 address units per K step
                        umma_bf16_lohi(tmem_base, a_lo, hi, b_lo, hi, idesc, (uint32_t)((chunk | ks) != 0));
                        a_lo += 64u;
                        b_lo += 64u;
                    }
                } else {
                    for (int ks = 0; ks < p.ksteps; ++ks) {
                        const uint32_t accf = (uint32_t)((chunk | ks) != 0);
                        umma_bf16_lohi(tmem_base + 0u, a_lo, hi, b_lo + 2u * pw4, hi, idesc, accf);
                        umma_bf16_lohi(tmem_base + 96u, a_lo, hi, b_lo + pw4, hi, idesc, accf);
                        umma_bf16_lohi(tmem_base + 192u, a_lo, hi, b_lo, hi, idesc, accf);
                        a_lo += 64u;
                        b_lo += 64u;
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accum);
        }
        __syncwarp();
    } else {
        // =============================== epilogue: TMEM -> fp32 partial block [tap][ci][32]
        mbar_wait(accum, 0);
        tc_fence_after();
        const int lq = warp & 3, half = warp >> 2;
        float* dst_base = p.part + ((size_t)(split * (p.Cout >> 5) + n_tile) * 9) * p.cin_pad * 32;
        if (p.stacked) {
            // lanes dr*32 + ci (dr = lq < 3), columns b*32 + co with kernel column dc = 2 - b
            if (lq < 3) {
                for (int b = half; b < 3; b += 2) {
                    uint32_t v[32];
                    tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * 32), v);
                    tmem_ld_wait();
                    const int tap = lq * 3 + (2 - b);
                    uint4* d = reinterpret_cast<uint4*>(dst_base + ((size_t)tap * p.cin_pad + lane) * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        } else {
            const int ci = ci0 + lq * 32 + lane;
            for (int j = half; j < 9; j += 2) {   // column block j = dr*3 + b, kernel column dc = 2 - b
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(j * 32), v);
                tmem_ld_wait();
                const int tap = (j / 3) * 3 + (2 - j % 3);
                if (ci < p.cin_pad) {
                    uint4* d = reinterpret_cast<uint4*>(dst_base + ((size_t)tap * p.cin_pad + ci) * 32);
#pragma unroll
                    for (int q = 0; q < 8; ++q) d[q] = make_uint4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem_base);
}

// dW[co][ci][tap] (fp32 OIHW) = sum over splits, fixed order.  part: [splits][Cout/32][9][cin_pad][32]
__global__ void wgrad_reduce_nt_kernel(const float* __restrict__ part, int splits, int Cout, int cin_pad, int Cin, float* __restrict__ dw) {
    const size_t block_elems = (size_t)9 * cin_pad * Cout;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < block_elems; i += (size_t)gridDim.x * blockDim.x) {
        size_t rr = i;
        const int col = (int)(rr & 31); rr >>= 5;
        const int ci = (int)(rr % cin_pad); rr /= cin_pad;
        const int tap = (int)(rr % 9); rr /= 9;
        const int nt = (int)rr;
        if (ci >= Cin) continue;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(size_t)k * block_elems + i];
        dw[((size_t)(nt * 32 + col) * Cin + ci) * 9 + tap] = s;
    }
}

}  // namespace nunet

using namespace nunet;

struct TmaTiling {
    int PW, ry, cpi, kimg, chunks, S, ksteps, stacked, x_box_rows, g_box_rows, x_block_bytes, x_stage_bytes, g_stage_bytes, smem_bytes;
};

// Picks the chunk shape (rows of one image, or whole images) that minimises the number of 16-position K steps of the whole
// layer (+1 per chunk for the pipeline hand-off) among those that leave at least two pipeline stages.
// Rows wider than a tensor-load box: N*strips column strips of Ws (even) pixels, as in nunet_conv3x3_plan.
static int wgrad_strips(int W, int* Ws) {
    static const bool no_strips = getenv("NUNET_NO_STRIPS") != nullptr;
    int strips = (W + 2 > 256 && !no_strips) ? (W + 253) / 254 : 1;
    int w = (W + strips - 1) / strips;
    if (strips > 1) w += w & 1;
    *Ws = w;
    return strips;
}

static bool tma_tiling(int N, int H, int W, int cin_pad, TmaTiling* t) {
    int Ws;
    const int strips = wgrad_strips(W, &Ws);
    N *= strips;
    W = Ws;
    const int PW = (W + 2 + 1) & ~1;          // even: every padded row starts 128-byte aligned
    const int Hp1 = H + 1;
    if (PW > 256) return false;               // TMA box dimension limit
    const int stacked = (cin_pad == 32) ? 1 : 0;
    const int budget = 225 * 1024;
    double best = 1e30;
    const int halo = 2 * PW + 3;
    for (int pass = 0; pass < 2; ++pass) {
        const int lim = pass == 0 ? (Hp1 < 254 ? Hp1 : 254) : (N < 256 ? N : 256);
        if (pass == 1 && (Hp1 > 256 || strips > 1)) break;   // whole-image chunks need consecutive images in the batch dimension
        for (int q = 1; q <= lim; ++q) {
            const int rows = pass == 0 ? q : q * Hp1;
            const int ksteps = (rows * PW + 15) / 16, kpos = ksteps * 16;
            const int xb = ((kpos + (stacked ? halo : 0)) * 64 + 1023) & ~1023;
            const int gb = ((kpos + (stacked ? 3 : halo)) * 64 + 1023) & ~1023;
            const int xs = stacked ? xb : 4 * xb;
            if (2 * (xs + gb) + 64 > budget) break;
            int S = 4;
            while (S * (xs + gb) + (2 * S + 1) * 8 + 16 > budget) --S;
            const long long chunks = pass == 0 ? (long long)N * ((Hp1 + q - 1) / q) : (N + q - 1) / q;
            double cost = (double)chunks * (ksteps + 1);
            if (S == 2) cost *= 1.05;
            if (cost < best) {
                best = cost;
                t->PW = PW; t->ry = pass == 0 ? q : 0; t->cpi = pass == 0 ? (Hp1 + q - 1) / q : 0; t->kimg = pass == 0 ? 0 : q;
                t->chunks = (int)chunks; t->S = S; t->ksteps = ksteps; t->stacked = stacked;
                const int halo_rows = pass == 0 ? 2 : 0;
                t->x_box_rows = rows + (stacked ? halo_rows : 0); t->g_box_rows = rows + (stacked ? 0 : halo_rows);
                t->x_block_bytes = xb; t->x_stage_bytes = xs; t->g_stage_bytes = gb;
                t->smem_bytes = S * (xs + gb) + (2 * S + 1) * 8 + 16;
            }
        }
    }
    return best < 1e30;
}

// Tiling of the TMA variant; returns 0 and fills the tma_* fields of `pl`, or nonzero if it does not apply.
int nunet_wgrad_tma_plan(int N, int H, int W, int cin_pad, int Cout, nunet_wgrad_plan_t* pl) {
    pl->tma = 0;
    static const bool off = getenv("NUNET_NO_WGRAD_TMA") != nullptr;
    TmaTiling t;
    if (off || cin_pad % 32 != 0 || Cout % 32 != 0 || !tma_tiling(N, H, W, cin_pad, &t)) return 1;
    const int chunks_total = t.chunks;
    const int m_tiles = (cin_pad + 127) / 128, n_tiles = Cout / 32;
    // one wave of CTAs (1 CTA per SM: 288 of 512 TMEM columns, > 113 KB shared memory), at least 2 chunks each
    int splits = 148 / (m_tiles * n_tiles);
    if (splits < 1) splits = 1;
    int max_splits = chunks_total / 2;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    const int cpc = (chunks_total + splits - 1) / splits;
    splits = (chunks_total + cpc - 1) / cpc;
    int S = t.S;
    if (S > cpc) S = cpc < 1 ? 1 : cpc;
    pl->tma = 1; pl->tma_PW = t.PW; pl->tma_r = t.kimg ? -t.kimg : t.ry; pl->tma_S = S; pl->tma_G = t.stacked ? 1 : 3; pl->tma_Nt = 32;
    pl->tma_splits = splits; pl->tma_cpc = cpc; pl->tma_chunks = chunks_total;
    pl->tma_smem_bytes = S * (t.x_stage_bytes + t.g_stage_bytes) + (2 * S + 1) * 8 + 16;
    return 0;
}

int nunet_wgrad_tma_launch(const nunet_wgrad_desc_t* d, const nunet_wgrad_plan_t& pl, int cin_real, cudaStream_t stream) {
    TmaTiling t;
    if (!tma_tiling(d->N, d->H, d->W, pl.cin_pad, &t)) return -1;
    WgradTmaParams p{};
    p.Hp1 = d->H + 1;
    p.PW = t.PW; p.ry = t.ry; p.cpi = t.cpi; p.kimg = t.kimg; p.x_box_rows = t.x_box_rows; p.g_box_rows = t.g_box_rows;
    p.chunks_total = pl.tma_chunks; p.cpc = pl.tma_cpc; p.S = pl.tma_S; p.ksteps = t.ksteps;
    p.stacked = t.stacked;
    p.strips = wgrad_strips(d->W, &p.Ws);
    p.div_strips = make_fastdiv((uint32_t)(p.strips > 1 ? p.strips : 2));
    p.cin_pad = pl.cin_pad; p.Cout = d->Cout; p.num_src = d->num_src;
    p.x_block_bytes = t.x_block_bytes; p.x_stage_bytes = t.x_stage_bytes; p.g_stage_bytes = t.g_stage_bytes;
    p.part = d->part;
    const int xbh = t.kimg ? p.Hp1 : t.x_box_rows, xbn = t.kimg ? t.kimg : 1;
    const int gbh = t.kimg ? p.Hp1 : t.g_box_rows, gbn = t.kimg ? t.kimg : 1;
    for (int i = 0; i < d->num_src; ++i) {
        p.srcC[i] = d->src[i].C;
        if (p.strips > 1 ? encode_nhwc_map(&p.xmap[i], d->src[i].data, d->N, d->H, d->W, d->src[i].C, 32, p.Ws, 64, 1, 1) != 0
                         : encode_nhwc_map(&p.xmap[i], d->src[i].data, d->N, d->H, d->W, d->src[i].C, 32, p.PW, 64, xbh, xbn) != 0) return -1;
    }
    if (encode_nhwc_map(&p.gmap, d->gy, d->N, d->H, d->W, d->Cout, 32, p.PW, 64, gbh, gbn) != 0) return -1;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.tma_smem_bytes);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    const int m_tiles = (pl.cin_pad + 127) / 128;
    conv3x3_wgrad_tma_kernel<<<dim3(pl.tma_splits, m_tiles, d->Cout / 32), kTThreads, pl.tma_smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    const size_t block_elems = (size_t)9 * pl.cin_pad * d->Cout;
    int blocks = (int)((block_elems + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_nt_kernel<<<blocks, 256, 0, stream>>>(d->part, pl.tma_splits, d->Cout, pl.cin_pad, cin_real, d->dw);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
