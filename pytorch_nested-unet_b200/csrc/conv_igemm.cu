// Fused 3x3 convolution (forward and dgrad) as an implicit GEMM on tcgen05 tensor cores, sm_100a.
//
// Geometry ("flattened padded" implicit GEMM): the batch is viewed as one 1-D sequence of padded positions
//   q = R*Wp + c,  Wp = W+2,  R in [0, N*(H+1)]  (row n*(H+1) is the shared zero row between images n-1 and n),
// so the nine taps of a 3x3 stencil are nine constant offsets (dr*Wp + dc) of q.  A CTA owns 128*T consecutive
// positions (T accumulator tiles of M=128 in TMEM).  Its producer warps stage P = 128*T + 2*Wp + 2 positions of the
// (virtual) concatenated input ONCE per K chunk into shared memory — applying on the way the producer's BatchNorm
// scale/shift + ReLU, the 2x2 max-pool or the align_corners bilinear 2x upsample of the skip pathway, and the channel
// concat (the K loop simply walks the sources) — in the no-swizzle K-major core-matrix layout
//   A[cgroup][position][8 channels]   (16 B per (position, cgroup); core matrix = 8 consecutive positions).
// Tap (dr,dc) of tile t is then the SAME buffer read through a descriptor whose start address is advanced by
// (t*128 + (dr+1)*Wp + (dc+1)) * 16 bytes (verified on B200 by tools/umma_probe.cu), so nothing is ever im2col'ed
// and the concatenated / pooled / upsampled tensors never exist in HBM.  Weights arrive pre-packed as the exact
// shared-memory image of a (N tile, K chunk) block and are fetched with one cp.async.bulk per chunk.
//
// Warp roles (320 threads): warps 0-7 producers then epilogue, warp 8 MMA issuer (one lane) + TMEM owner,
// warp 9 weight fetcher (one lane).  Pipelines: full[s]/empty[s] per smem stage, accum for TMEM -> epilogue.
// Epilogue: tcgen05.ld -> bf16 NHWC store (+ per-channel sum / sum-of-squares of the fp32 accumulators for
// train-mode BatchNorm, reduced by a transposing warp butterfly, one partial row per CTA).
#include "conv_common.cuh"
#include "loaders.cuh"
#include "tma.cuh"
#include <cstdlib>

namespace nunet {

// ------------------------------------------------------------------------------------------------ kernel
constexpr int kMmaWarp = 16;                      // warp 16   : tcgen05.mma issuer (one lane) + TMEM owner
constexpr int kFetchWarp = 17;                    // warp 17   : weight fetcher (cp.async.bulk, one lane; TMA mode: the activation tensor loads too)
constexpr int kThreads = 18 * 32;
// Warps 0-15.  Loader modes: warps 0-7 stage the transformed input halo, warps 8-15 drain TMEM (warp % 4 = TMEM lane quadrant, two
// warps per quadrant).  TMA mode (MODE_ACT_TMA): nothing is staged by threads, so ALL SIXTEEN warps drain TMEM (four per quadrant) —
// the epilogue of the level-0 layers is a latency chain (tcgen05.ld -> convert -> store) and eight warps could not hide it
// (profiles/r02_ncu_stalls.md: 23 % of the dgrad 32->192 samples sat on the store waiting for its shared-memory transposition).
constexpr int kProducerWarps = 8;
#ifndef NUNET_DBG_SKIP
#define NUNET_DBG_SKIP 0      // experiment builds only (tools/ab_epilogue.sh): 1 no epilogue stores, 2 no MMAs, 4 no activation tensor loads
#endif
constexpr int kEpiSlot = 32 * 80;                 // bytes of transpose scratch per epilogue warp
constexpr uint32_t kStatBytes = 8192;             // per-thread-owned statistic slots

// Shared-memory carve-up, identical on host (sizing) and device.
struct SmemLayout {
    uint32_t a_bytes, b_bytes, off_b, off_tab, off_coef, off_mask, off_stat, off_seg, off_epi, off_bar, total;
};
__host__ __device__ inline SmemLayout conv_smem_layout(bool tma, int T, int NT, int CK, int S, int P, int cin_pad, int a_stage_bytes, bool res_b, int epi_warps) {
    SmemLayout L;
    L.a_bytes = tma ? (uint32_t)a_stage_bytes : (uint32_t)(CK / 8) * P * 16u;
    L.b_bytes = 9u * CK * NT * 2u;
    uint32_t o = (uint32_t)S * L.a_bytes;
    L.off_b = o;       o += (res_b ? 1u : (uint32_t)S) * L.b_bytes;
    L.off_tab = o;     o += tma ? 0u : 2u * P * 8u;                    // position tables (loader modes only)
    L.off_coef = o;    o += tma ? 0u : (uint32_t)cin_pad * 12u;        // per-input-channel loader coefficients
    L.off_mask = o;    o += 2u * NT * 4u;                              // ReLU-mask coefficients of the fused-router epilogue
    L.off_seg = o;     o += 24u * 16u;                                 // output segment of each 32-channel slab {pointer, pitch}
    // statistic slots [4 * TG rows][2][NT] floats, TG = tile groups per lane quadrant (<= 8 KB), then the transpose scratch
    L.off_stat = o;    o += kStatBytes;
    L.off_epi = o;     o += (uint32_t)epi_warps * kEpiSlot;
    L.off_bar = o;     o += (2u * S + 6u) * 8u + 16u;
    L.total = o + 128u;
    return L;
}

struct OutSeg {
    __nv_bfloat16* base;   // already offset to the slab's first channel inside its segment
    int pitch, pad_;
};

// Optional wait/work cycle accounting per warp role (build with -DNUNET_PROFILE_WAITS; tools/prof_waits.py reads it).
#ifdef NUNET_PROFILE_WAITS
__device__ unsigned long long g_wait_prof[16];
#define PROF_DECL long long prof_t = clock64(); long long prof_acc[4] = {0, 0, 0, 0};
#define PROF_MARK(slot) { long long now_ = clock64(); prof_acc[slot] += now_ - prof_t; prof_t = now_; }
#define PROF_FLUSH(base) { for (int i_ = 0; i_ < 4; ++i_) atomicAdd(&g_wait_prof[(base) + i_], (unsigned long long)prof_acc[i_]); }
#define PROF_ENTRY const long long prof_entry = clock64();
#define PROF_PROLOGUE(slot) prof_acc[slot] += prof_t - prof_entry;                 /* kernel entry -> role loop start */
#define PROF_MAX_SINCE(idx) atomicMax(&g_wait_prof[idx], (unsigned long long)(clock64() - prof_t));   /* slowest CTA: the BatchNorm tail */
#else
#define PROF_ENTRY
#define PROF_PROLOGUE(slot)
#define PROF_MAX_SINCE(idx)
#define PROF_DECL
#define PROF_MARK(slot) {}
#define PROF_FLUSH(base) {}
#endif

// Persistent: CTA (x, n_tile) processes tile groups x, x+gridDim.x, ... of 128*T positions each.  Accumulators (and, in the loader
// modes, the position table) are double-buffered so the operand pipeline and the tensor core work on group i+1 while the epilogue
// drains group i.
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_igemm_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    PROF_ENTRY
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cgs = p.CK >> 3;
    constexpr bool kTma = (MODE == MODE_ACT_TMA);
    const int kEpi = kTma ? p.epi_warps : 8;       // epilogue warps (TMA mode: 16 where their transpose scratch fits, else 8)
    const int kParts = kEpi >> 2;                  // warps per TMEM lane quadrant
    // A single K chunk (Cin <= 32: the level-0 second convs and every level-0 dgrad): the weight block is the same for every tile
    // group, so it is fetched ONCE per CTA and stays resident; only the activation stages cycle.
    const bool res_b = p.res_b != 0;
    const SmemLayout L = conv_smem_layout(kTma, p.T, p.NT, p.CK, p.S, p.P, p.cin_pad, p.a_stage_bytes, res_b, kEpi);
    // TMA mode: an A stage is (P + Wp) whole padded rows of 32 channels (64-byte rows, 64B hardware swizzle) written by
    // tensor loads; otherwise the no-swizzle [channel group][position][8] layout written by the producer warps.
    const uint32_t a_bytes = L.a_bytes, b_bytes = L.b_bytes;
    const uint32_t b_stride = res_b ? 0u : b_bytes;
    uint8_t* sA = smem;
    uint8_t* sB = smem + L.off_b;
    PixInfo* pixtab = reinterpret_cast<PixInfo*>(smem + L.off_tab);  // [2][P]
    float* s_scale = reinterpret_cast<float*>(smem + L.off_coef);
    float* s_shift = s_scale + p.cin_pad;
    float* s_k2 = s_shift + p.cin_pad;
    float* s_mask = reinterpret_cast<float*>(smem + L.off_mask);     // [2][NT]: scale, shift of the masking BatchNorm
    float* s_stat = reinterpret_cast<float*>(smem + L.off_stat);
    OutSeg* s_seg = reinterpret_cast<OutSeg*>(smem + L.off_seg);
    uint8_t* s_epi = smem + L.off_epi;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    uint64_t* full = bars;                 // [S]  producers + fetcher -> MMA
    uint64_t* empty = bars + p.S;          // [S]  MMA -> producers + fetcher
    uint64_t* acc_full = bars + 2 * p.S;   // [2]  MMA -> epilogue
    uint64_t* acc_empty = acc_full + 2;    // [2]  epilogue -> MMA (+ producers: accumulator AND table buffer are free)
    uint64_t* tab_full = acc_empty + 2;    // [2]  producers -> epilogue (position table of the group is written)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tab_full + 2);

    const int n_tile = blockIdx.y;
    const int group_positions = 128 * p.T;

    // ---- one-time setup
    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int s = 0; s < p.S; ++s) {
                mbar_init(&full[s], kTma ? 1 : kProducerWarps + 1);
                mbar_init(&empty[s], 1);
            }
            for (int b = 0; b < 2; ++b) {
                mbar_init(&acc_full[b], 1);
                mbar_init(&acc_empty[b], kEpi);
                mbar_init(&tab_full[b], kProducerWarps);
            }
            mbar_init_fence();
        }
        __syncwarp();
        switch (p.tmem_cols) {
            case 32: tmem_alloc<32>(tmem_ptr); break;
            case 64: tmem_alloc<64>(tmem_ptr); break;
            case 128: tmem_alloc<128>(tmem_ptr); break;
            case 256: tmem_alloc<256>(tmem_ptr); break;
            default: tmem_alloc<512>(tmem_ptr); break;
        }
    }
    {
        if constexpr (!kTma) {
            int cbase = 0;
            for (int si = 0; si < p.num_src; ++si) {
                const ConvSrc& s = p.src[si];
                const int cpad = (s.kind == NUNET_SRC_INPUT) ? 16 : s.C;
                for (int c = threadIdx.x; c < cpad; c += kThreads) {
                    const bool has = (s.kind != NUNET_SRC_INPUT) && s.scale != nullptr;
                    s_scale[cbase + c] = has ? s.scale[c] : 1.f;
                    s_shift[cbase + c] = has ? s.shift[c] : 0.f;
                    s_k2[cbase + c] = (s.kind == NUNET_SRC_GRAD) ? s.k2[c] : 0.f;
                }
                cbase += cpad;
            }
        }
        for (int i = threadIdx.x; i < (int)(kStatBytes / 4); i += kThreads) s_stat[i] = 0.f;
        if (p.mask_y)
            for (int c = threadIdx.x; c < p.NT; c += kThreads) {
                s_mask[c] = p.mask_scale[n_tile * p.NT + c];
                s_mask[p.NT + c] = p.mask_shift[n_tile * p.NT + c];
            }
        // where each 32-channel slab of this N tile goes: the dense output, or the per-source segment that owns those channels
        for (int sl = threadIdx.x; sl < (p.NT >> 5); sl += kThreads) {
            int gc = n_tile * p.NT + sl * 32;
            OutSeg o;
            o.base = reinterpret_cast<__nv_bfloat16*>(p.out) + gc;
            o.pitch = p.Cout;
            o.pad_ = 0;
            for (int k = 0; k < p.num_out_seg; ++k) {
                if (gc < p.out_seg_c[k]) { o.base = reinterpret_cast<__nv_bfloat16*>(p.out_seg[k]) + gc; o.pitch = p.out_seg_c[k]; break; }
                gc -= p.out_seg_c[k];
            }
            s_seg[sl] = o;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const bool is_epi = kTma ? (warp < kEpi) : (warp >= kProducerWarps && warp < 16);

    if (!kTma && warp < kProducerWarps) {
        // =============================== producers (loader modes)
        const int nblk = (p.P + 7) >> 3;
        const int pl = lane & 7, r = lane >> 3;
        // CK=32: the four lane groups of a warp take the four channel groups of one 8-position block;
        // CK=16: two channel groups, the lane-group pairs take two different blocks
        const int cg = r % cgs;
        const int first_blk = (cgs == 4) ? warp : (warp * 2 + (r >> 1));
        const int blk_stride = (cgs == 4) ? kProducerWarps : 2 * kProducerWarps;
        int cc = 0;  // running chunk counter across tile groups
        int it = 0;
        PROF_DECL
        for (int tg = blockIdx.x; tg < p.ntg; tg += gridDim.x, ++it) {
            const int b = it & 1;
            const uint32_t k = (uint32_t)(it >> 1);
            PixInfo* tab = pixtab + (size_t)b * p.P;
            const int qbase = tg * group_positions - p.Wp - 1;
            mbar_wait(&acc_empty[b], (k & 1) ^ 1);  // the epilogue of group it-2 no longer reads table b
            PROF_MARK(0)
            for (int i = threadIdx.x; i < p.P; i += kProducerWarps * 32) tab[i] = decode_position(qbase + i, p.N, p.H, p.W, p.Wp, p.Hp1, p.Q);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (lane == 0) mbar_arrive(&tab_full[b]);
            PROF_MARK(1)
            int si = 0, src_c0 = 0;  // current source and its first concatenated channel
            for (int chunk = 0; chunk < p.nchunks; ++chunk, ++cc) {
                const int s = cc % p.S;
                const uint32_t ph = (uint32_t)(cc / p.S) & 1;
                const int cglob = chunk * p.CK;  // first concatenated channel of this chunk
                while (true) {
                    const int cpad = (p.src[si].kind == NUNET_SRC_INPUT) ? 16 : p.src[si].C;
                    if (cglob < src_c0 + cpad) break;
                    src_c0 += cpad;
                    ++si;
                }
                const int ch = cglob - src_c0 + cg * 8;  // channel inside the source
                mbar_wait(&empty[s], ph ^ 1);
                PROF_MARK(2)
                uint8_t* dstA = sA + (size_t)s * a_bytes + (size_t)cg * p.P * 16;
                stage_blocks_mode<MODE>(p.src[si], p.H, p.W, tab, 0, p.P, first_blk, nblk, blk_stride, pl, ch, s_scale + cglob + cg * 8,
                                      s_shift + cglob + cg * 8, s_k2 + cglob + cg * 8, dstA);
                if constexpr (MODE == MODE_ACT || MODE == MODE_ACT_UP) {
                    cp_async_commit();
                    cp_async_wait<0>();
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&full[s]);
                PROF_MARK(3)
            }
        }
        if (threadIdx.x == 0) PROF_FLUSH(0)
    } else if (is_epi) {
        // =============================== epilogue: kEpi warps, kParts per TMEM lane quadrant
        const int eidx = kTma ? warp : warp - kProducerWarps;
        const int lq = eidx & 3, part = eidx >> 2;      // lq == warp % 4 in both numberings
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        const bool want_stats = p.stat_part != nullptr;
        const bool masked = p.mask_y != nullptr;         // fused block-internal gradient router (dgrad of a block's second conv)
        const __nv_bfloat16* mask_y = reinterpret_cast<const __nv_bfloat16*>(p.mask_y);
        auto epi_pix = [&](int q) {   // output pixel of flattened position q (-1: padding / halo)
            return p.strips > 1 ? decode_pix_strips(q, p.N, p.H, p.W, p.Wp, p.Hp1, p.Q, p.div_wp, p.div_hp1, p.strips, p.Wreal, p.div_strips)
                                : decode_pix_fast(q, p.N, p.H, p.W, p.Wp, p.Hp1, p.Q, p.div_wp, p.div_hp1);
        };
        // statistic paths: 32-column slabs; a warp owns slab group sg (of G) and tile group tsel (of TG)
        const int nslab32 = p.NT >> 5;
        int G = kParts;
        while (G > nslab32) G >>= 1;                     // a divisor of kParts (a power of two)
        const int TG = kParts / G;
        const int sg = part % G, tsel = part / G;
        float* my_stat = s_stat + (size_t)((lq * TG + tsel) * 2) * p.NT;   // slots written by this warp only (deterministic)
        const int chunk = lane & 3, rsub = lane >> 2;
        // the tile-group loop is written twice (EPI_GROUP_BEGIN / _END): the statistics path keeps its sums in registers across all
        // groups, which must not cost the plain path registers
#ifdef NUNET_POLL1
#define EPI_GROUP_WAIT if (lane == 0) { if constexpr (!kTma) mbar_wait(&tab_full[b], k & 1); mbar_wait(&acc_full[b], k & 1); } __syncwarp();
#else
#define EPI_GROUP_WAIT if constexpr (!kTma) mbar_wait(&tab_full[b], k & 1); mbar_wait(&acc_full[b], k & 1);
#endif
#define EPI_GROUP_BEGIN                                                                                                   \
        int it = 0;                                                                                                       \
        PROF_DECL                                                                                                         \
        for (int tg = blockIdx.x; tg < p.ntg; tg += gridDim.x, ++it) {                                                    \
            const int b = it & 1;                                                                                         \
            const uint32_t k = (uint32_t)(it >> 1);                                                                       \
            const PixInfo* tab = pixtab + (size_t)b * p.P;                                                                \
            const int q0 = tg * group_positions + lq * 32 + lane; /* this lane's output position in tile 0 of the group */ \
            EPI_GROUP_WAIT                                                                                                \
            PROF_MARK(0)                                                                                                  \
            tc_fence_after();                                                                                             \
            const uint32_t acc = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(b * p.T * p.NT);                    \
            (void)tab;
#define EPI_GROUP_END                                                                                                     \
            tc_fence_before();                                                                                            \
            __syncwarp();                                                                                                 \
            if (lane == 0) mbar_arrive(&acc_empty[b]);                                                                    \
            PROF_MARK(1)                                                                                                  \
        }                                                                                                                 \
        if (eidx == 0 && lane == 0) PROF_FLUSH(4)
        if (!want_stats) {
            EPI_GROUP_BEGIN
                // plain store path (dgrad, eval): 32 columns per TMEM load = 64 bytes per pixel row.  A lane owns a ROW, so a direct
                // store would touch 32 different 128-byte lines per instruction; the 32x64 B tile is transposed through a padded
                // per-warp smem scratch so that four lanes write one row's 64 contiguous bytes (8 lines per store instruction).
                uint8_t* scratch = s_epi + (size_t)eidx * kEpiSlot;
                for (int wi = part; wi < p.T * nslab32; wi += kParts) {
                    const int t = wi / nslab32, cs = wi - t * nslab32;
                    int pix;
                    if constexpr (kTma) pix = epi_pix(q0 + t * 128);
                    else pix = tab[t * 128 + lq * 32 + lane + p.Wp + 1].pix;
                    uint32_t v[32];
                    tmem_ld32(acc + (uint32_t)(t * p.NT + cs * 32), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                        o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                        o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                        o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                        *reinterpret_cast<uint4*>(scratch + lane * 80 + j * 16) = o;
                    }
                    __syncwarp();
                    const OutSeg seg = s_seg[cs];
                    __nv_bfloat16* obase = seg.base + (lane & 3) * 8;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + (lane >> 2);
                        const int rpix = __shfl_sync(0xffffffffu, pix, row);
                        const uint4 o = *reinterpret_cast<const uint4*>(scratch + row * 80 + (lane & 3) * 16);
                        if (rpix >= 0 && !(NUNET_DBG_SKIP & 1)) *reinterpret_cast<uint4*>(obase + (size_t)rpix * seg.pitch) = o;
                    }
                    __syncwarp();
                }
            EPI_GROUP_END
        } else {
            // a lane owns 8 channels (chunk) of every 8th row (rsub) of a 32x32 slab; its sums stay in registers across ALL tile groups
            // when the warp owns one slab (NT <= 128 with sixteen epilogue warps), else they are flushed after every group
            // (a second register set would spill: 113 registers per thread at 576 threads)
            constexpr int kAccSets = 1;
            const bool hoist = nslab32 <= kAccSets * G;
            float s1[kAccSets][8], s2[kAccSets][8];
#pragma unroll
            for (int sl = 0; sl < kAccSets; ++sl)
#pragma unroll
                for (int e = 0; e < 8; ++e) { s1[sl][e] = 0.f; s2[sl][e] = 0.f; }
            auto flush_stats = [&](int cs0) {
                // fold the eight lanes that hold the same channel chunk (fixed order); one writer per slot
#pragma unroll
                for (int sl = 0; sl < kAccSets; ++sl) {
                    const int cs = cs0 + sl * G;
                    if (cs >= nslab32) break;
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
#pragma unroll
                        for (int off = 4; off <= 16; off <<= 1) {
                            s1[sl][e] += __shfl_xor_sync(0xffffffffu, s1[sl][e], off);
                            s2[sl][e] += __shfl_xor_sync(0xffffffffu, s2[sl][e], off);
                        }
                    }
                    if (rsub == 0) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            my_stat[cs * 32 + chunk * 8 + e] += s1[sl][e];
                            my_stat[p.NT + cs * 32 + chunk * 8 + e] += s2[sl][e];
                        }
                    }
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s1[sl][e] = 0.f; s2[sl][e] = 0.f; }
                }
            };
            EPI_GROUP_BEGIN
                // statistics epilogues (training forward / fused block-internal router): the same transposed 64-byte-per-row store
                // as the plain path (round 1 stored 32 bytes per lane: 32 cache lines per instruction).  The per-channel sums are
                // taken in the transposed domain, from the bf16 values that are actually stored: a lane owns 8 channels of every 8th
                // row; the eight lanes of a channel chunk are folded by three shuffles per slab.
                uint8_t* scratch = s_epi + (size_t)eidx * kEpiSlot;
                for (int cs0 = sg; cs0 < nslab32; cs0 += kAccSets * G) {
#pragma unroll
                    for (int sl = 0; sl < kAccSets; ++sl) {
                        const int cs = cs0 + sl * G;
                        if (cs >= nslab32) break;
                        const int gc = n_tile * p.NT + cs * 32 + chunk * 8;
                        for (int t = tsel; t < p.T; t += TG) {
                            int pix;
                            if constexpr (kTma) pix = epi_pix(q0 + t * 128);
                            else pix = tab[t * 128 + lq * 32 + lane + p.Wp + 1].pix;
                            uint32_t v[32];
                            tmem_ld32(acc + (uint32_t)(t * p.NT + cs * 32), v);
                            int rpix[4];
                            uint4 ym[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                rpix[i] = __shfl_sync(0xffffffffu, pix, i * 8 + rsub);
                                if (masked && rpix[i] >= 0) ym[i] = __ldg(reinterpret_cast<const uint4*>(mask_y + (size_t)rpix[i] * p.Cout + gc));
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 o;
                                o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                                o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                                o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                                o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                                *reinterpret_cast<uint4*>(scratch + lane * 80 + j * 16) = o;
                            }
                            __syncwarp();
                            float msc[8], msh[8];
                            if (masked) {
                                *reinterpret_cast<float4*>(&msc[0]) = *reinterpret_cast<const float4*>(&s_mask[cs * 32 + chunk * 8]);
                                *reinterpret_cast<float4*>(&msc[4]) = *reinterpret_cast<const float4*>(&s_mask[cs * 32 + chunk * 8 + 4]);
                                *reinterpret_cast<float4*>(&msh[0]) = *reinterpret_cast<const float4*>(&s_mask[p.NT + cs * 32 + chunk * 8]);
                                *reinterpret_cast<float4*>(&msh[4]) = *reinterpret_cast<const float4*>(&s_mask[p.NT + cs * 32 + chunk * 8 + 4]);
                            }
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                if (rpix[i] < 0) continue;
                                uint4 o = *reinterpret_cast<const uint4*>(scratch + (i * 8 + rsub) * 80 + chunk * 16);
                                float f[8];
                                unpack8(o, f);
                                if (masked) {
                                    // g_z = dX * [scale*y + shift > 0]; sums: S1 = sum g_z, S2' = sum g_z*y (combined with mean / invstd at the end)
                                    float yv[8];
                                    unpack8(ym[i], yv);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float z = fmaf(yv[e], msc[e], msh[e]);
                                        f[e] = (z > 0.f) ? f[e] : 0.f;
                                        s1[sl][e] += f[e];
                                        s2[sl][e] = fmaf(f[e], yv[e], s2[sl][e]);
                                    }
                                    o = pack8(f);
                                } else {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        s1[sl][e] += f[e];
                                        s2[sl][e] = fmaf(f[e], f[e], s2[sl][e]);
                                    }
                                }
                                *reinterpret_cast<uint4*>(out + (size_t)rpix[i] * p.Cout + gc) = o;
                            }
                            __syncwarp();
                        }
                    }
                    if (!hoist) flush_stats(cs0);
                }
            EPI_GROUP_END
            if (hoist) flush_stats(sg);
            asm volatile("bar.sync 2, %0;" ::"r"(kEpi * 32) : "memory");  // the epilogue warps only
            const int tid_e = eidx * 32 + lane;
            float* dst = p.stat_part + (size_t)blockIdx.x * 2 * p.Cout + n_tile * p.NT;
            for (int c = tid_e; c < p.NT; c += kEpi * 32) {
                float a = 0.f, bb = 0.f;
                for (int q = 0; q < 4 * TG; ++q) {       // fixed order over the rows of slots
                    a += s_stat[(q * 2 + 0) * p.NT + c];
                    bb += s_stat[(q * 2 + 1) * p.NT + c];
                }
                if (masked) bb = p.mask_invstd[n_tile * p.NT + c] * (bb - p.mask_mean[n_tile * p.NT + c] * a);   // sum g_z*yhat
                dst[c] = a;
                dst[p.Cout + c] = bb;
            }
            // BatchNorm finalisation by the last CTA of the grid (all operand stages are idle by now: sA is free scratch)
            if (p.tail.counter && tid_e < 256) bn_tail_run(p.tail, p.stat_part, (int)gridDim.x, p.Cout, gridDim.x * gridDim.y, tid_e, 3, sA);
            if (tid_e == 0) { PROF_MAX_SINCE(6) }
        }
    } else if (warp == kMmaWarp) {
        // =============================== MMA issuer
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.NT, kMajorK, kMajorK);
            const uint32_t a_lbo = (uint32_t)p.P * 16u, b_lbo = (uint32_t)p.NT * 16u;
            const uint32_t hi = umma_desc_hi(128u);          // SBO = 128 B for both operands
            const uint32_t a_kstep = 2u * (uint32_t)p.P;     // address-field units (16 B): two channel groups
            const uint32_t b_step = 2u * (uint32_t)p.NT;
            uint32_t tapoff[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tapoff[tap] = (uint32_t)((tap / 3) * p.Wp + (tap % 3));
            int cc = 0, it = 0;
            PROF_DECL
            PROF_PROLOGUE(3)
            for (int tg = blockIdx.x; tg < p.ntg; tg += gridDim.x, ++it) {
                const int b = it & 1;
                const uint32_t k = (uint32_t)(it >> 1);
                mbar_wait(&acc_empty[b], (k & 1) ^ 1);
                PROF_MARK(0)
                tc_fence_after();
                const uint32_t acc = tmem_base + (uint32_t)(b * p.T * p.NT);
                for (int chunk = 0; chunk < p.nchunks; ++chunk, ++cc) {
                    const int s = cc % p.S;
                    const uint32_t ph = (uint32_t)(cc / p.S) & 1;
                    mbar_wait(&full[s], ph);
                    PROF_MARK(1)
                    tc_fence_after();
                    const uint32_t b_lo0 = umma_desc_lo(smem_u32(sB + (size_t)s * b_stride), b_lbo);
                    if constexpr (kTma) {
                        // A rows are positions (64 B each, 64B swizzle, 8-row atoms of 512 B); the stage holds whole padded rows
                        // starting at row R_lo, so position (q0 - Wp - 1) sits `woff` rows into it.  Address-field units are 16 B.
                        const int qb = tg * group_positions - p.Wp - 1;
                        const int R_lo = (qb >= 0) ? (int)fastdiv((uint32_t)qb, p.div_wp) : -((-qb + p.Wp - 1) / p.Wp);
                        const uint32_t woff = (uint32_t)(qb - R_lo * p.Wp);
                        const uint32_t a_hi_sw = (512u >> 4) | (1u << 14) | (4u << 29);   // SBO = 512 B, version 1, SWIZZLE_64B
                        const uint32_t a_base = ((smem_u32(sA + (size_t)s * a_bytes) >> 4) & 0x3FFFu) | (1u << 16);
                        for (int t = 0; t < p.T; ++t) {
                            const uint32_t d_addr = acc + (uint32_t)(t * p.NT);
                            uint32_t b_lo = b_lo0;
#pragma unroll
                            for (int tap = 0; tap < ((NUNET_DBG_SKIP & 2) ? 1 : 9); ++tap) {
                                const uint32_t a_lo = a_base + (woff + (uint32_t)(t * 128) + tapoff[tap]) * 4u;   // 64 B per position
                                umma_bf16_lohi(d_addr, a_lo, a_hi_sw, b_lo, hi, idesc, (uint32_t)((chunk | tap) != 0));
                                umma_bf16_lohi(d_addr, a_lo + 2u, a_hi_sw, b_lo + b_step, hi, idesc, 1u);         // next 16 channels: +32 B
                                b_lo += 2 * b_step;
                            }
                        }
                    } else {
                    const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA + (size_t)s * a_bytes), a_lbo);
                    for (int t = 0; t < p.T; ++t) {
                        const uint32_t d_addr = acc + (uint32_t)(t * p.NT);
                        const uint32_t a_t = a_lo0 + (uint32_t)(t * 128);
                        uint32_t b_lo = b_lo0;
                        if (cgs == 4) {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const uint32_t a_lo = a_t + tapoff[tap];
                                umma_bf16_lohi(d_addr, a_lo, hi, b_lo, hi, idesc, (uint32_t)((chunk | tap) != 0));
                                umma_bf16_lohi(d_addr, a_lo + a_kstep, hi, b_lo + b_step, hi, idesc, 1u);
                                b_lo += 2 * b_step;
                            }
                        } else {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                umma_bf16_lohi(d_addr, a_t + tapoff[tap], hi, b_lo, hi, idesc, (uint32_t)((chunk | tap) != 0));
                                b_lo += b_step;
                            }
                        }
                    }
                    }
                    umma_commit(&empty[s]);
                    PROF_MARK(2)
                }
                umma_commit(&acc_full[b]);
            }
            PROF_FLUSH(8)
        }
        __syncwarp();
    } else if (warp == kFetchWarp) {
        // =============================== weight fetcher (bulk async copy engine)
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)n_tile * p.nchunks * b_bytes;
            int cc = 0;
            PROF_DECL
            if constexpr (kTma) {
                for (int si = 0; si < p.num_src; ++si) tma_prefetch_desc(&p.tmap[si]);
            }
            for (int tg = blockIdx.x; tg < p.ntg; tg += gridDim.x) {
                int si = 0, src_c0 = 0;
                // TMA mode: the padded rows [R_lo, R_lo + nrows) cover positions [q0 - Wp - 1, q0 + 128T + Wp + 1)
                const int qb = tg * group_positions - p.Wp - 1;
                const int R_lo = (qb >= 0) ? (int)fastdiv((uint32_t)qb, p.div_wp) : -((-qb + p.Wp - 1) / p.Wp);
                const int nrows = (qb - R_lo * p.Wp + p.P + p.Wp - 1) / p.Wp;
                const uint32_t row_bytes = (uint32_t)p.Wp * 64u;
                // image / row of padded row R_lo (R = n*(H+1) + 1 + y); the row loop below advances them without dividing
                const int n_lo = (R_lo >= 0) ? (int)fastdiv((uint32_t)R_lo, p.div_hp1) : -1;
                const int y_lo = (R_lo >= 0) ? R_lo - n_lo * p.Hp1 - 1 : -1;
                for (int chunk = 0; chunk < p.nchunks; ++chunk, ++cc) {
                    const int s = cc % p.S;
                    const uint32_t ph = (uint32_t)(cc / p.S) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    PROF_MARK(0)
                    if constexpr (kTma) {
                        const int cglob = chunk * 32;
                        while (cglob >= src_c0 + p.src[si].C) { src_c0 += p.src[si].C; ++si; }
                        const bool fetch_b = !res_b || cc == 0;
                        mbar_arrive_expect_tx(&full[s], (fetch_b ? b_bytes : 0u) + ((NUNET_DBG_SKIP & 4) ? 0u : (uint32_t)nrows * row_bytes));
                        if (fetch_b) bulk_g2s(sB + (size_t)s * b_stride, wsrc + (size_t)chunk * b_bytes, b_bytes, &full[s]);
                        uint8_t* dstA = sA + (size_t)s * a_bytes;
                        int n = n_lo, y = y_lo, R = R_lo;
                        for (int r = 0; r < ((NUNET_DBG_SKIP & 4) ? 0 : nrows); ++r) {
                            // separator rows (y = -1) and rows outside the batch (n = -1 or n >= N) are out of bounds -> zeros
                            if (p.strips > 1) {
                                // "image" n is strip n % strips of image n / strips: the box starts one pixel left of the strip, so its
                                // halo columns are the neighbouring strips' pixels (zeros only at the image border)
                                const int img = (n >= 0) ? (int)fastdiv((uint32_t)n, p.div_strips) : -1;
                                const int st = (n >= 0) ? n - img * p.strips : 0;
                                tma_load_4d(dstA + (size_t)r * row_bytes, &p.tmap[si], cglob - src_c0, st * p.W - 1, y, img, &full[s]);
                            } else {
                                tma_load_4d(dstA + (size_t)r * row_bytes, &p.tmap[si], cglob - src_c0, -1, y, n, &full[s]);
                            }
                            ++R;
                            if (R <= 0) { n = (R == 0) ? 0 : -1; y = -1; }
                            else if (++y == p.H) { y = -1; ++n; }
                        }
                    } else {
                        if (!res_b || cc == 0) {
                            mbar_arrive_expect_tx(&full[s], b_bytes);
                            bulk_g2s(sB + (size_t)s * b_stride, wsrc + (size_t)chunk * b_bytes, b_bytes, &full[s]);
                        } else {
                            mbar_arrive(&full[s]);
                        }
                    }
                    PROF_MARK(1)
                }
            }
            PROF_FLUSH(12)
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        switch (p.tmem_cols) {
            case 32: tmem_dealloc<32>(tmem_base); break;
            case 64: tmem_dealloc<64>(tmem_base); break;
            case 128: tmem_dealloc<128>(tmem_base); break;
            case 256: tmem_dealloc<256>(tmem_base); break;
            default: tmem_dealloc<512>(tmem_base); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------ weight packing
// wpack[n_tile][chunk][tap][kg][NT][8]  (bf16);  FWD: n = co, k = ci ; DGRAD: n = ci, k = co, taps flipped.
__device__ __forceinline__ void pack_group(const float* __restrict__ w, int Cout, int Cin, int mode, int NT, int CK, int nchunks,
                                           __nv_bfloat16* __restrict__ out, size_t g) {
    if (mode >= NUNET_PACK_FWD_C32) {   // stacked layout of conv_c32.cu: [chunk][dr][kg 4][dc*32 + n][8], NT = 32, CK = 32
        size_t r = g;
        const int n96 = (int)(r % 96); r /= 96;
        const int kg = (int)(r % 4); r /= 4;
        const int dr = (int)(r % 3); r /= 3;
        const int chunk = (int)r;
        const int dc = n96 >> 5, n = n96 & 31, tap = dr * 3 + dc;
        const int k0 = chunk * 32 + kg * 8;
        __nv_bfloat16 v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + e;
            float f = 0.f;
            if (mode == NUNET_PACK_FWD_C32) {
                if (n < Cout && k < Cin) f = w[((size_t)n * Cin + k) * 9 + tap];
            } else {
                if (n < Cin && k < Cout) f = w[((size_t)k * Cin + n) * 9 + (8 - tap)];
            }
            v[e] = __float2bfloat16_rn(f);
        }
        *reinterpret_cast<uint4*>(out + g * 8) = *reinterpret_cast<const uint4*>(v);
        return;
    }
    const int cgs = CK >> 3;
    size_t r = g;
    const int nl = (int)(r % NT); r /= NT;
    const int kg = (int)(r % cgs); r /= cgs;
    const int tap = (int)(r % 9); r /= 9;
    const int chunk = (int)(r % nchunks); r /= nchunks;
    const int nt = (int)r;
    const int n = nt * NT + nl;
    const int k0 = chunk * CK + kg * 8;
    __nv_bfloat16 v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int k = k0 + e;
        float f = 0.f;
        if (mode == NUNET_PACK_FWD) {
            if (n < Cout && k < Cin) f = w[((size_t)n * Cin + k) * 9 + tap];
        } else {
            if (n < Cin && k < Cout) f = w[((size_t)k * Cin + n) * 9 + (8 - tap)];
        }
        v[e] = __float2bfloat16_rn(f);
    }
    *reinterpret_cast<uint4*>(out + g * 8) = *reinterpret_cast<const uint4*>(v);
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int mode, int NT, int CK, int nchunks,
                                    int n_tiles, __nv_bfloat16* __restrict__ out) {
    const size_t total = (size_t)n_tiles * nchunks * 9 * (CK >> 3) * NT;  // 16-byte groups
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x)
        pack_group(w, Cout, Cin, mode, NT, CK, nchunks, out, g);
}

// All layers of a step in ONE launch (the optimizer touches every weight every step; 60 separate ~8 us launches otherwise).
struct PackBatch {
    nunet_pack_job_t jobs[NUNET_MAX_PACK_JOBS];
    int start[NUNET_MAX_PACK_JOBS + 1];   // first block of each job (blocks of 256 groups)
    int njobs;
};
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const __grid_constant__ PackBatch b) {
    int j = 0;
    while (j + 1 < b.njobs && (int)blockIdx.x >= b.start[j + 1]) ++j;
    const nunet_pack_job_t& jb = b.jobs[j];
    const size_t total = (size_t)jb.n_tiles * jb.nchunks * 9 * (jb.CK >> 3) * jb.NT;
    const size_t g = (size_t)(blockIdx.x - b.start[j]) * 256 + threadIdx.x;
    if (g < total) pack_group(jb.w, jb.Cout, jb.Cin, jb.mode, jb.NT, jb.CK, jb.nchunks, reinterpret_cast<__nv_bfloat16*>(jb.out), g);
}

}  // namespace nunet

// ------------------------------------------------------------------------------------------------ host side
using namespace nunet;

static int pick_nt(int C) {
    if (C <= 256) return C;
    for (int nt = 256; nt >= 32; nt -= 32)
        if (C % nt == 0) return nt;
    return 0;
}

// Resident weight block for single-chunk layers: measured 3-13 % faster for N tiles >= 96 (the 55-110 KB block no longer
// cycles through the stage ring; it also frees a stage's worth of shared memory for the sixteen-warp epilogue scratch), no gain for
// narrower tiles (profiles/r02_conv_tiling_sweep.log) -> used for NT >= 96.
static bool resident_weights(int NT) {
    static const bool on = getenv("NUNET_NO_RESIDENT_B") == nullptr;
#ifdef NUNET_EXPERIMENT
    if (const char* e = getenv("NUNET_RESB_MIN_NT")) return on && NT >= atoi(e);
#endif
    return on && NT >= 96;
}

static size_t conv_smem_bytes(int T, int NT, int CK, int S, int P, int cin_pad) {
    return conv_smem_layout(false, T, NT, CK, S, P, cin_pad, 0, cin_pad == CK && resident_weights(NT), 8).total;
}

static size_t conv_tma_a_stage(int P, int W) {
    return (((size_t)(P + 2 * (W + 2)) * 64) + 1023) & ~(size_t)1023;  // whole padded rows: up to 2 Wp - 2 extra positions
}
static size_t conv_tma_smem_bytes(int T, int NT, int S, int P, int W, int cin_pad, int epi_warps) {
    return (size_t)conv_smem_layout(true, T, NT, 32, S, P, cin_pad, (int)conv_tma_a_stage(P, W), cin_pad == 32 && resident_weights(NT), epi_warps).total + 1024;
}

extern "C" int nunet_conv3x3_plan(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout, nunet_conv_plan_t* pl) {
    if (!pl || N <= 0 || H <= 0 || W <= 0 || cin_total <= 0 || Cout <= 0) return NUNET_E_BADARG;
    if (Cout % 32 != 0) return NUNET_E_UNSUPPORTED;
    if (H >= 65536 || W >= 65536) return NUNET_E_UNSUPPORTED;
    const int cin_pad = (cin_total <= 16) ? 16 : cin_total;
    if (cin_pad % 16 != 0) return NUNET_E_UNSUPPORTED;
    if ((long long)N * H * W * (Cout > cin_pad ? Cout : cin_pad) >= (1LL << 40)) return NUNET_E_UNSUPPORTED;
    // Rows wider than a tensor-load box (W + 2 > 256 positions): the batch is described as N*strips images of ceil(W/strips) columns;
    // the box of a strip row starts one pixel to its left, so the halo columns hold the neighbouring strips' pixels.  Needs 32-channel
    // chunks (tensor loads); falls back to the whole-row geometry (cp.async / transforming loaders) when no such tiling fits.
    const int real_N = N, real_W = W;
    static const bool no_strips = getenv("NUNET_NO_STRIPS") != nullptr;
    int strips = (W + 2 > 256 && cin_pad % 32 == 0 && !no_strips) ? (W + 253) / 254 : 1;
retry_geometry:
    N = real_N * strips;
    W = (real_W + strips - 1) / strips;
    if (strips > 1) W += W & 1;      // even strip width: every row of a stage starts 128-byte aligned (tensor-load destination)
    const long long Q = ((long long)N * (H + 1) + 1) * (W + 2);
    if (Q + 4LL * 128 + 2 * (W + 2) + 2 >= (1LL << 31)) return NUNET_E_UNSUPPORTED;
    int NT = pick_nt(Cout);
    if (NT == 0) return NUNET_E_UNSUPPORTED;
#ifdef NUNET_EXPERIMENT
    if (const char* e = getenv("NUNET_CONV_NT")) { const int v = atoi(e); if (v >= 32 && v % 32 == 0 && Cout % v == 0 && v < NT) NT = v; }
#endif
    // (256-wide tiles only fit shared memory with 16-channel chunks, i.e. on the cp.async loader path; two 128-wide tiles on tensor
    // loads were measured 1.5-2.8x SLOWER at 12x12 / 6x6 — 256->256: 0.088 vs 0.058 ms, 512->512: 0.183 vs 0.066 ms — the weight
    // stream per position doubles; profiles/r02_summary.md)
    const size_t budget = 220 * 1024;
    int n_tiles = 0, bestT = 0, bestCK = 0, bestS = 0;
    auto search = [&](int NT) {
    n_tiles = Cout / NT; bestT = 0; bestCK = 0; bestS = 0;
    // prefer: large T (halo amortisation) while keeping >= 2 waves of CTAs, CK=32, 2 stages
    for (int minS = 2; minS >= 1 && !bestT; --minS)
    for (int T = (256 / NT < 4 ? 256 / NT : 4); T >= 1 && !bestT; --T) {
        if (T == 3) continue;
        const long long groups = ((Q + 128LL * T - 1) / (128LL * T)) * n_tiles;
        if (T > 1 && groups < 4 * 148) continue;  // keep >= 4 tile groups per SM for balance and pipelining
        for (int CK = 32; CK >= 16 && !bestT; CK -= 16) {
            if (cin_pad % CK != 0) continue;
            const int nch = cin_pad / CK;
            const int P = 128 * T + 2 * (W + 2) + 2;
            // the stage ring runs across tile groups, so S may exceed the chunks per group; with a resident weight block (level-0 dgrads)
            // deeper rings measured 35-50 % SLOWER (32->192 at 96x96: S=2 0.264 ms, S=3 0.399 ms; profiles/r02_conv_tiling_sweep.log)
            for (int S = ((nch == 1 && resident_weights(NT)) ? 2 : 3); S >= minS; --S) {
                if ((size_t)P * 16 >= (1u << 18)) continue;
                if (strips > 1 ? CK != 32 : conv_smem_bytes(T, NT, CK, S, P, cin_pad) > budget) continue;   // strips: only the TMA layout (below) is ever launched
                // rows that fit a tensor-load box: the same (T, S) must also fit the TMA variant's larger stages (copy-mode sources
                // then run on tensor loads with sixteen epilogue warps; a deeper ring that only fits the cp.async layout is slower)
                if (CK == 32 && W + 2 <= 256 && conv_tma_smem_bytes(T, NT, S, P, W, cin_pad, 8) > 227 * 1024) continue;
                bestT = T; bestCK = CK; bestS = S;
                break;
            }
        }
    }
    };
    search(NT);
    // Wide outputs whose 160..256-column weight stage only fits with 16-channel chunks fall off the tensor-load path (cp.async
    // loaders, eight epilogue warps).  Where the layer is large enough for the weight stream not to matter, narrower tiles on tensor
    // loads win: 64-wide at >= 3000 tiles (48x48: 64->320 0.313 -> 0.226 ms, 64->192 0.171 -> 0.134), 128-wide at >= 1000 tiles
    // (24x24: 128->384 0.149 -> 0.123); at 12x12 / 6x6 the default stays (256->768: 0.128 vs 0.197 / 0.334 ms).
    if ((bestCK != 32 || bestS < 2) && W + 2 <= 256 && cin_pad % 32 == 0) {
        const long long tiles = (Q + 127) / 128;
        const int alt = (tiles >= 3000 && Cout % 64 == 0 && NT > 64) ? 64 : ((tiles >= 1000 && Cout % 128 == 0 && NT > 128) ? 128 : 0);
        if (alt) {
            const int keepT = bestT, keepCK = bestCK, keepS = bestS;
            search(alt);
            if (bestT && bestCK == 32 && bestS >= 2) NT = alt;
            else { n_tiles = Cout / NT; bestT = keepT; bestCK = keepCK; bestS = keepS; }
        }
    }
    if (strips > 1 && (!bestT || bestCK != 32 || conv_tma_smem_bytes(bestT, NT, bestS, 128 * bestT + 2 * (W + 2) + 2, W, cin_pad, 8) > 227 * 1024)) {
        strips = 1;
        goto retry_geometry;
    }
    if (!bestT) return NUNET_E_SMEM;
#ifdef NUNET_EXPERIMENT
    // tiling overrides for sweeps (tools/sweep_conv.sh); never compiled into the shipped library
    if (const char* e = getenv("NUNET_CONV_T")) {
        const int T = atoi(e), P = 128 * T + 2 * (W + 2) + 2;
        int S = getenv("NUNET_CONV_S") ? atoi(getenv("NUNET_CONV_S")) : bestS;
        if (T >= 1 && 2 * T * NT <= 512 && cin_pad % 32 == 0) {
            while (S > 1 && conv_smem_bytes(T, NT, 32, S, P, cin_pad) > budget) --S;
            if (conv_smem_bytes(T, NT, 32, S, P, cin_pad) <= budget) { bestT = T; bestCK = 32; bestS = S; }
        }
    }
#endif
    pl->T = bestT; pl->NT = NT; pl->CK = bestCK; pl->S = bestS;
    pl->P = 128 * bestT + 2 * (W + 2) + 2;
    pl->nchunks = cin_pad / bestCK;
    pl->n_tiles = n_tiles;
    pl->ntg = (int)((Q + 128LL * bestT - 1) / (128LL * bestT));
    {
        int ctas_x = 148 / n_tiles;  // persistent: one CTA per SM in total
        if (ctas_x < 1) ctas_x = 1;
        pl->grid_m = pl->ntg < ctas_x ? pl->ntg : ctas_x;
    }
    pl->cin_pad = cin_pad;
    pl->smem_bytes = (int)conv_smem_bytes(bestT, NT, bestCK, bestS, pl->P, cin_pad);
    // TMA variant of the same tiling: A stage = (P + Wp) rows of 64 B rounded to 1 KB (first in shared memory: 1 KB aligned)
    pl->tma_a_stage_bytes = 0;
    pl->tma_smem_bytes = 0;
    pl->tma_epi_warps = 0;
    pl->strips = strips;
    if (bestCK == 32 && W + 2 <= 256) {
        const size_t a_tma = conv_tma_a_stage(pl->P, W);
        for (int ew = 16; ew >= 8 && !pl->tma_smem_bytes; ew -= 8) {
            const size_t total = conv_tma_smem_bytes(bestT, NT, bestS, pl->P, W, cin_pad, ew);
            if (total <= 227 * 1024) { pl->tma_a_stage_bytes = (int)a_tma; pl->tma_smem_bytes = (int)total; pl->tma_epi_warps = ew; }
        }
    }
    if (strips > 1) pl->smem_bytes = pl->tma_smem_bytes;   // the only layout a strip plan is launched with
    int cols = 2 * bestT * NT, tc = 32;
    while (tc < cols) tc <<= 1;
    pl->tmem_cols = tc;
    pl->wpack_elems = (int64_t)n_tiles * pl->nchunks * 9 * bestCK * NT;
    return NUNET_OK;
}

template <int MODE>
static int launch_conv(const ConvParams& p, const nunet_conv_plan_t& pl, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_igemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    conv3x3_igemm_kernel<MODE><<<dim3(pl.grid_m, pl.n_tiles), kThreads, pl.smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}

int nunet_conv_c32_launch(const nunet_conv_desc_t* d, int cin, int grid, const nunet::BnTail& tail, cudaStream_t stream);

extern "C" int nunet_conv3x3_fwd(const nunet_conv_desc_t* d, void* stream) {
    if (!d || d->num_src < 1 || d->num_src > NUNET_MAX_SRC || !d->wpack || (!d->out && d->num_out_seg <= 0)) return NUNET_E_BADARG;
    int cin = 0;
    for (int i = 0; i < d->num_src; ++i) {
        const nunet_src_t& s = d->src[i];
        if (!s.data) return NUNET_E_BADARG;
        if (s.kind == NUNET_SRC_INPUT) {
            if (s.C < 1 || s.C > 16 || d->num_src != 1) return NUNET_E_UNSUPPORTED;
            cin += 16;
        } else {
            const bool act = (s.kind == NUNET_SRC_ACT || s.kind == NUNET_SRC_POOL_ACT || s.kind == NUNET_SRC_UP_ACT);
            if (s.C % 32 != 0 || (!act && (!s.scale || !s.shift))) return NUNET_E_BADARG;
            if (s.kind == NUNET_SRC_GRAD && (!s.data2 || !s.k2)) return NUNET_E_BADARG;
            if ((s.kind == NUNET_SRC_UP || s.kind == NUNET_SRC_UP_ACT) && ((d->H | d->W) & 1)) return NUNET_E_UNSUPPORTED;
            cin += s.C;
        }
    }
    nunet_conv_plan_t pl;
    int rc = nunet_conv3x3_plan(d->N, d->H, d->W, cin, d->Cout, &pl);
    if (rc) return rc;
    ConvParams p{};
    const int strips = pl.strips > 1 ? pl.strips : 1;
    p.strips = strips; p.Wreal = d->W; p.div_strips = make_fastdiv((uint32_t)(strips > 1 ? strips : 2));
    p.N = d->N * strips; p.H = d->H; p.W = (d->W + strips - 1) / strips; p.Hp1 = d->H + 1;
    if (strips > 1) p.W += p.W & 1;
    p.Wp = p.W + 2;
    p.div_wp = make_fastdiv((uint32_t)p.Wp); p.div_hp1 = make_fastdiv((uint32_t)p.Hp1);
    p.Q = (p.N * (d->H + 1) + 1) * (p.W + 2);
    p.T = pl.T; p.NT = pl.NT; p.CK = pl.CK; p.S = pl.S; p.P = pl.P; p.nchunks = pl.nchunks;
    p.cin_pad = pl.cin_pad; p.Cout = d->Cout; p.num_src = d->num_src; p.tmem_cols = pl.tmem_cols; p.ntg = pl.ntg;
    for (int i = 0; i < d->num_src; ++i) {
        p.src[i].data = d->src[i].data; p.src[i].data2 = d->src[i].data2;
        p.src[i].scale = d->src[i].scale; p.src[i].shift = d->src[i].shift; p.src[i].k2 = d->src[i].k2;
        p.src[i].C = d->src[i].C; p.src[i].kind = d->src[i].kind;
    }
    p.wpack = d->wpack; p.out = d->out; p.stat_part = d->stat_part;
    if (!bn_tail_ok(d->bn, d->stat_part)) return NUNET_E_BADARG;
    p.tail = make_bn_tail(d->bn);
    if (d->num_out_seg < 0 || d->num_out_seg > NUNET_MAX_SRC) return NUNET_E_BADARG;
    if (d->num_out_seg > 0) {   // output channels scattered to one dense tensor per segment (plain epilogue only)
        int tot = 0;
        for (int i = 0; i < d->num_out_seg; ++i) {
            if (!d->out_seg[i] || d->out_seg_c[i] < 32 || d->out_seg_c[i] % 32 != 0) return NUNET_E_BADARG;
            p.out_seg[i] = d->out_seg[i]; p.out_seg_c[i] = d->out_seg_c[i];
            tot += d->out_seg_c[i];
        }
        if (tot != d->Cout || d->stat_part || d->mask_y) return NUNET_E_BADARG;
        p.num_out_seg = d->num_out_seg;
    }
    if (d->mask_y) {            // fused block-internal gradient router: ReLU mask of the producer + BatchNorm-backward sums
        if (!d->stat_part || !d->mask_scale || !d->mask_shift || !d->mask_mean || !d->mask_invstd || d->bn.counter) return NUNET_E_BADARG;
        p.mask_y = d->mask_y; p.mask_scale = d->mask_scale; p.mask_shift = d->mask_shift; p.mask_mean = d->mask_mean; p.mask_invstd = d->mask_invstd;
        const nunet_bn_bwd_tail_t& b = d->bn_bwd;
        if (b.counter) {
            if (b.count <= 0 || (b.gamma ? (!b.save_mean || !b.save_invstd || !b.dgamma || !b.dbeta || !b.k1 || !b.k2 || !b.k3) : !b.sums_out)) return NUNET_E_BADARG;
            p.tail.counter = b.counter; p.tail.sums_out = b.sums_out;
            p.tail.fin.mode = 1; p.tail.fin.count = b.count; p.tail.fin.gamma = b.gamma;
            p.tail.fin.save_mean = const_cast<float*>(b.save_mean); p.tail.fin.save_invstd = const_cast<float*>(b.save_invstd);
            p.tail.fin.dgamma = b.dgamma; p.tail.fin.dbeta = b.dbeta; p.tail.fin.k1 = b.k1; p.tail.fin.k2 = b.k2; p.tail.fin.k3 = b.k3;
            p.tail.fin.grad_scale = b.grad_scale;
            if (!exchange_ok(b.xch)) return NUNET_E_BADARG;
            copy_exchange(p.tail.xch, b.xch);
        }
    }
    p.res_b = (pl.nchunks == 1 && resident_weights(pl.NT)) ? 1 : 0;
    int kinds[NUNET_MAX_SRC];
    for (int i = 0; i < d->num_src; ++i) kinds[i] = d->src[i].kind;
    const int mode = sources_mode(kinds, d->num_src);
    if (d->wpack_layout != 0 && (d->wpack_layout != 1 || mode != MODE_ACT)) return NUNET_E_BADARG;
    if (strips > 1 && mode != MODE_ACT) return NUNET_E_UNSUPPORTED;   // column strips: tensor-load sources only (see nunet_conv_plan_t.strips)
    switch (mode) {
        case MODE_INPUT: return launch_conv<MODE_INPUT>(p, pl, (cudaStream_t)stream);
        case MODE_SAME: return launch_conv<MODE_SAME>(p, pl, (cudaStream_t)stream);
        case MODE_POOL: return launch_conv<MODE_POOL>(p, pl, (cudaStream_t)stream);
        case MODE_SAME_UP: return launch_conv<MODE_SAME_UP>(p, pl, (cudaStream_t)stream);
        case MODE_GRAD: return launch_conv<MODE_GRAD>(p, pl, (cudaStream_t)stream);
        case MODE_ACT: {
            if (d->wpack_layout == 1) {
                if (!nunet_conv3x3_c32_eligible(d->N, d->H, d->W, cin, d->Cout) || d->num_out_seg) return NUNET_E_BADARG;
                return nunet_conv_c32_launch(d, cin, pl.grid_m, p.tail, (cudaStream_t)stream);
            }
            // tensor-load path: box = (32 ch, W+2 px, 1, 1) per padded row, needs W+2 <= 256 and the CK=32 tiling
            static const bool no_tma = getenv("NUNET_NO_TMA") != nullptr;
            if (!no_tma && pl.CK == 32 && p.W + 2 <= 256 && (p.W & 1) == 0 && pl.tma_smem_bytes > 0) {   // odd rows: 64-byte-aligned stage rows, not a TMA destination
                bool ok = true;
                for (int i = 0; i < d->num_src && ok; ++i)   // maps over the REAL tensors; the box is one (strip) row + its two halo columns
                    ok = encode_nhwc_map(&p.tmap[i], d->src[i].data, d->N, d->H, d->W, d->src[i].C, 32, p.W + 2, 64) == 0;
                if (ok) {
                    nunet_conv_plan_t pt = pl;
                    pt.smem_bytes = pl.tma_smem_bytes;
                    p.a_stage_bytes = pl.tma_a_stage_bytes;
                    p.epi_warps = pl.tma_epi_warps;
                    return launch_conv<MODE_ACT_TMA>(p, pt, (cudaStream_t)stream);
                }
            }
            if (strips > 1) return NUNET_E_UNSUPPORTED;   // the strip geometry exists on tensor loads only
            return launch_conv<MODE_ACT>(p, pl, (cudaStream_t)stream);
        }
        case MODE_POOL_ACT: return launch_conv<MODE_POOL_ACT>(p, pl, (cudaStream_t)stream);
        case MODE_ACT_UP: return launch_conv<MODE_ACT_UP>(p, pl, (cudaStream_t)stream);
    }
    return NUNET_E_UNSUPPORTED;  // source-kind combination outside {input | same* | pool | same*+up | grad}
}

#ifdef NUNET_PROFILE_WAITS
extern "C" int nunet_debug_read_wait_prof(unsigned long long* out16, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_wait_prof, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    if (reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_wait_prof, z, sizeof(z));
    }
    return e == cudaSuccess ? 0 : NUNET_E_CUDA + (int)e;
}
#endif

extern "C" int nunet_pack_weights(const float* w, int32_t Cout, int32_t Cin, int32_t mode, const nunet_conv_plan_t* pl,
                                  void* wpack, void* stream) {
    if (!w || !pl || !wpack || mode < NUNET_PACK_FWD || mode > NUNET_PACK_DGRAD_C32) return NUNET_E_BADARG;
    if (mode >= NUNET_PACK_FWD_C32 && (pl->NT != 32 || pl->CK != 32 || pl->n_tiles != 1)) return NUNET_E_BADARG;
    const size_t groups = (size_t)pl->wpack_elems / 8;
    int blocks = (int)((groups + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, mode, pl->NT, pl->CK, pl->nchunks, pl->n_tiles,
                                                                   reinterpret_cast<__nv_bfloat16*>(wpack));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}

extern "C" int nunet_pack_weights_batched(const nunet_pack_job_t* jobs, int32_t njobs, void* stream) {
    if (!jobs || njobs < 1) return NUNET_E_BADARG;
    for (int j0 = 0; j0 < njobs; j0 += NUNET_MAX_PACK_JOBS) {
        PackBatch b{};
        const int nj = njobs - j0 < NUNET_MAX_PACK_JOBS ? njobs - j0 : NUNET_MAX_PACK_JOBS;
        int blocks = 0;
        for (int j = 0; j < nj; ++j) {
            const nunet_pack_job_t& jb = jobs[j0 + j];
            if (!jb.w || !jb.out || jb.mode < NUNET_PACK_FWD || jb.mode > NUNET_PACK_DGRAD_C32 || jb.NT < 1 || jb.CK < 8 || jb.nchunks < 1 || jb.n_tiles < 1)
                return NUNET_E_BADARG;
            b.jobs[j] = jb;
            b.start[j] = blocks;
            const size_t total = (size_t)jb.n_tiles * jb.nchunks * 9 * (jb.CK >> 3) * jb.NT;
            blocks += (int)((total + 255) / 256);
        }
        b.start[nj] = blocks;
        b.njobs = nj;
        pack_weights_batched_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(b);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    }
    return NUNET_OK;
}
