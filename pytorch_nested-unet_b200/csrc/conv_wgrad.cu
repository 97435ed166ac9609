// Weight gradient of the fused 3x3 convolution on tcgen05 tensor cores (sm_100a), split-K over positions.
//
//   dW[co][ci][tap] = sum_q gy[q][co] * X[q + off(tap)][ci] = sum_q' X[q'][ci] * gy[q' - off(tap)][co]
// over the flattened padded positions of conv_igemm.cu.  The reduction index (GEMM K) is the POSITION, so both
// operands are MN-major for the tensor core.  Per chunk of KP positions the producers stage
//   X   [cgroup(16)][KP][8 ci]            — the conv's virtual input through the FORWARD loaders, unshifted
//   gy  [dc(3)][cgroup(4)][KP + 2Wp][8 co] — the materialised BatchNorm-backward gradient (nunet_bn_bwd_apply), copied
//        three times with a one-position stagger, so that for a kernel row dr the twelve (dc, cgroup) planes form ONE
//        MN-major operand of N = 96 whose start address encodes the row shift (2-dr)*Wp.
// Each K step (16 positions) therefore issues 3 MMAs of 128x96x16 (~50 cycles each) instead of 9 of 128x32x16
// (~45 cycles each: the tensor pipe is operand-fetch bound at N=32, tools/umma_probe.cu test 100).
// A CTA owns a (128 ci) x (32 co) x (9 taps) accumulator block = 288 TMEM columns (column = tap*32 + co) and a
// contiguous range of chunks; partial blocks are written as fp32 and summed in a fixed order (deterministic).
#include "conv_common.cuh"
#include "loaders.cuh"
#include <cstdlib>

namespace nunet {

struct WgradParams {
    int N, H, W, Wp, Hp1;
    long long Q;
    int KP, S, Pg, Pg2, cpc, nchunks_total;  // Pg2: positions per gy plane (TRIPLE: KP + 2Wp, else Pg = KP + 2Wp + 2)
    int cin_pad, Cout, num_src;
    ConvSrc src[NUNET_MAX_SRC];
    const __nv_bfloat16* gy;  // bf16 NHWC [N,H,W,Cout]
    float* part;              // [splits][Cout/32][9][cin_pad][32]
};

#ifdef NUNET_PROFILE_WAITS
__device__ unsigned long long g_wgrad_prof[16];
#define WPROF_DECL long long prof_t = clock64(); long long prof_acc[6] = {0, 0, 0, 0, 0, 0};
#define WPROF_MARK(slot) { long long now_ = clock64(); prof_acc[slot] += now_ - prof_t; prof_t = now_; }
#define WPROF_FLUSH(base, n) { for (int i_ = 0; i_ < n; ++i_) atomicAdd(&g_wgrad_prof[(base) + i_], (unsigned long long)prof_acc[i_]); }
#else
#define WPROF_DECL
#define WPROF_MARK(slot) {}
#define WPROF_FLUSH(base, n) {}
#endif

#ifndef NUNET_WGRAD_PRODUCER_WARPS
#define NUNET_WGRAD_PRODUCER_WARPS 8
#endif
constexpr int kWProducerWarps = NUNET_WGRAD_PRODUCER_WARPS;   // warps 0..P-1 producers (then epilogue: first 8), warp P: MMA issuer
constexpr int kWThreads = (kWProducerWarps + 2) * 32;

// TRIPLE: three staggered gy copies, 3 MMAs of N=96 per K step; else one gy copy with the full halo and 9 MMAs of N=32
// per K step (needs a third of the shared memory for gy: wide images).
template <int MODE, bool TRIPLE>
__global__ void __launch_bounds__(kWThreads, 1) conv3x3_wgrad_kernel(const WgradParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t x_bytes = 16u * p.KP * 16u;
    const uint32_t g_bytes = (TRIPLE ? 12u : 4u) * p.Pg2 * 16u;
    uint8_t* sX = smem;
    uint8_t* sG = sX + (size_t)p.S * x_bytes;
    PixInfo* tabs = reinterpret_cast<PixInfo*>(sG + (size_t)p.S * g_bytes);  // [S][Pg]
    float* s_scale = reinterpret_cast<float*>(tabs + (size_t)p.S * p.Pg);    // [cin_pad]
    float* s_shift = s_scale + p.cin_pad;
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + p.cin_pad);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.S;
    uint64_t* accum = bars + 2 * p.S;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * p.S + 1);

    const int split = blockIdx.x, m_tile = blockIdx.y, n_tile = blockIdx.z;
    const int chunk0 = split * p.cpc;
    const int nchunks = min(p.cpc, p.nchunks_total - chunk0);
    const int ci0 = m_tile * 128;
    const int mcg = min(16, (p.cin_pad - ci0) >> 3);  // staged channel groups of this M tile

    if (warp == kWProducerWarps) {
        if (lane == 0) {
            for (int s = 0; s < p.S; ++s) {
                mbar_init(&full[s], kWProducerWarps);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum, 1);
            mbar_init_fence();
        }
        __syncwarp();
        tmem_alloc<512>(tmem_ptr);
    }
    {
        int cbase = 0;
        for (int si = 0; si < p.num_src; ++si) {
            const ConvSrc& s = p.src[si];
            const int cpad = (s.kind == NUNET_SRC_INPUT) ? 16 : s.C;
            for (int c = threadIdx.x; c < cpad; c += kWThreads) {
                const bool has = (s.kind != NUNET_SRC_INPUT) && s.scale != nullptr;
                s_scale[cbase + c] = has ? s.scale[c] : 1.f;
                s_shift[cbase + c] = has ? s.shift[c] : 0.f;
            }
            cbase += cpad;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < kWProducerWarps) {
        const int pl = lane & 7, r = lane >> 3;
        const int tid = threadIdx.x;  // producer thread index
        WPROF_DECL
        for (int chunk = 0; chunk < nchunks; ++chunk) {
            const int s = chunk % p.S;
            const uint32_t ph = (chunk / p.S) & 1;
            const int k0 = (chunk0 + chunk) * p.KP;  // first position of the chunk
            PixInfo* tab = tabs + (size_t)s * p.Pg;
            mbar_wait(&empty[s], ph ^ 1);
            WPROF_MARK(0)
            for (int i = tid; i < p.Pg; i += kWProducerWarps * 32) tab[i] = decode_position(k0 - p.Wp - 1 + i, p.N, p.H, p.W, p.Wp, p.Hp1, (int)p.Q);
            asm volatile("bar.sync 1, %0;" ::"n"(kWProducerWarps * 32) : "memory");
            WPROF_MARK(1)
            // ---- gy, all cp.async.  TRIPLE: three staggered copies (dc = 0,1,2) of 4 channel groups x Pg2 positions:
            //      plane (dc, cg) entry i holds position k0 - Wp - 1 + i + (2 - dc)  (= table index i + 2 - dc).
            //      else: one copy, entry i = table index i.
            {
                const __nv_bfloat16* gbase = p.gy + n_tile * 32 + r * 8;
                const int nblk = (p.Pg2 + 7) >> 3;
#pragma unroll
                for (int dc = 0; dc < (TRIPLE ? 3 : 1); ++dc) {
                    uint8_t* dst = sG + (size_t)s * g_bytes + (size_t)(dc * 4 + r) * p.Pg2 * 16;
                    const int toff = TRIPLE ? (2 - dc) : 0;
                    for (int blk = warp; blk < nblk; blk += kWProducerWarps * 4) {
                        int pix[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int pp = (blk + u * kWProducerWarps) * 8 + pl;
                            pix[u] = (pp < p.Pg2) ? tab[pp + toff].pix : -2;
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int pp = (blk + u * kWProducerWarps) * 8 + pl;
                            if (pix[u] != -2) cp_async16(dst + (size_t)pp * 16, gbase + (size_t)(pix[u] >= 0 ? pix[u] : 0) * p.Cout, pix[u] >= 0 ? 16u : 0u);
                        }
                    }
                }
            }
            WPROF_MARK(2)
            // ---- X: mcg channel groups x KP positions (positions [Wp+1, Wp+1+KP) of the table), one quad of channel
            //      groups (= 32 channels, never straddling a source) per pass
            {
                const int nquads = (mcg + 3) >> 2;
                for (int quad = 0; quad < nquads; ++quad) {
                    const int cg = quad * 4 + r;
                    if (cg >= mcg) continue;
                    const int cglob = ci0 + cg * 8;
                    int si = 0, c0 = 0;
                    while (true) {
                        const int cpad = (p.src[si].kind == NUNET_SRC_INPUT) ? 16 : p.src[si].C;
                        if (cglob < c0 + cpad) break;
                        c0 += cpad;
                        ++si;
                    }
                    uint8_t* dst = sX + (size_t)s * x_bytes + (size_t)cg * p.KP * 16;
                    stage_blocks_mode<MODE>(p.src[si], p.H, p.W, tab, p.Wp + 1, p.KP, warp, p.KP >> 3, kWProducerWarps, pl, cglob - c0, s_scale + cglob,
                                            s_shift + cglob, s_scale + cglob, dst);
                }
            }
            cp_async_commit();
            WPROF_MARK(3)
            cp_async_wait<0>();
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
            WPROF_MARK(4)
        }
        if (tid == 0) WPROF_FLUSH(0, 5)
        // ---- epilogue: TMEM -> fp32 partial block (first 8 producer warps)
        if (warp >= 8) goto done;
        mbar_wait(accum, 0);
        tc_fence_after();
        const int lq = warp & 3, half = warp >> 2;
        const int ci = ci0 + lq * 32 + lane;
        float* dst_base = p.part + ((size_t)(split * (p.Cout >> 5) + n_tile) * 9) * p.cin_pad * 32;
        for (int tap = half; tap < 9; tap += 2) {
            uint32_t v[32];
            tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(tap * 32), v);
            tmem_ld_wait();
            if (ci < p.cin_pad) {
                uint4* d = reinterpret_cast<uint4*>(dst_base + ((size_t)tap * p.cin_pad + ci) * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) d[j] = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            }
        }
    } else if (warp == kWProducerWarps) {
        if (elect_one()) {   // elect.sync: the compiler then knows ONE thread runs this (no per-instruction convergence loop around tcgen05 / TMA ops)
            const uint32_t idesc = umma_idesc_bf16(128, TRIPLE ? 96 : 32, kMajorMN, kMajorMN);
            const uint32_t a_hi = umma_desc_hi((uint32_t)p.KP * 16u), b_hi = umma_desc_hi((uint32_t)p.Pg2 * 16u);
            const int ksteps = p.KP >> 4;
            uint32_t tapoff[9];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) tapoff[tap] = (uint32_t)((2 - tap / 3) * p.Wp + (2 - tap % 3));
            WPROF_DECL
            for (int chunk = 0; chunk < nchunks; ++chunk) {
                const int s = chunk % p.S;
                const uint32_t ph = (chunk / p.S) & 1;
                mbar_wait(&full[s], ph);
                WPROF_MARK(0)
                tc_fence_after();
                uint32_t a_lo = umma_desc_lo(smem_u32(sX + (size_t)s * x_bytes), 128u);
                uint32_t b_lo = umma_desc_lo(smem_u32(sG + (size_t)s * g_bytes), 128u);
                // one K step = 16 positions = 16 address units
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t accf = (uint32_t)((chunk | ks) != 0);
                    if constexpr (TRIPLE) {  // kernel row dr reads the planes shifted by (2-dr)*Wp positions
                        umma_bf16_lohi(tmem_base + 0u, a_lo, a_hi, b_lo + 2u * (uint32_t)p.Wp, b_hi, idesc, accf);
                        umma_bf16_lohi(tmem_base + 96u, a_lo, a_hi, b_lo + (uint32_t)p.Wp, b_hi, idesc, accf);
                        umma_bf16_lohi(tmem_base + 192u, a_lo, a_hi, b_lo, b_hi, idesc, accf);
                    } else {
#pragma unroll
                        for (int tap = 0; tap < 9; ++tap)
                            umma_bf16_lohi(tmem_base + (uint32_t)(tap * 32), a_lo, a_hi, b_lo + tapoff[tap], b_hi, idesc, accf);
                    }
                    a_lo += 16u;
                    b_lo += 16u;
                }
                umma_commit(&empty[s]);
                WPROF_MARK(1)
            }
            umma_commit(accum);
            WPROF_FLUSH(8, 2)
        }
        __syncwarp();
    }

done:
    tc_fence_before();
    __syncthreads();
    if (warp == kWProducerWarps) tmem_dealloc<512>(tmem_base);
}

// dW[co][ci][tap] (fp32 OIHW, ci < Cin real) = sum over splits, fixed order.
__global__ void wgrad_reduce_kernel(const float* __restrict__ part, int splits, int Cout, int cin_pad, int Cin, float* __restrict__ dw) {
    const size_t block_elems = (size_t)(Cout >> 5) * 9 * cin_pad * 32;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < block_elems; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int col = (int)(r & 31); r >>= 5;
        const int ci = (int)(r % cin_pad); r /= cin_pad;
        const int tap = (int)(r % 9); r /= 9;
        const int nt = (int)r;
        if (ci >= Cin) continue;
        float s = 0.f;
        for (int k = 0; k < splits; ++k) s += part[(size_t)k * block_elems + i];
        dw[((size_t)(nt * 32 + col) * Cin + ci) * 9 + tap] = s;
    }
}

}  // namespace nunet

using namespace nunet;

int nunet_wgrad_tma_plan(int N, int H, int W, int cin_pad, int Cout, nunet_wgrad_plan_t* pl);
int nunet_wgrad_tma_launch(const nunet_wgrad_desc_t* d, const nunet_wgrad_plan_t& pl, int cin_real, cudaStream_t stream);

static size_t wgrad_smem_bytes(int KP, int S, int Pg, int planes, int Pg2, int cin_pad) {
    return (size_t)S * (256ull * KP + 64ull * planes * Pg2 + 8ull * Pg) + (size_t)cin_pad * 8 + (2 * S + 1) * 8 + 16 + 128;
}

#ifdef NUNET_PROFILE_WAITS
extern "C" int nunet_debug_read_wgrad_prof(unsigned long long* out16, int reset) {
    cudaError_t e = cudaMemcpyFromSymbol(out16, g_wgrad_prof, sizeof(unsigned long long) * 16);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    if (reset) {
        unsigned long long z[16] = {0};
        e = cudaMemcpyToSymbol(g_wgrad_prof, z, sizeof(z));
    }
    return e == cudaSuccess ? 0 : NUNET_E_CUDA + (int)e;
}
#endif

extern "C" int nunet_conv3x3_wgrad_plan(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout, nunet_wgrad_plan_t* pl) {
    if (!pl || N <= 0 || H <= 0 || W <= 0 || cin_total <= 0 || Cout <= 0 || Cout % 32 != 0) return NUNET_E_BADARG;
    if (H >= 65536 || W >= 65536) return NUNET_E_UNSUPPORTED;
    const int cin_pad = cin_total <= 16 ? 16 : cin_total;
    if (cin_pad % 16 != 0) return NUNET_E_UNSUPPORTED;
    const long long Q = ((long long)N * (H + 1) + 1) * (W + 2);
    if (Q >= (1LL << 31) - 4096) return NUNET_E_UNSUPPORTED;
    const size_t budget = 220 * 1024;
    int KP = 0, S = 0, triple = 0;
    const char* env_tr = getenv("NUNET_WGRAD_TRIPLE");
    const int tr_max = env_tr ? atoi(env_tr) : 0;  // default: single gy plane (the producers, not the tensor pipe, bound wgrad)
    // prefer the three-plane layout (3 MMAs of N=96 per K step) with >= 2 stages; fall back to one plane (9 MMAs of N=32)
    for (int tr = tr_max; tr >= 0 && !KP; --tr) {
        static const int kps[] = {256, 192, 128, 96, 64};
        for (int ki = 0; ki < (tr ? 3 : 5) && !KP; ++ki) {
            const int kp = kps[ki];
            const int Pg = kp + 2 * (W + 2) + 2, Pg2 = tr ? kp + 2 * (W + 2) : Pg;
            if ((size_t)Pg * 16 >= (1u << 18)) continue;
            for (int s = 3; s >= 2; --s)
                if (wgrad_smem_bytes(kp, s, Pg, tr ? 3 : 1, Pg2, cin_pad) <= budget) { KP = kp; S = s; triple = tr; break; }
        }
    }
    if (!KP) {
        // the cp.async variant does not fit (rows of >= ~600 pixels): the layer exists on the TMA variant (column strips) only
        pl->KP = 0; pl->S = 0; pl->Pg = 0; pl->cpc = 0; pl->nchunks_total = 0; pl->splits = 0; pl->m_tiles = (cin_pad + 127) / 128;
        pl->n_tiles = Cout / 32; pl->cin_pad = cin_pad; pl->triple = 0; pl->smem_bytes = 0; pl->part_elems = 0;
        if (nunet_wgrad_tma_plan(N, H, W, cin_pad, Cout, pl) != 0) return NUNET_E_SMEM;
        pl->part_elems = (int64_t)pl->tma_splits * 9 * cin_pad * Cout;
        return NUNET_OK;
    }
    const int nchunks_total = (int)((Q + KP - 1) / KP);
    const int m_tiles = (cin_pad + 127) / 128, n_tiles = Cout / 32;
    // enough CTAs for ~2 waves, but at least 4 chunks per CTA to amortise prologue + epilogue
    int splits = (2 * 148 + m_tiles * n_tiles - 1) / (m_tiles * n_tiles);
    int max_splits = nchunks_total / 4;
    if (max_splits < 1) max_splits = 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    int cpc = (nchunks_total + splits - 1) / splits;
    splits = (nchunks_total + cpc - 1) / cpc;
    if (S > cpc) S = cpc < 1 ? 1 : cpc;
    pl->KP = KP; pl->S = S; pl->Pg = KP + 2 * (W + 2) + 2; pl->cpc = cpc; pl->nchunks_total = nchunks_total;
    pl->splits = splits; pl->m_tiles = m_tiles; pl->n_tiles = n_tiles; pl->cin_pad = cin_pad;
    pl->triple = triple;
    pl->smem_bytes = (int)wgrad_smem_bytes(KP, S, pl->Pg, triple ? 3 : 1, triple ? KP + 2 * (W + 2) : pl->Pg, cin_pad);
    pl->part_elems = (int64_t)splits * n_tiles * 9 * cin_pad * 32;
    if (nunet_wgrad_tma_plan(N, H, W, cin_pad, Cout, pl) == 0) {
        const int64_t pe = (int64_t)pl->tma_splits * 9 * cin_pad * Cout;
        if (pe > pl->part_elems) pl->part_elems = pe;
    }
    return NUNET_OK;
}

template <int MODE, bool TRIPLE>
static int launch_wgrad2(const WgradParams& p, const nunet_wgrad_plan_t& pl, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<MODE, TRIPLE>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    conv3x3_wgrad_kernel<MODE, TRIPLE><<<dim3(pl.splits, pl.m_tiles, pl.n_tiles), kWThreads, pl.smem_bytes, stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
template <int MODE>
static int launch_wgrad(const WgradParams& p, const nunet_wgrad_plan_t& pl, cudaStream_t stream) {
    return pl.triple ? launch_wgrad2<MODE, true>(p, pl, stream) : launch_wgrad2<MODE, false>(p, pl, stream);
}

extern "C" int nunet_conv3x3_wgrad(const nunet_wgrad_desc_t* d, void* stream) {
    if (!d || d->num_src < 1 || d->num_src > NUNET_MAX_SRC || !d->part || !d->dw) return NUNET_E_BADARG;
    if (!d->gy) return NUNET_E_BADARG;
    int cin = 0, cin_real = 0;
    for (int i = 0; i < d->num_src; ++i) {
        const nunet_src_t& s = d->src[i];
        if (!s.data) return NUNET_E_BADARG;
        if (s.kind == NUNET_SRC_INPUT) {
            if (s.C < 1 || s.C > 16 || d->num_src != 1) return NUNET_E_UNSUPPORTED;
            cin += 16; cin_real += s.C;
        } else {
            const bool act = (s.kind == NUNET_SRC_ACT || s.kind == NUNET_SRC_POOL_ACT || s.kind == NUNET_SRC_UP_ACT);
            if (s.kind == NUNET_SRC_GRAD || s.C % 32 != 0 || (!act && (!s.scale || !s.shift))) return NUNET_E_BADARG;
            cin += s.C; cin_real += s.C;
        }
    }
    if (d->cin_real < 0 || d->cin_real > cin_real) return NUNET_E_BADARG;
    if (d->cin_real > 0) cin_real = d->cin_real;
    nunet_wgrad_plan_t pl;
    int rc = nunet_conv3x3_wgrad_plan(d->N, d->H, d->W, cin, d->Cout, &pl);
    if (rc) return rc;
    if (pl.tma) {
        bool all_act = true;
        for (int i = 0; i < d->num_src; ++i) all_act = all_act && d->src[i].kind == NUNET_SRC_ACT;
        if (all_act) {
            const int trc = nunet_wgrad_tma_launch(d, pl, cin_real, (cudaStream_t)stream);
            if (trc >= 0) return trc;   // < 0: tensor-map encoding unavailable -> cp.async kernel below
        }
    }
    if (!pl.KP) return NUNET_E_SMEM;   // no cp.async tiling for this shape (plan: TMA variant only)
    WgradParams p{};
    p.N = d->N; p.H = d->H; p.W = d->W; p.Wp = d->W + 2; p.Hp1 = d->H + 1;
    p.Q = ((long long)d->N * (d->H + 1) + 1) * (d->W + 2);
    p.KP = pl.KP; p.S = pl.S; p.Pg = pl.Pg; p.Pg2 = pl.triple ? pl.KP + 2 * (d->W + 2) : pl.Pg; p.cpc = pl.cpc; p.nchunks_total = pl.nchunks_total;
    p.cin_pad = pl.cin_pad; p.Cout = d->Cout; p.num_src = d->num_src;
    for (int i = 0; i < d->num_src; ++i) {
        p.src[i].data = d->src[i].data; p.src[i].data2 = nullptr; p.src[i].scale = d->src[i].scale;
        p.src[i].shift = d->src[i].shift; p.src[i].k2 = nullptr; p.src[i].C = d->src[i].C; p.src[i].kind = d->src[i].kind;
    }
    p.gy = (const __nv_bfloat16*)d->gy;
    p.part = d->part;
    int kinds[NUNET_MAX_SRC];
    for (int i = 0; i < d->num_src; ++i) kinds[i] = d->src[i].kind;
    int lrc;
    switch (sources_mode(kinds, d->num_src)) {
        case MODE_INPUT: lrc = launch_wgrad<MODE_INPUT>(p, pl, (cudaStream_t)stream); break;
        case MODE_SAME: lrc = launch_wgrad<MODE_SAME>(p, pl, (cudaStream_t)stream); break;
        case MODE_POOL: lrc = launch_wgrad<MODE_POOL>(p, pl, (cudaStream_t)stream); break;
        case MODE_SAME_UP: lrc = launch_wgrad<MODE_SAME_UP>(p, pl, (cudaStream_t)stream); break;
        case MODE_ACT: lrc = launch_wgrad<MODE_ACT>(p, pl, (cudaStream_t)stream); break;
        case MODE_POOL_ACT: lrc = launch_wgrad<MODE_POOL_ACT>(p, pl, (cudaStream_t)stream); break;
        case MODE_ACT_UP: lrc = launch_wgrad<MODE_ACT_UP>(p, pl, (cudaStream_t)stream); break;
        default: return NUNET_E_UNSUPPORTED;
    }
    if (lrc) return lrc;
    cudaError_t e;
    const size_t block_elems = (size_t)pl.n_tiles * 9 * pl.cin_pad * 32;
    int blocks = (int)((block_elems + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    wgrad_reduce_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d->part, pl.splits, d->Cout, pl.cin_pad, cin_real, d->dw);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
