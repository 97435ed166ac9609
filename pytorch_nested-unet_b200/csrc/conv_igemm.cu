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

namespace nunet {

// ------------------------------------------------------------------------------------------------ kernel
constexpr int kProducerWarps = 8;
constexpr int kThreads = 320;

__global__ void __launch_bounds__(kThreads, 1) conv3x3_igemm_kernel(const ConvParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cgs = p.CK >> 3;
    const uint32_t a_bytes = (uint32_t)cgs * p.P * 16u;
    const uint32_t b_bytes = 9u * p.CK * p.NT * 2u;

    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)p.S * a_bytes;
    PixInfo* pixtab = reinterpret_cast<PixInfo*>(sB + (size_t)p.S * b_bytes);
    float* s_scale = reinterpret_cast<float*>(pixtab + p.P);
    float* s_shift = s_scale + p.cin_pad;
    float* s_k2 = s_shift + p.cin_pad;
    float* s_stat = s_k2 + p.cin_pad;  // [8 warps][2][NT]: per-warp partials, summed in a fixed order (deterministic)
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_stat + 16 * p.NT);
    uint64_t* full = bars;
    uint64_t* empty = bars + p.S;
    uint64_t* accum = bars + 2 * p.S;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * p.S + 1);

    const int q0 = blockIdx.x * (128 * p.T);
    const int qbase = q0 - p.Wp - 1;
    const int n_tile = blockIdx.y;

    // ---- one-time setup
    if (warp == 8) {
        if (lane == 0) {
            for (int s = 0; s < p.S; ++s) {
                mbar_init(&full[s], kProducerWarps + 1);
                mbar_init(&empty[s], 1);
            }
            mbar_init(accum, 1);
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
    for (int i = threadIdx.x; i < p.P; i += kThreads) {
        const PixInfo pi = decode_position((long long)qbase + i, p.N, p.H, p.W, p.Wp, p.Hp1, p.Q);
        pixtab[i] = pi;
    }
    {
        int cbase = 0;
        for (int si = 0; si < p.num_src; ++si) {
            const ConvSrc& s = p.src[si];
            const int cpad = (s.kind == NUNET_SRC_INPUT) ? 16 : s.C;
            for (int c = threadIdx.x; c < cpad; c += kThreads) {
                const bool has = (s.kind != NUNET_SRC_INPUT);
                s_scale[cbase + c] = has ? s.scale[c] : 1.f;
                s_shift[cbase + c] = has ? s.shift[c] : 0.f;
                s_k2[cbase + c] = (s.kind == NUNET_SRC_GRAD) ? s.k2[c] : 0.f;
            }
            cbase += cpad;
        }
        for (int i = threadIdx.x; i < 16 * p.NT; i += kThreads) s_stat[i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < kProducerWarps) {
        // =============================== producers: stage the transformed input halo, chunk by chunk
        const int ppi = 8 * (4 / cgs);  // positions per warp iteration: 8 (CK=32) or 16 (CK=16)
        const int nblk = (p.P + ppi - 1) / ppi;
        const int pl = lane & 7, r = lane >> 3;
        const int cg = r % cgs, psub = pl + 8 * (r / cgs);
        int si = 0, src_c0 = 0;  // current source and its first concatenated channel
        for (int chunk = 0; chunk < p.nchunks; ++chunk) {
            const int s = chunk % p.S;
            const uint32_t ph = (chunk / p.S) & 1;
            const int cglob = chunk * p.CK;  // first concatenated channel of this chunk
            while (true) {
                const int cpad = (p.src[si].kind == NUNET_SRC_INPUT) ? 16 : p.src[si].C;
                if (cglob < src_c0 + cpad) break;
                src_c0 += cpad;
                ++si;
            }
            const ConvSrc& src = p.src[si];
            const int ch = cglob - src_c0 + cg * 8;  // channel inside the source
            float sc[8], sh[8], k2[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                sc[i] = s_scale[cglob + cg * 8 + i];
                sh[i] = s_shift[cglob + cg * 8 + i];
                k2[i] = s_k2[cglob + cg * 8 + i];
            }
            mbar_wait(&empty[s], ph ^ 1);
            uint8_t* dstA = sA + (size_t)s * a_bytes + (size_t)cg * p.P * 16;
            for (int blk = warp; blk < nblk; blk += kProducerWarps) {
                const int pp = blk * ppi + psub;
                if (pp < p.P) {
                    const PixInfo pi = pixtab[pp];
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (pi.pix >= 0) v = load_transform(src, p.H, p.W, pi, ch, sc, sh, k2);
                    *reinterpret_cast<uint4*>(dstA + (size_t)pp * 16) = v;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }

        // =============================== epilogue
        mbar_wait(accum, 0);
        tc_fence_after();
        const int lq = warp & 3, half = warp >> 2;
        const int nslab = p.NT >> 5;
        __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(p.out);
        for (int cs = 0; cs < nslab; ++cs) {
            float ssum[32], ssq[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
            bool any = false;
            for (int t = 0; t < p.T; ++t) {
                if (((cs * p.T + t) & 1) != half) continue;
                any = true;
                const int m = lq * 32 + lane;
                const PixInfo pi = pixtab[t * 128 + m + p.Wp + 1];
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)(t * p.NT + cs * 32), v);
                tmem_ld_wait();
                if (pi.pix >= 0) {
                    uint4* dst = reinterpret_cast<uint4*>(out + (size_t)pi.pix * p.Cout + n_tile * p.NT + cs * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 o;
                        o.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                        o.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                        o.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                        o.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                        dst[j] = o;
                    }
                    if (p.stat_part != nullptr) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const float f = __uint_as_float(v[i]);
                            ssum[i] += f;
                            ssq[i] = fmaf(f, f, ssq[i]);
                        }
                    }
                }
            }
            if (p.stat_part != nullptr && any) {
                // transposing butterfly: afterwards lane L holds the warp total of channel L
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const bool up = (lane & off) != 0;
#pragma unroll
                    for (int i = 0; i < off; ++i) {
                        const float sa = up ? ssum[i] : ssum[i + off];
                        const float ka = up ? ssum[i + off] : ssum[i];
                        ssum[i] = ka + __shfl_xor_sync(0xffffffffu, sa, off);
                        const float sb = up ? ssq[i] : ssq[i + off];
                        const float kb = up ? ssq[i + off] : ssq[i];
                        ssq[i] = kb + __shfl_xor_sync(0xffffffffu, sb, off);
                    }
                }
                s_stat[(warp * 2 + 0) * p.NT + cs * 32 + lane] = ssum[0];  // each (warp, cs) slot is written at most once
                s_stat[(warp * 2 + 1) * p.NT + cs * 32 + lane] = ssq[0];
            }
        }
        if (p.stat_part != nullptr) {
            asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 epilogue warps only
            float* dst = p.stat_part + (size_t)blockIdx.x * 2 * p.Cout + n_tile * p.NT;
            for (int c = threadIdx.x; c < p.NT; c += 256) {
                float a = 0.f, b = 0.f;
#pragma unroll
                for (int w8 = 0; w8 < 8; ++w8) {
                    a += s_stat[(w8 * 2 + 0) * p.NT + c];
                    b += s_stat[(w8 * 2 + 1) * p.NT + c];
                }
                dst[c] = a;
                dst[p.Cout + c] = b;
            }
        }
    } else if (warp == 8) {
        // =============================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.NT, kMajorK, kMajorK);
            const uint32_t a_lbo = (uint32_t)p.P * 16u, b_lbo = (uint32_t)p.NT * 16u;
            const int ksteps = p.CK >> 4;
            for (int chunk = 0; chunk < p.nchunks; ++chunk) {
                const int s = chunk % p.S;
                const uint32_t ph = (chunk / p.S) & 1;
                mbar_wait(&full[s], ph);
                tc_fence_after();
                const uint32_t a_base = smem_u32(sA + (size_t)s * a_bytes);
                const uint32_t b_base = smem_u32(sB + (size_t)s * b_bytes);
                for (int t = 0; t < p.T; ++t) {
                    const uint32_t d_addr = tmem_base + (uint32_t)(t * p.NT);
                    for (int tap = 0; tap < 9; ++tap) {
                        const int dr = tap / 3, dc = tap - 3 * dr;
                        const uint32_t a_tap = a_base + (uint32_t)(t * 128 + dr * p.Wp + dc) * 16u;
                        const uint32_t b_tap = b_base + (uint32_t)(tap * cgs) * b_lbo;
                        for (int ks = 0; ks < ksteps; ++ks) {
                            const uint64_t ad = umma_desc_nosw(a_tap + (uint32_t)(2 * ks) * a_lbo, a_lbo, 128u);
                            const uint64_t bd = umma_desc_nosw(b_tap + (uint32_t)(2 * ks) * b_lbo, b_lbo, 128u);
                            umma_bf16(d_addr, ad, bd, idesc, (chunk | tap | ks) != 0);
                        }
                    }
                }
                umma_commit(&empty[s]);
            }
            umma_commit(accum);
        }
        __syncwarp();
    } else {
        // =============================== weight fetcher (bulk async copy engine)
        if (lane == 0) {
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpack) + (size_t)n_tile * p.nchunks * b_bytes;
            for (int chunk = 0; chunk < p.nchunks; ++chunk) {
                const int s = chunk % p.S;
                const uint32_t ph = (chunk / p.S) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                mbar_arrive_expect_tx(&full[s], b_bytes);
                bulk_g2s(sB + (size_t)s * b_bytes, wsrc + (size_t)chunk * b_bytes, b_bytes, &full[s]);
            }
        }
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
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
__global__ void pack_weights_kernel(const float* __restrict__ w, int Cout, int Cin, int mode, int NT, int CK, int nchunks,
                                    int n_tiles, __nv_bfloat16* __restrict__ out) {
    const int cgs = CK >> 3;
    const size_t total = (size_t)n_tiles * nchunks * 9 * cgs * NT;  // 16-byte groups
    for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < total; g += (size_t)gridDim.x * blockDim.x) {
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

static size_t conv_smem_bytes(int T, int NT, int CK, int S, int P, int cin_pad) {
    size_t a = (size_t)(CK / 8) * P * 16, b = (size_t)9 * CK * NT * 2;
    return S * (a + b) + (size_t)P * 8 + (size_t)cin_pad * 12 + (size_t)16 * NT * 4 + (2 * S + 1) * 8 + 16 + 128;
}

extern "C" int nunet_conv3x3_plan(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout, nunet_conv_plan_t* pl) {
    if (!pl || N <= 0 || H <= 0 || W <= 0 || cin_total <= 0 || Cout <= 0) return NUNET_E_BADARG;
    if (Cout % 32 != 0) return NUNET_E_UNSUPPORTED;
    if (H >= 65536 || W >= 65536) return NUNET_E_UNSUPPORTED;
    const int cin_pad = (cin_total <= 16) ? 16 : cin_total;
    if (cin_pad % 16 != 0) return NUNET_E_UNSUPPORTED;
    const long long Q = ((long long)N * (H + 1) + 1) * (W + 2);
    if (Q + 4LL * 128 + 2 * (W + 2) + 2 >= (1LL << 31) || (long long)N * H * W * (Cout > cin_pad ? Cout : cin_pad) >= (1LL << 40))
        return NUNET_E_UNSUPPORTED;
    const int NT = pick_nt(Cout);
    if (NT == 0) return NUNET_E_UNSUPPORTED;
    const int n_tiles = Cout / NT;
    const size_t budget = 220 * 1024;
    int bestT = 0, bestCK = 0, bestS = 0;
    // prefer: large T (halo amortisation) while keeping >= 2 waves of CTAs, CK=32, 2 stages
    for (int T = (512 / NT < 4 ? 512 / NT : 4); T >= 1 && !bestT; --T) {
        if (T == 3) continue;
        const long long ctas = ((Q + 128LL * T - 1) / (128LL * T)) * n_tiles;
        if (T > 1 && ctas < 2 * 148) continue;
        for (int CK = 32; CK >= 16 && !bestT; CK -= 16) {
            if (cin_pad % CK != 0) continue;
            const int nch = cin_pad / CK;
            const int P = 128 * T + 2 * (W + 2) + 2;
            for (int S = (nch >= 3 ? 3 : nch); S >= (nch >= 2 ? 2 : 1); --S) {
                if ((size_t)P * 16 >= (1u << 18)) continue;
                if (conv_smem_bytes(T, NT, CK, S, P, cin_pad) <= budget) { bestT = T; bestCK = CK; bestS = S; break; }
            }
        }
    }
    if (!bestT) return NUNET_E_SMEM;
    pl->T = bestT; pl->NT = NT; pl->CK = bestCK; pl->S = bestS;
    pl->P = 128 * bestT + 2 * (W + 2) + 2;
    pl->nchunks = cin_pad / bestCK;
    pl->n_tiles = n_tiles;
    pl->grid_m = (int)((Q + 128LL * bestT - 1) / (128LL * bestT));
    pl->cin_pad = cin_pad;
    pl->smem_bytes = (int)conv_smem_bytes(bestT, NT, bestCK, bestS, pl->P, cin_pad);
    int cols = bestT * NT, tc = 32;
    while (tc < cols) tc <<= 1;
    pl->tmem_cols = tc;
    pl->wpack_elems = (int64_t)n_tiles * pl->nchunks * 9 * bestCK * NT;
    return NUNET_OK;
}

extern "C" int nunet_conv3x3_fwd(const nunet_conv_desc_t* d, void* stream) {
    if (!d || d->num_src < 1 || d->num_src > NUNET_MAX_SRC || !d->wpack || !d->out) return NUNET_E_BADARG;
    int cin = 0;
    for (int i = 0; i < d->num_src; ++i) {
        const nunet_src_t& s = d->src[i];
        if (!s.data) return NUNET_E_BADARG;
        if (s.kind == NUNET_SRC_INPUT) {
            if (s.C < 1 || s.C > 16 || d->num_src != 1) return NUNET_E_UNSUPPORTED;
            cin += 16;
        } else {
            if (s.C % 32 != 0 || !s.scale || !s.shift) return NUNET_E_BADARG;
            if (s.kind == NUNET_SRC_GRAD && (!s.data2 || !s.k2)) return NUNET_E_BADARG;
            if ((s.kind == NUNET_SRC_UP) && ((d->H | d->W) & 1)) return NUNET_E_UNSUPPORTED;
            cin += s.C;
        }
    }
    nunet_conv_plan_t pl;
    int rc = nunet_conv3x3_plan(d->N, d->H, d->W, cin, d->Cout, &pl);
    if (rc) return rc;
    ConvParams p{};
    p.N = d->N; p.H = d->H; p.W = d->W; p.Wp = d->W + 2; p.Hp1 = d->H + 1;
    p.Q = (d->N * (d->H + 1) + 1) * (d->W + 2);
    p.T = pl.T; p.NT = pl.NT; p.CK = pl.CK; p.S = pl.S; p.P = pl.P; p.nchunks = pl.nchunks;
    p.cin_pad = pl.cin_pad; p.Cout = d->Cout; p.num_src = d->num_src; p.tmem_cols = pl.tmem_cols;
    for (int i = 0; i < d->num_src; ++i) {
        p.src[i].data = d->src[i].data; p.src[i].data2 = d->src[i].data2;
        p.src[i].scale = d->src[i].scale; p.src[i].shift = d->src[i].shift; p.src[i].k2 = d->src[i].k2;
        p.src[i].C = d->src[i].C; p.src[i].kind = d->src[i].kind;
    }
    p.wpack = d->wpack; p.out = d->out; p.stat_part = d->stat_part;
    cudaError_t e = cudaFuncSetAttribute(conv3x3_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes);
    if (e != cudaSuccess) return NUNET_E_CUDA + (int)e;
    conv3x3_igemm_kernel<<<dim3(pl.grid_m, pl.n_tiles), kThreads, pl.smem_bytes, (cudaStream_t)stream>>>(p);
    e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}

extern "C" int nunet_pack_weights(const float* w, int32_t Cout, int32_t Cin, int32_t mode, const nunet_conv_plan_t* pl,
                                  void* wpack, void* stream) {
    if (!w || !pl || !wpack || (mode != NUNET_PACK_FWD && mode != NUNET_PACK_DGRAD)) return NUNET_E_BADARG;
    const size_t groups = (size_t)pl->wpack_elems / 8;
    int blocks = (int)((groups + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    pack_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, Cin, mode, pl->NT, pl->CK, pl->nchunks, pl->n_tiles,
                                                                   reinterpret_cast<__nv_bfloat16*>(wpack));
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? NUNET_OK : NUNET_E_CUDA + (int)e;
}
