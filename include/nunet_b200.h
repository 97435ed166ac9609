/* nunet_b200.h — C ABI of the B200-native UNet++ hot path (libnunet_b200.so).
 *
 * The reference (husheng876/pytorch_nested-unet) has no FFI/plugin interface: its boundary is the Python module
 * surface (archs.NestedUNet / archs.UNet, losses.BCEDiceLoss, metrics.iou_score).  The entry
 * points below are what that surface binds to in this build; each comment cites the reference lines whose GPU
 * work the entry point replaces.  Conventions (all functions):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer owned by the caller unless stated otherwise
 *   - `stream` is a cudaStream_t passed as void*; nothing here synchronises, allocates or frees device memory
 *   - return 0 on success; NUNET_E_* (>0) for an invalid descriptor, 1000+cudaError_t for a CUDA launch error
 *   - re-entrant across streams/devices (no mutable global state)
 */
#ifndef NUNET_B200_H
#define NUNET_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define NUNET_ABI_VERSION 2
#define NUNET_MAX_SRC 5

enum { NUNET_OK = 0, NUNET_E_BADARG = 1, NUNET_E_UNSUPPORTED = 2, NUNET_E_SMEM = 3, NUNET_E_CUDA = 1000 };

/* How one concatenated source enters the consuming 3x3 convolution's loader. */
enum {
    NUNET_SRC_INPUT = 0, /* network input, fp32 NCHW [N,C,H,W], C <= 16 (zero-padded to 16 channels)            */
    NUNET_SRC_SAME = 1,  /* relu(scale*y+shift) of a raw conv output y at the same resolution (bf16 NHWC)       */
    NUNET_SRC_POOL = 2,  /* 2x2 max-pool of relu(scale*y+shift), y at (2H,2W)   — finished/archs1.py:82        */
    NUNET_SRC_UP = 3,    /* bilinear 2x (align_corners=True) of relu(scale*y+shift), y at (H/2,W/2) — :83       */
    NUNET_SRC_GRAD = 4,  /* backward operand: scale*gz + k2*y + shift (BatchNorm backward folded into loader)   */
    /* The same three skip-pathway kinds reading an ALREADY ACTIVATED bf16 NHWC tensor a = relu(scale*y+shift)
     * (written once per node output by nunet_bn_apply because up to ten loaders re-read it): no BN math in the loader,
     * and the same-resolution kind is staged with cp.async (global -> shared, zero-filled padding) without registers. */
    NUNET_SRC_ACT = 5,
    NUNET_SRC_POOL_ACT = 6,
    NUNET_SRC_UP_ACT = 7
};

typedef struct {
    const void* data;   /* see kind */
    const void* data2;  /* GRAD: raw conv output y (bf16 NHWC); else NULL */
    const float* scale; /* [C] per-channel; INPUT: unused */
    const float* shift; /* [C] */
    const float* k2;    /* [C] GRAD only */
    int32_t C;          /* channels contributed (multiple of 32; INPUT: real channel count <= 16) */
    int32_t kind;
} nunet_src_t;

/* BatchNorm finalisation run as the TAIL of the kernel that produced the per-CTA partial sums (csrc/bn_tail.cuh): the last CTA to
 * finish reduces the partial rows in a fixed order (fp64, deterministic) and writes the per-channel coefficients, so no separate
 * launch sits between a convolution and its consumer.  counter == NULL disables the tail (the caller then launches
 * nunet_bn_finalize_train / nunet_bn_backward_reduce on the partials itself).
 * counter: ONE zero-initialised uint32 owned by this layer (self-resetting; launches that may overlap need distinct counters).
 * Forward (nn.BatchNorm2d train-mode statistics, finished/archs1.py:19,21): fields as nunet_bn_finalize_train.
 * gamma == NULL: only sums_out ([2][C] fp64) is written — the SyncBN exchange point. */
/* SyncBN inside the tail (csrc/bn_tail.cuh): world > 1 makes the last CTA exchange its [2][C] fp64 sums with the other ranks of the
 * node through peer-mapped memory (flags with release / acquire at system scope, peer loads over NVLink, summed in rank order) before
 * it finalises with `count` = the GLOBAL element count.  data[q]: rank q's slot of THIS layer and direction, [2 epoch parities][2][C]
 * fp64; flags[q]: rank q's 8 flag words of this slot (word r is written by rank r), zero-initialised once; seq: local epoch counter
 * of the slot (device memory, zero-initialised once).  Every pointer must be dereferenceable from this device. */
typedef struct {
    int32_t world, rank;
    double* data[8];
    uint32_t* flags[8];
    uint32_t* seq;
} nunet_bn_exchange_t;

typedef struct {
    uint32_t* counter;
    double count;
    const float *gamma, *beta, *conv_bias;
    float *running_mean, *running_var;
    float momentum, eps;
    float *scale, *shift, *save_mean, *save_invstd;
    double* sums_out;
    nunet_bn_exchange_t xch;
} nunet_bn_tail_t;
/* Backward (sums of nunet_grad_router): fields as nunet_bn_backward_reduce / nunet_bn_backward_from_sums. */
typedef struct {
    uint32_t* counter;
    double count;
    const float *gamma, *save_mean, *save_invstd;
    float *dgamma, *dbeta, *k1, *k2, *k3;
    float grad_scale;
    double* sums_out;
    nunet_bn_exchange_t xch;
} nunet_bn_bwd_tail_t;

/* One fused 3x3 convolution:  out[n,y,x,co] = sum_{src, ci, r, s} T_src(a_src)[n,y+r-1,x+s-1,ci] * W[co,ci,r,s]
 * (zero padding applied AFTER the per-source transform, as torch.cat + nn.Conv2d(padding=1) does).
 * Replaces, for one VGGBlock conv: nn.Conv2d (finished/archs1.py:18,20,24,28) and — via the source kinds — the
 * preceding nn.BatchNorm2d + ReLU apply (:25-26,29-30), nn.MaxPool2d (:82), nn.Upsample (:83) and torch.cat (:117-132).
 * With stat_part != NULL the epilogue also emits per-CTA per-channel [sum, sum of squares] of the conv output (the batch
 * statistics nn.BatchNorm2d computes in train mode): of the bf16 values it stores — what the following BatchNorm normalises —
 * in the generic kernel, of the fp32 accumulators in the stacked Cout = 32 kernel (wpack_layout 1); both are summed in fp32 per
 * thread and in fp64 across CTAs, in a fixed order.  The same entry point runs dgrad (source kind GRAD,
 * weights packed with NUNET_PACK_DGRAD). */
typedef struct {
    int32_t N, H, W; /* output resolution */
    int32_t num_src;
    nunet_src_t src[NUNET_MAX_SRC];
    int32_t Cout;      /* output channels (row pitch of `out`) */
    const void* wpack; /* from nunet_pack_weights with the plan of this descriptor */
    void* out;         /* bf16 NHWC [N,H,W,Cout] */
    float* stat_part;  /* [plan.grid_m][2][Cout] fp32 or NULL */
    int32_t wpack_layout; /* 0: standard pack (NUNET_PACK_FWD / _DGRAD); 1: stacked pack (NUNET_PACK_*_C32) for the Cout = 32 kernel,
                           * only where nunet_conv3x3_c32_eligible() says so and every source is NUNET_SRC_ACT */
    nunet_bn_tail_t bn;   /* bn.counter != NULL (needs stat_part): BatchNorm finalisation in the kernel's tail */
    /* dgrad only.  num_out_seg > 0: the Cout output channels (= the forward conv's concatenated input channels) are written to one
     * DENSE bf16 NHWC tensor per segment (out_seg[i]: [N,H,W,out_seg_c[i]], channel counts multiples of 32 summing to Cout) instead
     * of `out`: the gradient router / nunet_up_transpose then read whole cache lines instead of 64-byte slices of a wide row
     * (measured: 1.77x DRAM over-read on the interleaved layout, profiles/r02_ncu_kernels.csv). */
    int32_t num_out_seg;
    void* out_seg[NUNET_MAX_SRC];
    int32_t out_seg_c[NUNET_MAX_SRC];
    /* dgrad of a VGGBlock's SECOND conv only: fuses the block-internal gradient router (autograd of finished/archs1.py:25-26, the
     * first conv's BatchNorm + ReLU) into the epilogue.  mask_y: the first conv's raw output y1 (bf16 NHWC, Cout channels);
     * out = dX * [mask_scale*y1 + mask_shift > 0] (= g_z of the first conv); stat_part rows receive [sum g_z, sum g_z*yhat] with
     * yhat = (y1 - mask_mean)*mask_invstd; bn_bwd.counter != NULL: BatchNorm-backward coefficients in the kernel's tail. */
    const void* mask_y;
    const float *mask_scale, *mask_shift, *mask_mean, *mask_invstd;
    nunet_bn_bwd_tail_t bn_bwd;
} nunet_conv_desc_t;

typedef struct {
    int32_t T, NT, CK, S, P; /* M-tiles per CTA, N tile, channels per K chunk, pipeline stages, staged positions */
    int32_t nchunks, n_tiles, grid_m, cin_pad; /* grid_m = CTAs along x = rows of stat_part (persistent kernel) */
    int32_t smem_bytes, tmem_cols, ntg;       /* ntg = tile groups of 128*T positions */
    int32_t tma_a_stage_bytes, tma_smem_bytes; /* > 0: the TMA tensor-load variant of this tiling fits (copy-mode sources) */
    int32_t tma_epi_warps;                     /* warps draining TMEM in the TMA variant (16 or 8) */
    int32_t strips;                            /* > 1: rows wider than a tensor-load box (W + 2 > 256): the plan describes the batch as
                                                  N*strips images of ceil(W/strips) columns whose halo columns are the neighbouring
                                                  strips' pixels (the box simply starts one pixel to the left of the strip); tensor-load
                                                  sources only — the transforming loader kinds return NUNET_E_UNSUPPORTED there */
    int64_t wpack_elems; /* bf16 elements the packed weights occupy */
} nunet_conv_plan_t;

int nunet_version(void);
/* Tiling decision for a conv of this shape (pure host function, no CUDA calls). */
int nunet_conv3x3_plan(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout, nunet_conv_plan_t* plan);
int nunet_conv3x3_fwd(const nunet_conv_desc_t* d, void* stream);
/* 1 if a conv of this shape over NUNET_SRC_ACT sources can run on the stacked Cout = 32 kernel (csrc/conv_c32.cu): kernel columns
 * stacked into N = 96 on the weight side, column shift applied to the accumulators.  Weights then use the *_C32 pack modes
 * ([chunk][kernel row][k-group][kernel column * 32 + n][8]; same size as the standard pack). */
int nunet_conv3x3_c32_eligible(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout);

enum { NUNET_PACK_FWD = 0, NUNET_PACK_DGRAD = 1, NUNET_PACK_FWD_C32 = 2, NUNET_PACK_DGRAD_C32 = 3 };
/* fp32 OIHW master weights [Cout,Cin,3,3] -> bf16 tile images consumed by cp.async.bulk (one contiguous block per
 * (N tile, K chunk)).  FWD: GEMM-N = Cout, GEMM-K = Cin (padded to cin_pad).  DGRAD: GEMM-N = Cin, GEMM-K = Cout,
 * taps flipped (transposed convolution).  `plan` is the plan of the conv that will consume the pack. */
int nunet_pack_weights(const float* w_oihw, int32_t Cout, int32_t Cin, int32_t mode, const nunet_conv_plan_t* plan,
                       void* wpack, void* stream);
/* Same for up to any number of layers in one launch per NUNET_MAX_PACK_JOBS jobs (jobs: host array; NT/CK/nchunks/n_tiles from
 * the layer's nunet_conv_plan_t, out: plan.wpack_elems bf16). */
#define NUNET_MAX_PACK_JOBS 64
typedef struct {
    const float* w;
    void* out;
    int32_t Cout, Cin, mode, NT, CK, nchunks, n_tiles, pad_;
} nunet_pack_job_t;
int nunet_pack_weights_batched(const nunet_pack_job_t* jobs, int32_t njobs, void* stream);

/* Reduce the per-CTA partial statistics and emit the BN apply coefficients (nn.BatchNorm2d train-mode forward,
 * finished/archs1.py:19,21): mean/var over count = N*H*W, scale = gamma*rsqrt(var+eps), shift = beta - mean*scale,
 * running_mean/var momentum update (unbiased var), conv bias folded into running_mean (the bias itself cancels in
 * train-mode BN and is never added to y).  save_mean / save_invstd are kept for the backward pass.
 * scratch: nunet_bn_workspace_bytes(C) bytes; counters: C/32 uint32, zero-initialised once (self-resetting).
 * sums_out (optional): [2][C] fp64 [sum, sum of squares] — with gamma == NULL only the sums are produced, which is
 * the SyncBN path: all-reduce sums_out across ranks, then nunet_bn_finalize_from_sums with the GLOBAL count. */
int nunet_bn_workspace_bytes(int32_t C);
int nunet_bn_finalize_train(const float* stat_part, int32_t rows, int32_t C, double count, const float* gamma,
                            const float* beta, const float* conv_bias, float* running_mean, float* running_var,
                            float momentum, float eps, float* scale, float* shift, float* save_mean,
                            float* save_invstd, double* scratch, uint32_t* counters, double* sums_out, void* stream);
int nunet_bn_finalize_from_sums(const double* sums, int32_t C, double count, const float* gamma, const float* beta,
                                const float* conv_bias, float* running_mean, float* running_var, float momentum,
                                float eps, float* scale, float* shift, float* save_mean, float* save_invstd,
                                void* stream);
/* a = relu(scale*y + shift), bf16 NHWC -> bf16 NHWC (nn.BatchNorm2d apply + ReLU of a node output that several later
 * convolutions consume, finished/archs1.py:29-30). npix = N*H*W, C % 8 == 0. */
int nunet_bn_apply(const void* y, const float* scale, const float* shift, long long npix, int32_t C, void* a, void* stream);
/* out[N,2h,2w,C] = bilinear 2x (align_corners=True) of an activated bf16 NHWC tensor (nn.Upsample, finished/archs1.py:83),
 * for callers that prefer one materialised operand over the NUNET_SRC_UP* loaders. */
int nunet_upsample2x(const void* in, int32_t N, int32_t h, int32_t w, int32_t C, void* out, void* stream);
/* "Activate" (csrc/activate.cu): everything the consumers of a raw conv output need in ONE pass over y — a = relu(scale*y+shift)
 * (finished/archs1.py:25-26,29-30), its 2x2 max-pool (:82), its bilinear 2x upsample (:83) and the 1x1 head's logits (:106-111,135-142,
 * computed from the fp32 activation; C == 32, K <= 8, fp32 NCHW).  Any output pointer may be NULL (at least one must be given);
 * replaces the separate nunet_bn_apply / nunet_maxpool2x2 / nunet_upsample2x / nunet_head_fwd launches. */
typedef struct nunet_activate_desc_s {
    const void* y; /* bf16 NHWC [N,H,W,C] */
    int32_t N, H, W, C;
    const float *scale, *shift;
    void *a, *a_pool, *a_up; /* bf16 NHWC [N,H,W,C], [N,H/2,W/2,C], [N,2H,2W,C] */
    const float *wh, *bh;    /* head: [K][32], [K] */
    int32_t K, pad_;
    float* logits;           /* [N,K,H,W] */
} nunet_activate_desc_t;
int nunet_bn_activate(const nunet_activate_desc_t* d, void* stream);
/* Eval mode: scale = gamma*rsqrt(running_var+eps), shift = beta + (conv_bias - running_mean)*scale. */
int nunet_bn_fold_eval(int32_t C, const float* gamma, const float* beta, const float* conv_bias,
                       const float* running_mean, const float* running_var, float eps, float* scale, float* shift,
                       void* stream);

/* ---- 1x1 heads, loss, metric (finished/archs1.py:106-111,135-142; losses.py:107-117; metrics.py:6-18) ---------------
 * head_fwd: logits[n,k,y,x] (fp32 NCHW) = sum_c relu(scale*y+shift)[n,y,x,c]*wh[k,c] + bh[k]; y bf16 NHWC with 32 ch. */
int nunet_head_fwd(const void* y, const float* scale, const float* shift, const float* wh, const float* bh, int32_t N,
                   int32_t H, int32_t W, int32_t K, float* logits, void* stream);
/* One pass over (logits, target) [N][E = K*H*W] fp32: BCE-with-logits + per-sample dice sums + integer IoU counts.
 * part: [N][bps][4] fp32, cnt: [N][bps][2] int32 with bps = nunet_loss_blocks_per_sample(N,E);
 * sums: [N][3] fp64 (sum p*t, sum p, sum t — kept for loss_bwd); loss: fp32 scalar (0.5*bce + 1 - mean dice);
 * counts: [2] int64 (intersection, union over the whole batch and all classes). loss / counts may be NULL. */
int nunet_loss_blocks_per_sample(int32_t N, int32_t E);
int nunet_loss_iou_fwd(const float* logits, const float* target, int32_t N, int32_t E, float* part, int32_t* cnt,
                       double* sums, float* loss, long long* counts, void* stream);
/* metrics.iou_score counts only: counts[0] = |o & t|, counts[1] = |o | t| with o = sigmoid(logit) > 0.5, t = target > 0.5. */
int nunet_iou_counts(const float* logits, const float* target, long long total, unsigned long long* counts, void* stream);
/* val.py:100-105 output stage: out[i] = (uint8)(sigmoid(logits[i]) * 255) (numpy astype('uint8') truncation), same NCHW order. */
int nunet_sigmoid_u8(const float* logits, long long total, uint8_t* out, void* stream);
/* dlogits = (gscale ? *gscale : 1) * gmul * d(loss)/d(logits). */
int nunet_loss_bwd(const float* logits, const float* target, const double* sums, const float* gscale, float gmul, int32_t N,
                   int32_t E, float* dlogits, void* stream);
/* The deep-supervision loop (trains.py:120-123: `for output in outputs: loss += criterion(output, target)`) in ONE pass: nheads <= 4
 * logit tensors of the same shape against the same target.  logits / gscale / dlogits are HOST arrays of nheads device pointers.
 * part: [nheads][N][bps][4], cnt: [nheads][N][bps][2], sums: [nheads][N][3], loss: [nheads], counts: [nheads][2]; every head's
 * results are bitwise those of nunet_loss_iou_fwd / nunet_loss_bwd on that head alone.  bwd: gscale[k] NULL = 1, dlogits[k] NULL =
 * head k's loss received no gradient (nothing is written for it). */
int nunet_loss_iou_fwd_multi(const float* const* logits, int32_t nheads, const float* target, int32_t N, int32_t E, float* part,
                             int32_t* cnt, double* sums, float* loss, long long* counts, void* stream);
int nunet_loss_bwd_multi(const float* const* logits, int32_t nheads, const float* target, const double* sums,
                         const float* const* gscale, float gmul, int32_t N, int32_t E, float* const* dlogits, void* stream);
/* LovaszHingeLoss (losses.py:49-96,120-129; csrc/lovasz.cu): logits / target [N][P] fp32 (P = C*H*W of one image, target in {0,1});
 * per image a descending sort of the hinge errors, the Jaccard gradient of the sorted labels, loss = mean_n dot(relu(e_sorted), grad).
 * coef [N][P]: d loss_image / d logit (kept for the backward pass); loss_img [N] fp64; loss: fp32 scalar.  scratch:
 * nunet_lovasz_scratch_elems(N, P) uint64 (0 when an image's pairs fit shared memory: P <= 16384).
 * bwd: dlogits = coef * (gscale ? *gscale : 1) * gmul   (gmul = 1/N). */
long long nunet_lovasz_scratch_elems(int32_t N, int32_t P);
int nunet_lovasz_hinge_fwd(const float* logits, const float* target, int32_t N, int32_t P, unsigned long long* scratch, float* coef,
                           double* loss_img, float* loss, void* stream);
int nunet_lovasz_hinge_bwd(const float* coef, const float* gscale, float gmul, long long total, float* dlogits, void* stream);
/* Head parameter gradients: dwh[k][c] = sum dlogit*a, dbh[k] = sum dlogit. part: [nunet_head_bwd_params_rows()][K][33]. */
int nunet_head_bwd_params_rows(void);
int nunet_head_bwd_params(const void* y, const float* scale, const float* shift, const float* dlogits, int32_t N, int32_t H,
                          int32_t W, int32_t K, float* part, float* dwh, float* dbh, int32_t accumulate, void* stream);

/* ---- backward of the skip pathway + BatchNorm/ReLU (autograd of finished/archs1.py:23-32,82-83,117-142) ---------------
 * One consumer of this conv output's activation: the consumer's dgrad buffer (bf16 NHWC at the consumer's resolution,
 * `pitch` channels per pixel) holds our slice at channel offset `coff`; kind says how the consumer read us. */
typedef struct {
    const void* dx;
    int32_t pitch, coff, kind; /* NUNET_SRC_SAME | NUNET_SRC_POOL | NUNET_SRC_UP */
} nunet_grad_cons_t;

typedef struct {
    int32_t N, H, W, C; /* this conv output */
    const void* y;      /* raw conv output, bf16 NHWC */
    const float *scale, *shift, *mean, *invstd;
    int32_t num_cons;
    nunet_grad_cons_t cons[6];
    const float* dlogits; /* optional head consumer: fp32 NCHW [N,K,H,W] */
    const float* wh;      /* [K][C] */
    int32_t K;
    void* gz;    /* out: g_z = g_a * [a>0], bf16 NHWC */
    float* part; /* out: [nunet_grad_router_rows()][2][C]: sum g_z, sum g_z*yhat */
    nunet_bn_bwd_tail_t bn; /* bn.counter != NULL: BatchNorm-backward coefficients in the kernel's tail */
    /* head_part != NULL (needs dlogits, K == 1, C == 32, bn.counter, no POOL / UP consumer): the head's own parameter gradients
     * dwh[c] = sum dlogit*a[c] ([32]) and dbh = sum dlogit ([1], may be NULL) are accumulated in the same pass and finished by the
     * last CTA (replaces nunet_head_bwd_params).  head_part: [nunet_grad_router_rows()][33] fp32 scratch. */
    float* head_part;
    float *dwh, *dbh;
} nunet_router_desc_t;
int nunet_grad_router_rows(int32_t N, int32_t H, int32_t W, int32_t C);
int nunet_grad_router(const nunet_router_desc_t* d, void* stream);
/* Transposed align_corners bilinear 2x upsample: out[N,h,w,C] (bf16 NHWC, dense) = backward of the UP loader applied to
 * the channel slice [coff, coff+C) of a consumer's dgrad buffer dx[N,2h,2w,pitch]; the router then reads `out` as a
 * NUNET_SRC_SAME consumer (pitch C, offset 0). */
int nunet_up_transpose(const void* dx, int32_t pitch, int32_t coff, int32_t N, int32_t h, int32_t w, int32_t C, void* out,
                       void* stream);
/* Finishes BatchNorm backward: dbeta = S1, dgamma = S2 and the loader coefficients of g_y = k1*g_z + k2*y + k3.
 * gamma == NULL: only sums_out ([2][C] fp64) is produced (SyncBN: all-reduce, then nunet_bn_backward_from_sums). */
int nunet_bn_backward_reduce(const float* part, int32_t rows, int32_t C, double count, const float* gamma,
                             const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* k1,
                             float* k2, float* k3, double* scratch, uint32_t* counters, double* sums_out, void* stream);
int nunet_bn_backward_from_sums(const double* sums, int32_t C, double count, const float* gamma, const float* save_mean,
                                const float* save_invstd, float* dgamma, float* dbeta, float* k1, float* k2, float* k3,
                                float grad_scale, void* stream);

/* g_y = k1*g_z + k2*y + k3 (bf16 NHWC, may alias gz): the affine rest of BatchNorm backward, materialised once because
 * both dgrad (as a pre-activated NUNET_SRC_ACT source) and every wgrad CTA re-read it. */
int nunet_bn_bwd_apply(const void* gz, const void* y, const float* k1, const float* k2, const float* k3, long long npix,
                       int32_t C, void* gy, void* stream);

/* ---- weight gradient (autograd of nn.Conv2d, finished/archs1.py:18,20) --------------------------------------------------
 * dw[co][ci][3][3] (fp32 OIHW, overwritten) = sum over pixels of g_y (x) X, with X read through the same fused loaders
 * as the forward conv (src[]).  part: plan.part_elems fp32 of scratch. */
typedef struct {
    int32_t KP, S, Pg, cpc, nchunks_total, splits, m_tiles, n_tiles, cin_pad, smem_bytes, triple;
    /* TMA variant (all sources NUNET_SRC_ACT, rows <= 254 px; csrc/conv_wgrad_tma.cu).  tma = 1: applies.  tma_PW: padded row
     * length; tma_r > 0: padded rows of one image per K chunk, tma_r < 0: -(whole images per chunk); tma_S: pipeline stages;
     * tma_G: MMAs per 16-position K step (3: kernel columns stacked into N = 96; 1: Cin = 32, kernel rows stacked on A too);
     * tma_Nt: output channels per CTA (32); tma_splits x tma_cpc >= tma_chunks: split-K partition (one wave of CTAs). */
    int32_t tma, tma_PW, tma_r, tma_S, tma_G, tma_Nt, tma_splits, tma_cpc, tma_chunks, tma_smem_bytes;
    int64_t part_elems;
} nunet_wgrad_plan_t;
typedef struct {
    int32_t N, H, W;
    int32_t num_src;
    nunet_src_t src[NUNET_MAX_SRC];
    int32_t Cout;
    const void* gy; /* bf16 NHWC [N,H,W,Cout]: g_y = k1*g_z + k2*y + k3 from nunet_bn_bwd_apply */
    float* part;
    float* dw;
    int32_t cin_real; /* 0: sum of the sources' channels; else the leading channels that exist in dw (zero-padded stem input) */
} nunet_wgrad_desc_t;
/* 2x2/stride-2 max-pool of an activated bf16 NHWC tensor (nn.MaxPool2d(2,2), finished/archs1.py:80 applied to a node output):
 * out [N,H/2,W/2,C]. */
int nunet_maxpool2x2(const void* in, int32_t N, int32_t H, int32_t W, int32_t C, void* out, void* stream);
/* fp32 NCHW network input [N,C,H,W], C <= 32 -> bf16 NHWC [N,H,W,32] with zero channels >= C (stem conv operand). */
int nunet_input_to_nhwc32(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, void* out, void* stream);
/* Device-side input pipeline (csrc/input.cu): uint8 NHWC image -> ((u8/max_pixel - mean)/std)/post_div (albumentations Normalize(),
 * trains.py:266, then dataset.py:71's /255; IEEE fp32 divisions in that order) as out_kind 0: bf16 NHWC with 32 zero-padded channels
 * (the stem conv's operand) or 1: fp32 NCHW.  mean == NULL (out_kind 1 only): value/max_pixel — the mask path of dataset.py:73-74.
 * aug: optional [N] uint8 per-image geometric augmentation, bits 0-1 = np.rot90 quarter turns (trains.py:257, needs H == W), bit 2 /
 * bit 3 = horizontal / vertical flip (trains.py:258), applied after the turns. */
int nunet_input_u8(const uint8_t* img, int32_t N, int32_t H, int32_t W, int32_t C, const float* mean, const float* stdv, float max_pixel,
                   float post_div, const uint8_t* aug, void* out, int32_t out_kind, void* stream);
int nunet_conv3x3_wgrad_plan(int32_t N, int32_t H, int32_t W, int32_t cin_total, int32_t Cout, nunet_wgrad_plan_t* plan);
int nunet_conv3x3_wgrad(const nunet_wgrad_desc_t* d, void* stream);

/* ---- fused optimizer step (trains.py:131-133 with the optimizers of trains.py:225-233; csrc/optim.cu) ---------------------------------
 * One launch updates every parameter tensor of the model with torch.optim's arithmetic (SGD: momentum, dampening, weight decay,
 * nesterov; Adam: betas, eps, weight decay).  tensors_dev: DEVICE array, one entry per parameter (grad == NULL: skipped); chunks_dev:
 * DEVICE array of int32 pairs (tensor index, chunk index) — one CTA per nunet_optim_chunk_elems() elements; hyper_dev: DEVICE fp32
 * [lr, momentum, dampening, weight_decay, nesterov] (SGD) or [lr, beta1, beta2, eps, weight_decay] (Adam), so a captured CUDA graph
 * follows the host's learning-rate schedule; step_dev (Adam): DEVICE int32 [ntensors] step counters (torch keeps one per parameter),
 * advanced by the call for every tensor that had a gradient. */
enum { NUNET_OPTIM_SGD = 0, NUNET_OPTIM_ADAM = 1 };
typedef struct {
    float* param;       /* fp32 master, updated in place */
    const float* grad;  /* fp32 gradient or NULL */
    float* state1;      /* SGD: momentum buffer (zero-initialised; NULL if momentum == 0) ; Adam: exp_avg */
    float* state2;      /* Adam: exp_avg_sq */
    int64_t numel;      /* bit 62 set (SGD): this is the buffer's first step: buf = g' (torch clones the gradient; dampening is not applied) */
} nunet_optim_tensor_t;
#define NUNET_OPTIM_FIRST_STEP (1LL << 62)
int nunet_optim_chunk_elems(void);
int nunet_optim_step(const nunet_optim_tensor_t* tensors_dev, int32_t ntensors, const int32_t* chunks_dev, int32_t nchunks,
                     const float* hyper_dev, int32_t algo, int32_t* step_dev, void* stream);

/* ---- fp32 check mode (forward and backward; csrc/check_fp32.cu) -----------------------------------------------------
 * The reference's fp32 arithmetic on CUDA cores: fp32 NHWC activations with a channel pitch, fp64 BatchNorm statistics.
 * Used to hold the graph / BatchNorm / pooling / upsampling / head semantics to <= 1e-4 against the reference's logits and
 * <= 1e-3 (rel-L2) against the reference's gradients. */
int nunet_check_nchw_to_nhwc(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, float* out, void* stream);
/* nn.Conv2d(Cin, Cout, 3, padding=1) incl. bias (finished/archs1.py:18,20); wt_scratch: 9*Cin*Cout floats */
int nunet_check_conv3x3(const float* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t xpitch, const float* w_oihw,
                        const float* bias, int32_t Cout, float* wt_scratch, float* y, void* stream);
/* nn.BatchNorm2d (finished/archs1.py:19,21): train != 0: batch statistics (+ running-stat update), else running statistics;
 * emits the per-channel scale/shift of a = relu(scale*y + shift) */
int nunet_check_bn(const float* y, long long M, int32_t C, const float* gamma, const float* beta, float* running_mean,
                   float* running_var, float momentum, float eps, int32_t train, float* scale, float* shift, float* save_mean,
                   float* save_invstd, void* stream);
int nunet_check_bn_relu(const float* y, long long M, int32_t C, const float* scale, const float* shift, float* dst, int32_t pitch,
                        int32_t off, void* stream);
/* mode 0 copy | 1 nn.MaxPool2d(2,2) | 2 nn.Upsample(2, bilinear, align_corners=True) of src [N,h,w,C] into the channel
 * slice [off, off+C) of dst (torch.cat, finished/archs1.py:117-132) */
int nunet_check_route(const float* src, int32_t N, int32_t h, int32_t w, int32_t C, int32_t mode, float* dst, int32_t pitch,
                      int32_t off, void* stream);
int nunet_check_head(const float* a, int32_t N, int32_t H, int32_t W, int32_t C, const float* wh, const float* bh, int32_t K,
                     float* logits, void* stream);
/* Backward of the check-mode graph (autograd of finished/archs1.py:23-32,82-83,117-142 in fp32 with fp64 reductions).
 * relu_bn_bwd: g_z = g_a*[scale*y+shift > 0]; dbeta = sum g_z; dgamma = sum g_z*yhat; g_y = gamma*invstd*(g_z - dbeta/M - yhat*dgamma/M).
 * conv3x3_dgrad: dx [N,H,W,Cin] dense (overwritten); conv3x3_wgrad: dw OIHW (overwritten), dbias [Cout] optional (sum of g_y).
 * route_bwd: ga [N,h,w,C] (dense, ACCUMULATED) += transpose of nunet_check_route applied to the slice [off, off+C) of dx; a_src is the
 * routed activation (needed for mode 1: gradient goes to the first maximum of each 2x2 window in scan order, as ATen).
 * head_bwd: ga += dlogits * wh (ACCUMULATED); dwh [K][C], dbh [K] (optional) overwritten. */
int nunet_check_relu_bn_bwd(const float* ga, const float* y, long long M, int32_t C, const float* scale, const float* shift,
                            const float* gamma, const float* save_mean, const float* save_invstd, float* dgamma, float* dbeta, float* gy,
                            void* stream);
int nunet_check_conv3x3_dgrad(const float* gy, int32_t N, int32_t H, int32_t W, int32_t Cout, const float* w_oihw, int32_t Cin,
                              float* wt_scratch, float* dx, void* stream);
int nunet_check_conv3x3_wgrad(const float* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t xpitch, const float* gy, int32_t Cout,
                              float* dw_oihw, float* dbias, void* stream);
int nunet_check_route_bwd(const float* dx, int32_t N, int32_t h, int32_t w, int32_t C, int32_t mode, int32_t pitch, int32_t off,
                          const float* a_src, float* ga, void* stream);
int nunet_check_head_bwd(const float* dlogits, const float* a, int32_t N, int32_t H, int32_t W, int32_t C, const float* wh, int32_t K,
                         float* ga, float* dwh, float* dbh, void* stream);

#ifdef __cplusplus
}
#endif
#endif
