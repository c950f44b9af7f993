/* libtris_sm100.so -- C ABI of the B200-native TRIS Stage-1 hot path.
 *
 * The reference (fawnliu/TRIS) is pure Python/PyTorch and has NO native interface; its plugin
 * boundary is the nn.Module surface model/model_stage1.py:14-123.  This header is therefore the NEW C boundary
 * that the Python shim (tris_b200/) binds with ctypes; each entry cites the reference arithmetic it replaces.
 *
 * Conventions: extern "C"; raw device pointers + sizes only (no torch types); caller owns every buffer; kernels
 * never allocate; every call is asynchronous on `stream`; return 0 on success, <0 = TRIS_ERR_*; the message of the
 * last failure on the calling thread is available from tris_last_error().  bf16 buffers are uint16 storage.
 */
#ifndef TRIS_SM100_H
#define TRIS_SM100_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* tris_stream_t; /* cudaStream_t */

#define TRIS_OK 0
#define TRIS_ERR_SHAPE (-1)
#define TRIS_ERR_ALIGN (-2)
#define TRIS_ERR_ARCH (-3)
#define TRIS_ERR_CUDA (-4)

const char* tris_last_error(void);
int tris_abi_version(void);
/* 0 if the current device is sm_100 (B200) and the driver exposes cuTensorMapEncodeTiled. */
int tris_check_device(void);

/* ------------------------------------------------------------------------------------------------------------
 * tcgen05 GEMM / implicit-GEMM convolution (bf16 operands, fp32 accumulation in TMEM, TMA-fed, persistent).
 * Replaces: nn.Conv2d 1x1/3x3 of CLIP/clip/model.py:17-40 (fwd, dgrad, wgrad), nn.Linear / in_proj / out_proj /
 * mlp of model.py:366-378, vis_project / lan_project (model_stage1.py:36-37) and the 1x1 convs / linears of
 * model/attn.py:69-109.
 *
 * Operand storage modes:
 *   TRIS_OP_K2D    row-major [rows, K]   (contraction contiguous)
 *   TRIS_OP_MN2D   row-major [K, rows]   (contraction strided; used for dgrad weights and wgrad operands)
 *   TRIS_OP_CONV   NHWC activation [n,h,w,c] read through a 4-D TMA box with zero-filled halo;
 *                  as A of a forward/dgrad 3x3 conv (contraction = taps x channels, K-major) or, with
 *                  `wgrad`=1, as A and B of a weight-gradient GEMM (contraction = pixels, MN-major).
 */
enum { TRIS_OP_K2D = 0, TRIS_OP_MN2D = 1, TRIS_OP_CONV = 2 };
enum { TRIS_ACT_NONE = 0, TRIS_ACT_RELU = 1, TRIS_ACT_QUICKGELU = 2 };
enum { TRIS_DT_BF16 = 0, TRIS_DT_F32 = 1 };

typedef struct tris_gemm_desc {
    const void* a;        /* bf16 */
    const void* b;        /* bf16 */
    void* d;              /* bf16 or f32, row-major [M, ldd] (conv: NHWC pixels x channels) */
    const float* bias;    /* [N] or NULL (added before activation) */
    const void* residual; /* bf16 [M, ldd] or NULL (added after activation) */
    float* stats;         /* [stats_parts][2N] or NULL: per-CTA partial column statistics of the stored (bf16-rounded)
                             output -- row r is written by CTA r with plain stores (rows beyond the grid are cleared), so
                             the consumer's in-order sum is bit-reproducible (BatchNorm batch statistics, model.py:18-28) */
    int32_t a_mode, b_mode;
    int32_t M, N, K;      /* GEMM extents.  conv fwd/dgrad: M = n*h*w pixels, K = taps*channels(A).
                             conv wgrad: M = channels(A), N = channels(B), K = n*h*w pixels (per tap) */
    int32_t lda, ldb, ldd; /* leading dimensions in elements (2-D modes); ldd also for conv outputs */
    int32_t img_n, img_h, img_w; /* conv geometry */
    int32_t tile_h, tile_w;      /* spatial patch handled per 128-row tile (fwd/dgrad: tile_h*tile_w <= 128;
                                    wgrad: k-block = patch, tile_h*tile_w % 16 == 0, <= 96) */
    int32_t taps;         /* 1 or 9 */
    int32_t flip;         /* 1 = dgrad tap offsets (transpose convolution) */
    int32_t wgrad;        /* 1 = conv weight-gradient form */
    int32_t b_tap_stride; /* MN2D B of a conv dgrad: element offset between taps along the contiguous dim */
    int32_t block_n;      /* 32..256, multiple of 32 (UMMA N) */
    int32_t split_k;      /* >=1; >1 requires out f32 and `splitk_ws`: the partials of the splits are stored to the
                             workspace and added in split order by a second kernel (bit-reproducible) */
    int32_t act;          /* TRIS_ACT_* */
    int32_t out_dtype;    /* TRIS_DT_* */
    int32_t atomic;       /* 1 = accumulate into d (d += result; f32 only): TMA reduce-add of the single partial when
                             split_k == 1, `d += sum of partials` in the split-K second stage otherwise */
    int32_t max_ctas;     /* 0 = one per SM */
    void* d_pre;          /* optional bf16 [M, ldd]: pre-activation (post-bias) values, saved for backward */
    const void* dact_src; /* optional bf16 [M, ldd]: epilogue computes (acc [+ residual]) * act'(dact_src) instead of
                             act(acc) [+ residual] (fuses the QuickGELU / ReLU derivative into a dgrad GEMM) */
    int32_t batch;        /* >1: `batch` independent GEMMs of extents M,N,K (2-D modes only; per-image products of the
                             cross-modal attention, model/attn.py:118-131).  Rows beyond M / K of one batch entry are
                             zero-filled by TMA, so M and K need not be multiples of the tile */
    float scale;          /* accumulator scale applied before the bias; 0 is read as 1 (no scaling) */
    int64_t a_batch_stride, b_batch_stride, d_batch_stride; /* elements; 0 for A/B = operand shared by all batches */
    int32_t stats_parts;  /* rows of `stats` (>= 1 when stats != NULL); the grid is capped to it */
    int32_t stats_mode;   /* 0: (sum x, sum x^2).  1: BatchNorm-backward sums (sum g, sum g*(y - mu)) of the stored g against
                             `stats_y` / `stats_mu` (fuses the dgamma/dbeta reduction of model.py:18-28 into the GEMM
                             that produces the upstream gradient) */
    const void* stats_y;  /* bf16, same shape / leading dimension as d */
    const float* stats_mu; /* [N] */
    const float* mask_sc; /* optional [N] with mask_sh: d *= ((stats_y * mask_sc + mask_sh) > 0), the ReLU mask of */
    const float* mask_sh; /*   relu(bn(y)) recomputed from y (2-D and conv fwd/dgrad modes, bf16 output) */
    float* splitk_ws;     /* fp32 workspace, >= split_k * M * W floats (W = taps*N for conv wgrad, N rounded up to 4 else) */
    int32_t defer_reduce; /* 1 = split-K: only store the partials; the caller reduces them later, many tensors per launch, with
                             tris_splitk_reduce_multi (rows = M, w = W, split = the effective split this call reports back) */
    int32_t split_used;   /* OUT: effective split count of this call (<= split_k) */
    void* d_norm;         /* optional bf16, same shape as d: tile-local InstanceNorm epilogue (batched K-major form, M <= 128 rows
                             per batch entry = the pixels of one image; model/attn.py:75-105).  d receives the raw product
                             (saved for backward), d_norm = in_mix * act(IN(d) * in_gamma + in_beta) [+ in_add], with the
                             per-(image, channel) mean / invstd written to in_mean / in_invstd [batch, N] */
    const float* in_gamma; const float* in_beta;
    float* in_mean; float* in_invstd;
    const void* in_add;   /* optional bf16 [batch*M, ldd] */
    float in_eps, in_mix; /* in_mix 0 is read as 1 */
    int32_t in_relu;
    int32_t conv_halo;    /* 1 = 3x3 conv forward / dgrad with halo re-use: the (16+2) x (8+2) pixel box of a 16 x 8 tile is loaded
                             once per 64-channel block and the nine taps are nine start addresses into it (instead of nine
                             shifted box loads); needs tile_h = 16, tile_w = 8 */
    const void* res_bits; /* optional uint8 [M, N/8] (N % 32 == 0): the residual is added only where its bit is set -- the residual
                             join of a bottleneck in backward: dx = dy1 W1 + dout * [out > 0] without materialising the masked
                             gradient (bits from tris_bn_apply_fwd) */
    uint64_t* tstamp;     /* optional uint64[2], initialised {UINT64_MAX, 0}: the kernel leaves min(%globaltimer at CTA start)
                             and max(%globaltimer at CTA end) there -- the device-side duration of this launch, also inside
                             a CUDA-graph replay (measurement aid of bench.py; NULL in production) */
    int32_t residual_f32; /* 1 = `residual` is fp32 [M, ldd] (the fp32 residual stream of the transformer towers; use with
                             out_dtype F32); 0 = bf16 */
} tris_gemm_desc;

int tris_gemm(tris_gemm_desc* desc, tris_stream_t stream);

/* Second stage of deferred split-K weight gradients: d[m*ldd + j] (+)= sum_s ws[(s*rows + m)*w + j], s in order. */
#define TRIS_REDUCE_MAX 64
typedef struct tris_reduce_item {
    const float* ws;
    float* d;
    int32_t rows, w, ldd, split, accumulate, pad_;
} tris_reduce_item;
int tris_splitk_reduce_multi(const tris_reduce_item* items, int n, tris_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Non-GEMM kernels.  Same conventions as above.  `void*` activation buffers are bf16 unless stated; NHWC for
 * image-tower activations, [rows, channels] row-major elsewhere; statistics / losses / gradients of parameters fp32.
 * Each entry names the reference arithmetic it replaces (paths relative to fawnliu/TRIS). */

/* ---- bn_act.cu */
/* nn.BatchNorm2d (train: batch statistics = in-order sum of the GEMM epilogue's partial rows stats[stats_parts][2C] in a
 * finalize kernel + running-stat update; eval: running stats) + ReLU + AvgPool2d(pool) + residual add / second BN branch
 * (downsample) -- CLIP/clip/model.py:18-28,36-40,42-55.  fold_half = c/2: channels c and c + c/2 are one BatchNorm channel
 * (image-pair-packed stem).  save_scale0 / save_shift0 (optional, [C]): gamma*invstd and beta - mean*gamma*invstd of branch 0,
 * kept for the backward GEMM epilogues that recompute the ReLU mask from y.  relu_bits (optional, uint8 [rows, C/8]): sign bits
 * of the output, so that the backward pass masks gradients without re-reading the bf16 output.  unpair (pool 2 only): the two
 * channel halves are the two images of a pair; the output is written un-paired as [2n, h/2, w/2, c/2]. */
int tris_bn_apply_fwd(const void* y0, const float* stats0, const float* gamma0, const float* beta0, float* rm0,
    float* rv0, float* save_mean0, float* save_invstd0, const void* y1, const float* stats1, const float* gamma1,
    const float* beta1, float* rm1, float* rv1, float* save_mean1, float* save_invstd1, const void* residual, void*
    out, int n, int h, int w, int c, int pool, int relu, int train, float momentum, float eps, int stats_parts,
    int fold_half, float* save_scale0, float* save_shift0, void* relu_bits, int unpair, tris_stream_t stream);
/* backward of the above: per-channel reductions (two-stage, fixed order: partial rows in `ws`, then a finalize kernel
 * that also adds into dgamma / dbeta) then dy (and the residual gradient g_out) -- model.py:42-55.  ws holds
 * [nparts][K][C] partial rows + [K][C] finalized sums (K = 3 with a second branch, else 2); ext_parts > 0: the rows were
 * written by the GEMM that produced `dout` (tris_gemm stats_mode 1) and `dout` is already the masked gradient.
 * relu_bits (optional): the mask written by tris_bn_apply_fwd, used instead of `out`.  unpair (pool 2 only): dout is the
 * un-paired [2n, h/2, w/2, c/2] gradient of the forward's un-paired output. */
int tris_bn_bwd(const void* dout, const void* out, const void* y0, const float* gamma0, const float* beta0, const
    float* save_mean0, const float* save_invstd0, float* dgamma0, float* dbeta0, void* dy0, const void* y1, const
    float* gamma1, const float* beta1, const float* save_mean1, const float* save_invstd1, float* dgamma1, float*
    dbeta1, void* dy1, void* g_out, int n, int h, int w, int c, int pool, int relu, int fold_half, float* ws,
    long ws_floats, int ext_parts, const void* relu_bits, int unpair, tris_stream_t stream);
/* nn.AvgPool2d(2) on NHWC bf16 (the anti-aliased stride of the downsample branch, model.py:37). */
int tris_avgpool2_fwd(const void* x, void* out, int n, int h, int w, int c, tris_stream_t stream);
/* its adjoint (+ optional accumulate source). */
int tris_avgpool2_bwd(const void* dout, const void* add, void* dx, int n, int h, int w, int c, tris_stream_t
    stream);

/* ---- head.cu */
/* the two softmaxes of bilateral_prompt (model/attn.py:122 text axis, :125 pixel axis) + pixel-centred copy of the first. */
int tris_xattn_softmax_fwd(const float* S1, const float* S2T, void* PA, void* PAc, void* PTt, int B, int P, int T,
    int Tp, float scale, tris_stream_t stream);
/* softmax backward for both. */
int tris_xattn_softmax_bwd(const void* PA, const float* dPA, const void* PTt, const float* dPTt, void* dS1, void*
    dS2T, int B, int P, int T, int Tp, float scale, tris_stream_t stream);
/* norm_lan + 0.1 * new_lan with norm_lan shared by all images (model/model_stage1.py:66,74). */
int tris_bcast_mix(const void* base, const void* x, void* out, long per, int B, float a, tris_stream_t stream);
/* adjoint of the broadcast above (sum over images). */
int tris_batch_sum(const void* x, const void* add, void* out, long per, int B, float a, tris_stream_t stream);
/* ReLU backward of the text projections (attn.py:87-97) from an fp32 gradient. */
int tris_relu_mask(const float* g, const void* y, void* dst, long n, tris_stream_t stream);
/* response head: bg class, 49-way softmax, mean/max/focal scores, diagonal maps (model/model_stage1.py:80-114, focal_loss :122-123). */
int tris_head_fwd(const float* R, const float* logit_scale, float* cls_out, float* cls_fg, float* maps, float* mbar,
    int* argmax, float* es_out, int B, int P, int T, int Tp, float focal_p, float focal_l, int train, tris_stream_t
    stream);
/* its backward: dL/d(score/exp(logit_scale)) in bf16 and d logit_scale. */
int tris_head_bwd(const float* R, const float* logit_scale, const float* dcls_out, const float* dcls_fg, const
    float* dmaps, const float* mbar, const int* argmax, void* D, float* dlogit_scale, int B, int P, int T, int Tp,
    float focal_p, float focal_l, tris_stream_t stream);
/* Upsample(bilinear, align_corners=False) + ReLU / sigmoid (model/utils.py:5-10, model_stage1.py:114-119). */
int tris_upsample_fwd(const float* maps, float* relu_out, float* sig_out, int B, int h, int w, int H, int W,
    tris_stream_t stream);
/* adjoint (gather form) onto the h x w logits. */
int tris_upsample_bwd(const float* drelu, const float* dsig, const float* sig, float* dmaps, int B, int h, int w,
    int H, int W, tris_stream_t stream);
/* fg = bilinear_ac(sig->O) * bilinear_ac(img->O) as ViT patches and/or NCHW fp32 (train_stage1.py:327-339); sig NULL = patchify only. */
int tris_mask_resize_fwd(const float* sig, const float* img, void* patches, float* fg, int B, int S, int O, int ps,
    tris_stream_t stream);
/* gradient of the above w.r.t. the sigmoid map. */
int tris_mask_resize_bwd(const void* dpatches, const float* img, float* dcam, float* dsig, int B, int S, int O, int
    ps, tris_stream_t stream);
/* PRMS map selection (validate.py:120-127, 304-332): cosine scores of the S masked-image features f [S, D] against the S
 * sentence features g [S, D] (bf16), summed over the sentences; *best = first arg-max, scores [S] optional. */
int tris_prms_select(const void* f, const void* g, float* scores, int* best, int S, int D, tris_stream_t stream);
/* validate.py:183-191 on the device: out = cam / (max + 1e-5), pred = out > 1e-9, stats = {|pred & target|, |pred | target|,
 * hit (target at the first arg-max), max}.  cams f32 [S, HW] with *sel choosing the map (NULL: map 0); target int64 [HW]. */
int tris_cam_metrics(const float* cams, const int* sel, const long long* target, float* out, float* stats, int HW,
    tris_stream_t stream);
/* F.interpolate(mode="bilinear", align_corners=True) of fp32 [nc, H, W] maps to the original image size (validate.py:180, demo.py:94). */
int tris_resize_bilinear_ac(const float* src, float* dst, long nc, int H, int W, int OH, int OW, tris_stream_t stream);
/* fg loss (clip_forward + MaxLoss, train_stage1.py:263-284,340), negative loss (:342-353), multilabel soft margin (:354), weighted sum (:364). */
int tris_stage1_loss_fwd(const void* f, const void* g, const float* cls, float* out, int B, int D, int K, float w1,
    float w4, float w5, tris_stream_t stream);
int tris_stage1_loss_fwd_f32(const float* f, const float* g, const float* cls, float* out, int B, int D, int K, float w1, float w4,
    float w5, tris_stream_t stream);
/* gradients of the weighted sum w.r.t. the image features and cls_out. */
int tris_stage1_loss_bwd(const void* f, const void* g, const float* cls, const float* dout, void* df, float* dcls,
    int B, int D, int K, float w1, float w4, float w5, tris_stream_t stream);

/* ---- misc.cu */
/* data movement of the stride-2 3x3 stem conv on the fp32 NCHW image (model.py:212-217,255-258). */
int tris_stem_im2col(const float* img, void* col, int n, int h, int w, tris_stream_t stream);
/* pair-packed stem (two images per 64-channel row, block-diagonal weights): im2col of an image pair, block-diagonal weight
 * packing / gradient un-packing (the fold of the two halves' statistics happens in the BatchNorm finalize kernels). */
int tris_stem_im2col_pair(const float* img, void* col, int n, int h, int w, tris_stream_t stream);
int tris_pack_conv_blockdiag(const float* w, void* out, int co, int ci, int khw, int reps, tris_stream_t stream);
int tris_unpack_conv_grad_blockdiag(const float* gp, float* gw, int co, int ci, int khw, int reps, tris_stream_t stream);
/* fp32 master -> bf16 operand copy of all parameters. */
int tris_f32_to_bf16(const float* src, void* dst, long n, tris_stream_t stream);
/* OIHW fp32 -> [Cout, taps*Cin] bf16 (tap-major K) for the implicit-GEMM convs. */
int tris_pack_conv(const float* w, void* out, int co, int ci, int khw, int co_pad, int ci_pad, tris_stream_t
    stream);
/* inverse layout change for the weight gradient (adds into the OIHW gradient). */
int tris_unpack_conv_grad(const float* gp, float* gw, int co, int ci, int khw, int ci_pad, tris_stream_t stream);
/* x / x.norm(dim=-1) (model_stage1.py:68-69), bf16 rows. */
int tris_l2norm_fwd(const void* x, void* y, float* inv_norm, int rows, int D, tris_stream_t stream);
/* same from fp32 rows, optionally also emitting fp32 output. */
int tris_l2norm_fwd_f32(const float* x, void* y, float* y32, float* inv_norm, int rows, int D, tris_stream_t
    stream);
/* x - mean over the pixels of each image (exact under the InstanceNorm that follows, attn.py:72-86). */
int tris_center_pixels(const float* x, void* out, int B, int P, int C, tris_stream_t stream);
/* backward of the L2 normalisation. */
int tris_l2norm_bwd(const void* dy, const void* y, const float* inv_norm, void* dx, int rows, int D, tris_stream_t
    stream);
/* nn.InstanceNorm2d(affine) [+ ReLU] [+ 0.1-residual mix] (attn.py:72-86,102-105; model_stage1.py:73). */
int tris_instnorm_fwd(const void* x, const float* gamma, const float* beta, const void* mix_add, void* out, float*
    mean, float* invstd, int batch, int P, int C, float mix_scale, int relu, float eps, tris_stream_t stream);
/* its backward.  ws (optional): fp32 [2][batch][C] per-image partial rows of dgamma (plane 0) / dbeta (plane 1), plain
 * stores -- add the rows in order with tris_splitk_reduce_multi (split = batch). */
int tris_instnorm_bwd(const void* dout, const void* x, const float* gamma, const float* beta, const float* mean,
    const float* invstd, void* dx, float* ws, int batch, int P, int C, float mix_scale, int relu,
    tris_stream_t stream);
/* y = a*x + b*y on bf16. */
int tris_axpby(const void* x, void* y, float a, float b, long n, tris_stream_t stream);

/* ---- optim.cu */
/* torch.optim.AdamW over the flat buffer, two lr groups, poly-0.9 LambdaLR, 1/world gradient scaling, bf16 shadow refresh
 * (train_stage1.py:133-144,368-372). */
int tris_adamw_step(float* p, const float* g, float* m, float* v, void* shadow, long n, long n_group0, int* step,
    float max_iter, float lr0, float lr1, float beta1, float beta2, float eps, float wd, float grad_scale, float
    power, tris_stream_t stream);
/* Device-side signal between a captured step graph and the communication stream (data-parallel overlap of the gradient
 * all-reduce with the image-tower backward, train_stage1.py:68-70): tris_flag_inc bumps *flag from inside the graph,
 * tris_flag_wait parks a one-thread kernel on `stream` until *flag has reached `target` (bounded spin, traps after ~10 s). */
int tris_flag_inc(uint32_t* flag, tris_stream_t stream);
int tris_flag_wait(const uint32_t* flag, uint32_t target, tris_stream_t stream);

/* ---- transformer.cu */
/* token_embedding[ids] + positional_embedding, EOT index = argmax(ids) (model.py:552-564).  x_f32: x is fp32 -- the residual
 * stream of a transformer stays in fp32 between blocks, as it does in the reference under autocast (fp32 embeddings +
 * low-precision branch outputs promote to fp32); branch inputs (LayerNorm outputs) are bf16. */
int tris_embed_fwd(const int* ids, const float* E, const float* P, void* x, int* eot, int n, int L, int D, int x_f32,
    tris_stream_t stream);
/* dE[ids] += dx, dP += sum_n dx: owner-computes (first occurrence of a token id adds all its occurrences in order), no
 * atomics. */
int tris_embed_bwd(const int* ids, const void* dx, float* dE, float* dP, int n, int L, int D, tris_stream_t stream);
/* LayerNorm in fp32 (model.py:352-358).  x_f32 / y_f32: input / output rows are fp32 instead of bf16. */
int tris_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd, int
    rows, int D, float eps, int x_f32, int y_f32, tris_stream_t stream);
/* its backward (+ residual-gradient add).  ws (optional): fp32 [2][ws_rows][D] per-CTA partial rows of dgamma (plane 0) /
 * dbeta (plane 1) from exactly ws_rows CTAs (1 <= ws_rows <= ceil(rows / 8)); add them in order with
 * tris_splitk_reduce_multi (split = ws_rows).  x_f32: the saved forward input x is fp32 (dy, add, dx stay bf16). */
int tris_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* mean, const float* rstd,
    const void* add, void* dx, float* ws, int ws_rows, int rows, int D, int x_f32, tris_stream_t stream);
/* softmax(q k^T / 8 [+ causal mask]) v per (sample, head), head dim 64 (nn.MultiheadAttention in model.py:366-386). */
int tris_attn_fwd(const void* qkv, void* out, int n, int L, int heads, int causal, tris_stream_t stream);
/* its backward (dq, dk, dv packed like qkv). */
int tris_attn_bwd(const void* qkv, const void* dout, void* dqkv, int n, int L, int heads, int causal, tris_stream_t
    stream);
/* row gather (EOT / class-token pooling, model.py:562, :445). */
int tris_gather_rows(const void* x, const int* idx, void* out, int rows, int D, tris_stream_t stream);
/* adjoint of the gather into a zero tensor. */
int tris_scatter_rows(const void* src, const int* idx, void* out, int rows, int D, tris_stream_t stream);
/* ws[chunk][c] = sum over the rows of chunk `chunk` of x[r,c] (bias gradients: add the chunk rows in order with
 * tris_splitk_reduce_multi, split = chunks). */
int tris_colsum(const void* x, float* ws, int chunks, int rows, int N, tris_stream_t stream);
/* class token + patch tokens + positional embedding (model.py:436-441). */
int tris_vit_assemble(const void* patch, const float* cls, const float* pos, void* tok, int n, int T, int D,
    tris_stream_t stream);

/* ---- precise.cu: fp32 parity mode (forward only, CUDA cores; all buffers fp32; BatchNorm batch sums fp64).  Same reference lines
 * as the bf16 kernels above: convolutions as im2col + SGEMM, BatchNorm train/eval + ReLU + residual, AvgPool2d(2), embedding, LayerNorm,
 * attention, L2 / InstanceNorm / softmax / 0.1-mix of the fusion. */
int tris_sgemm(const float* A, const float* B, float* C, const float* bias, const float* res, int M, int N, int K,
    int lda, int ldb, int ldc, int b_kn, int act, float alpha, int batch, long sa, long sb, long sc, tris_stream_t
    stream);
int tris_im2col3x3_f32(const float* x, float* col, int n, int h, int w, int c, int stride, int nchw, tris_stream_t
    stream);
int tris_colstats_f32(const float* x, double* stats, long rows, int c, tris_stream_t stream);
int tris_bn_f32(const float* y0, const double* stats0, const float* gamma0, const float* beta0, const float* rm0,
    const float* rv0, const float* y1, const double* stats1, const float* gamma1, const float* beta1, const float*
    rm1, const float* rv1, const float* res, float* out, long rows, int c, int relu, float eps, tris_stream_t
    stream);
int tris_avgpool2_f32(const float* x, float* out, int n, int h, int w, int c, tris_stream_t stream);
int tris_embed_f32(const int* ids, const float* E, const float* P, float* x, int* eot, int n, int L, int D,
    tris_stream_t stream);
int tris_layernorm_f32(const float* x, const float* gamma, const float* beta, float* y, int rows, int D, float eps,
    tris_stream_t stream);
int tris_attn_f32(const float* qkv, float* out, int n, int L, int heads, int causal, tris_stream_t stream);
int tris_vit_assemble_f32(const float* patch, const float* cls, const float* pos, float* tok, int n, int T, int D, tris_stream_t stream);
int tris_gather_rows_f32(const float* x, const int* idx, float* out, int rows, int D, tris_stream_t stream);
int tris_l2norm_f32(const float* x, float* y, int rows, int D, tris_stream_t stream);
int tris_instnorm_f32(const float* x, const float* gamma, const float* beta, const float* mix_add, float* out, int
    batch, int P, int C, float mix_scale, int relu, float eps, tris_stream_t stream);
int tris_softmax_f32(const float* x, float* y, int rows, int n, int ld, float scale, tris_stream_t stream);
int tris_bcast_mix_f32(const float* base, const float* x, float* out, long per, int B, float a, tris_stream_t
    stream);

#ifdef __cplusplus
}
#endif
#endif /* TRIS_SM100_H */
