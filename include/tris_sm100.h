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
    float* stats;         /* [2N] column sum / sum of squares of the stored (bf16-rounded) output, accumulated per CTA
                             in shared memory and flushed with one atomic per column (BatchNorm batch statistics,
                             model.py:18-28) or NULL */
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
    int32_t split_k;      /* >=1; >1 requires out f32 + atomic */
    int32_t act;          /* TRIS_ACT_* */
    int32_t out_dtype;    /* TRIS_DT_* */
    int32_t atomic;       /* 1 = red.add.f32 into d (d pre-zeroed by the caller) */
    int32_t max_ctas;     /* 0 = one per SM */
    void* d_pre;          /* optional bf16 [M, ldd]: pre-activation (post-bias) values, saved for backward */
    const void* dact_src; /* optional bf16 [M, ldd]: epilogue multiplies by act'(dact_src) instead of applying act
                             (fuses the QuickGELU / ReLU derivative into a dgrad GEMM) */
    int32_t batch;        /* >1: `batch` independent GEMMs of extents M,N,K (2-D modes only; per-image products of the
                             cross-modal attention, model/attn.py:118-131).  Rows beyond M / K of one batch entry are
                             zero-filled by TMA, so M and K need not be multiples of the tile */
    float scale;          /* accumulator scale applied before the bias; 0 is read as 1 (no scaling) */
    int64_t a_batch_stride, b_batch_stride, d_batch_stride; /* elements; 0 for A/B = operand shared by all batches */
} tris_gemm_desc;

int tris_gemm(const tris_gemm_desc* desc, tris_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TRIS_SM100_H */
