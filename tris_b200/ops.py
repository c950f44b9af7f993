"""Tensor-level wrappers of the non-GEMM kernels of libtris_sm100.so (see include/tris_sm100.h).

Every function takes contiguous CUDA tensors, launches on torch's current stream and returns torch tensors.
No torch arithmetic happens here; missing library / wrong device raises (``_lib``).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

bf16, f32 = torch.bfloat16, torch.float32
P = L.ptr


def _vp(t):
    return C.c_void_p(P(t))


def empty(shape, like, dtype=bf16):
    return torch.empty(shape, device=like.device, dtype=dtype)


# ------------------------------------------------------------------ BatchNorm family (NHWC)
class BNState:
    """Per-layer BatchNorm tensors handed to the kernels (all fp32 [C])."""
    __slots__ = ("gamma", "beta", "rm", "rv", "mean", "invstd", "dgamma", "dbeta", "scale", "shift")

    def __init__(self, gamma, beta, rm, rv, dgamma=None, dbeta=None):
        self.gamma, self.beta, self.rm, self.rv = gamma, beta, rm, rv
        self.mean = torch.empty_like(gamma)
        self.invstd = torch.empty_like(gamma)
        self.scale = torch.empty_like(gamma)      # gamma * invstd, beta - mean * gamma * invstd of the last train-mode forward
        self.shift = torch.empty_like(gamma)
        self.dgamma, self.dbeta = dgamma, dbeta


STAT_PARTS = 148      # rows of a partial-statistics buffer: one per CTA of the persistent GEMM (= SMs of a B200)


def bn_apply(y, stats, bn: BNState, train, relu=True, pool=1, y1=None, stats1=None, bn1: BNState = None, residual=None,
             momentum=0.1, eps=1e-5, fold_half=0, bits=None, unpair=False):
    """stats / stats1: the GEMM epilogue's partial rows [STAT_PARTS, 2C] (train mode); they are summed in a fixed order by
    the finalize kernel that precedes the apply kernel.  bits: optional uint8 [rows, C/8] receiving the sign bits of the output
    (the ReLU mask the backward kernels need, at 1/16 of the bytes of the output)."""
    n, h, w, c = y.shape
    out = empty((2 * n, h // pool, w // pool, c // 2) if unpair else (n, h // pool, w // pool, c), y)
    L.call("tris_bn_apply_fwd", _vp(y), _vp(stats), _vp(bn.gamma), _vp(bn.beta), _vp(bn.rm), _vp(bn.rv), _vp(bn.mean),
           _vp(bn.invstd), _vp(y1), _vp(stats1), _vp(bn1.gamma if bn1 else None), _vp(bn1.beta if bn1 else None),
           _vp(bn1.rm if bn1 else None), _vp(bn1.rv if bn1 else None), _vp(bn1.mean if bn1 else None),
           _vp(bn1.invstd if bn1 else None), _vp(residual), _vp(out), n, h, w, c, pool, int(relu), int(train),
           C.c_float(momentum), C.c_float(eps), STAT_PARTS if train else 0, int(fold_half), _vp(bn.scale), _vp(bn.shift),
           _vp(bits), int(unpair), launches=2 if train else 1)
    return out


_BN_BWD_PARTS = 4 * 148     # CTAs of the reduction kernel (4 resident per SM)


def bn_bwd(dout, out, y, bn: BNState, relu=True, pool=1, y1=None, bn1: BNState = None, want_g=False, fold_half=0, ext=None,
           bits=None, unpair=False):
    """Returns (dy, dy1|None, g|None); adds dgamma/dbeta into bn.dgamma/bn.dbeta (fp32).  The per-channel reductions are
    two-stage in a private workspace (never in the gradient buffers) and bit-reproducible.
    ext: partial rows [STAT_PARTS, 2C] (+ 2C spare floats) written by the GEMM that produced `dout` (gemm.py bwd_stats=):
    `dout` is then the already masked gradient and the reduction kernel is skipped."""
    n, h, w, c = y.shape
    dy = torch.empty_like(y)
    dy1 = torch.empty_like(y1) if y1 is not None else None
    g = torch.empty_like(y) if want_g else None
    k = 3 if y1 is not None else 2
    if ext is not None:
        ws, ext_parts = ext, STAT_PARTS
    else:
        ws, ext_parts = torch.empty(((_BN_BWD_PARTS + 1) * k * c,), device=y.device, dtype=f32), 0
    L.call("tris_bn_bwd", _vp(dout), _vp(out), _vp(y), _vp(bn.gamma), _vp(bn.beta), _vp(bn.mean), _vp(bn.invstd),
           _vp(bn.dgamma), _vp(bn.dbeta), _vp(dy), _vp(y1), _vp(bn1.gamma if bn1 else None),
           _vp(bn1.beta if bn1 else None), _vp(bn1.mean if bn1 else None), _vp(bn1.invstd if bn1 else None),
           _vp(bn1.dgamma if bn1 else None), _vp(bn1.dbeta if bn1 else None), _vp(dy1), _vp(g), n, h, w, c, pool,
           int(relu), int(fold_half), _vp(ws), C.c_long(ws.numel()), ext_parts, _vp(bits), int(unpair), launches=3 if ext is None else 2)
    return dy, dy1, g


def avgpool2(x):
    n, h, w, c = x.shape
    out = empty((n, h // 2, w // 2, c), x)
    L.call("tris_avgpool2_fwd", _vp(x), _vp(out), n, h, w, c)
    return out


def avgpool2_bwd(dout, add=None):
    n, ho, wo, c = dout.shape
    dx = empty((n, ho * 2, wo * 2, c), dout)
    L.call("tris_avgpool2_bwd", _vp(dout), _vp(add), _vp(dx), n, ho * 2, wo * 2, c)
    return dx


# ------------------------------------------------------------------ transformer pieces
def embed_fwd(ids, E, Ppos, want_eot=True, out_dtype=bf16):
    n, l = ids.shape
    d = E.shape[1]
    x = torch.empty((n * l, d), device=E.device, dtype=out_dtype)
    eot = torch.empty((n,), device=E.device, dtype=torch.int32) if want_eot else None
    L.call("tris_embed_fwd", _vp(ids), _vp(E), _vp(Ppos), _vp(x), _vp(eot), n, l, d, int(out_dtype == f32))
    return x, eot


def embed_bwd(ids, dx, dE, dP):
    n, l = ids.shape
    L.call("tris_embed_bwd", _vp(ids), _vp(dx), _vp(dE), _vp(dP), n, l, dE.shape[1])


def layernorm_fwd(x, gamma, beta, save=True, eps=1e-5, out_dtype=bf16):
    """x: bf16 or fp32 rows (the fp32 residual stream of the transformer towers); y: out_dtype."""
    rows, d = x.shape
    y = torch.empty((rows, d), device=x.device, dtype=out_dtype)
    mean = torch.empty((rows,), device=x.device, dtype=f32) if save else None
    rstd = torch.empty((rows,), device=x.device, dtype=f32) if save else None
    L.call("tris_layernorm_fwd", _vp(x), _vp(gamma), _vp(beta), _vp(y), _vp(mean), _vp(rstd), rows, d, C.c_float(eps),
           int(x.dtype == f32), int(out_dtype == f32))
    return y, mean, rstd


def _queue(queue):
    """(queue to push to, flush-now flag): without a caller-owned gemm.SplitKQueue the reduction runs immediately."""
    if queue is not None:
        return queue, False
    from .gemm import SplitKQueue
    return SplitKQueue(), True


def layernorm_bwd(dy, x, gamma, mean, rstd, add=None, dgamma=None, dbeta=None, queue=None):
    """dgamma / dbeta (fp32 [D], accumulated): per-CTA partial rows + an in-order sum through the queue (no atomics)."""
    rows, d = x.shape
    dx = torch.empty((rows, d), device=x.device, dtype=bf16)      # the gradient stream is bf16 whatever the dtype of x
    ws, nrows = None, 0
    if dgamma is not None:
        nrows = max(1, min((rows + 7) // 8, 148))
        ws = torch.empty((2, nrows, d), device=x.device, dtype=f32)
    L.call("tris_layernorm_bwd", _vp(dy), _vp(x), _vp(gamma), _vp(mean), _vp(rstd), _vp(add), _vp(dx), _vp(ws), nrows, rows, d,
           int(x.dtype == f32))
    if ws is not None:
        q, now = _queue(queue)
        q.push(ws[0], dgamma, 1, d, d, nrows, True)
        q.push(ws[1], dbeta, 1, d, d, nrows, True)
        if now:
            q.flush()
    return dx


def attn_fwd(qkv, n, l, heads, causal):
    out = torch.empty((n * l, heads * 64), device=qkv.device, dtype=bf16)
    L.call("tris_attn_fwd", _vp(qkv), _vp(out), n, l, heads, int(causal))
    return out


def attn_bwd(qkv, dout, n, l, heads, causal):
    dqkv = torch.empty_like(qkv)
    L.call("tris_attn_bwd", _vp(qkv), _vp(dout), _vp(dqkv), n, l, heads, int(causal))
    return dqkv


def gather_rows(x, idx):
    """Rows of a bf16 or fp32 matrix (fp32 rows are copied as pairs of 16-bit words by the same kernel)."""
    if x.dtype == f32:
        return gather_rows(x.view(bf16), idx).view(f32)
    out = torch.empty((idx.numel(), x.shape[1]), device=x.device, dtype=bf16)
    L.call("tris_gather_rows", _vp(x), _vp(idx), _vp(out), idx.numel(), x.shape[1])
    return out


def scatter_rows(src, idx, rows_total):
    out = torch.zeros((rows_total, src.shape[1]), device=src.device, dtype=bf16)
    L.call("tris_scatter_rows", _vp(src), _vp(idx), _vp(out), idx.numel(), src.shape[1])
    return out


def colsum(x, out, queue=None):
    """out[c] += sum_r x[r, c] (bias gradients): chunk partial rows + an in-order sum through the queue (no atomics)."""
    rows, n = x.shape
    assert n % 4 == 0
    chunks = max(1, min(128, (rows + 31) // 32))
    ws = torch.empty((chunks, n), device=x.device, dtype=f32)
    L.call("tris_colsum", _vp(x), _vp(ws), chunks, rows, n)
    q, now = _queue(queue)
    q.push(ws, out, 1, n, n, chunks, True)
    if now:
        q.flush()


def vit_assemble(patch, cls, pos, n):
    t, d = pos.shape
    tok = torch.empty((n * t, d), device=patch.device, dtype=bf16)
    L.call("tris_vit_assemble", _vp(patch), _vp(cls), _vp(pos), _vp(tok), n, t, d)
    return tok


# ------------------------------------------------------------------ misc
def stem_im2col(img):
    n, _, h, w = img.shape
    col = torch.empty((n * (h // 2) * (w // 2), 64), device=img.device, dtype=bf16)
    L.call("tris_stem_im2col", _vp(img), _vp(col), n, h, w)
    return col


def stem_im2col_pair(img):
    """img fp32 [N,3,H,W], N even -> col bf16 [(N/2)*(H/2)*(W/2), 64]: two images per row (columns 0..26 | 32..58)."""
    n, _, h, w = img.shape
    col = torch.empty(((n // 2) * (h // 2) * (w // 2), 64), device=img.device, dtype=bf16)
    L.call("tris_stem_im2col_pair", _vp(img), _vp(col), n, h, w)
    return col


def pack_conv_blockdiag(w, out, reps=2):
    co, ci, kh, kw = w.shape
    L.call("tris_pack_conv_blockdiag", _vp(w), _vp(out), co, ci, kh * kw, reps)


def unpack_conv_grad_blockdiag(gp, gw, reps=2):
    co, ci, kh, kw = gw.shape
    L.call("tris_unpack_conv_grad_blockdiag", _vp(gp), _vp(gw), co, ci, kh * kw, reps)


def f32_to_bf16(src, dst):
    L.call("tris_f32_to_bf16", _vp(src), _vp(dst), C.c_long(src.numel()))


def pack_conv(w, out, co_pad=None, ci_pad=None):
    co, ci, kh, kw = w.shape
    L.call("tris_pack_conv", _vp(w), _vp(out), co, ci, kh * kw, co_pad or co, ci_pad or ci)


def unpack_conv_grad(gp, gw, ci_pad=None):
    co, ci, kh, kw = gw.shape
    L.call("tris_unpack_conv_grad", _vp(gp), _vp(gw), co, ci, kh * kw, ci_pad or ci)


def l2norm_fwd(x):
    y = torch.empty_like(x)
    inv = torch.empty((x.shape[0],), device=x.device, dtype=f32)
    L.call("tris_l2norm_fwd", _vp(x), _vp(y), _vp(inv), x.shape[0], x.shape[1])
    return y, inv


def l2norm_fwd_f32(x, want_f32=True):
    """fp32 rows in -> (y bf16, y fp32 | None, inv_norm)."""
    y = torch.empty(x.shape, device=x.device, dtype=bf16)
    y32 = torch.empty_like(x) if want_f32 else None
    inv = torch.empty((x.shape[0],), device=x.device, dtype=f32)
    L.call("tris_l2norm_fwd_f32", _vp(x), _vp(y), _vp(y32), _vp(inv), x.shape[0], x.shape[1])
    return y, y32, inv


def center_pixels(x32, batch):
    """x fp32 [B*P, C] -> bf16 x - mean_pixels(x) per (image, channel)."""
    rows, c = x32.shape
    out = torch.empty((rows, c), device=x32.device, dtype=bf16)
    L.call("tris_center_pixels", _vp(x32), _vp(out), batch, rows // batch, c)
    return out


def l2norm_bwd(dy, y, inv):
    dx = torch.empty_like(y)
    L.call("tris_l2norm_bwd", _vp(dy), _vp(y), _vp(inv), _vp(dx), y.shape[0], y.shape[1])
    return dx


def instnorm_fwd(x, gamma, beta, batch, relu, mix_scale=1.0, mix_add=None, eps=1e-5):
    rows, c = x.shape
    p = rows // batch
    out = torch.empty_like(x)
    mean = torch.empty((batch, c), device=x.device, dtype=f32)
    invstd = torch.empty((batch, c), device=x.device, dtype=f32)
    L.call("tris_instnorm_fwd", _vp(x), _vp(gamma), _vp(beta), _vp(mix_add), _vp(out), _vp(mean), _vp(invstd), batch, p, c,
           C.c_float(mix_scale), int(relu), C.c_float(eps))
    return out, mean, invstd


def instnorm_bwd(dout, x, gamma, beta, mean, invstd, dgamma, dbeta, batch, relu, mix_scale=1.0, queue=None):
    rows, c = x.shape
    dx = torch.empty_like(x)
    ws = torch.empty((2, batch, c), device=x.device, dtype=f32)
    L.call("tris_instnorm_bwd", _vp(dout), _vp(x), _vp(gamma), _vp(beta), _vp(mean), _vp(invstd), _vp(dx), _vp(ws),
           batch, rows // batch, c, C.c_float(mix_scale), int(relu))
    q, now = _queue(queue)
    q.push(ws[0], dgamma, 1, c, c, batch, True)
    q.push(ws[1], dbeta, 1, c, c, batch, True)
    if now:
        q.flush()
    return dx


def axpby(x, y, a, b):
    """y = a*x + b*y (bf16, in place on y)."""
    L.call("tris_axpby", _vp(x), _vp(y), C.c_float(a), C.c_float(b), C.c_long(x.numel()))
    return y


# ------------------------------------------------------------------ Stage-1 head (head.cu)
def xattn_softmax_fwd(S1, S2T, T, scale):
    B, Pn, Tp = S1.shape
    PA = torch.empty((B, Pn, Tp), device=S1.device, dtype=bf16)
    PAc, PTt = torch.empty_like(PA), torch.empty_like(PA)
    L.call("tris_xattn_softmax_fwd", _vp(S1), _vp(S2T), _vp(PA), _vp(PAc), _vp(PTt), B, Pn, T, Tp, C.c_float(scale))
    return PA, PAc, PTt


def xattn_softmax_bwd(PA, dPA, PTt, dPTt, T, scale):
    B, Pn, Tp = PA.shape
    dS1, dS2T = torch.empty_like(PA), torch.empty_like(PA)
    L.call("tris_xattn_softmax_bwd", _vp(PA), _vp(dPA), _vp(PTt), _vp(dPTt), _vp(dS1), _vp(dS2T), B, Pn, T, Tp, C.c_float(scale))
    return dS1, dS2T


def bcast_mix(base, x, a):
    """out[b] = base + a * x[b]   (bf16; base has the shape of one batch entry)."""
    out = torch.empty_like(x)
    L.call("tris_bcast_mix", _vp(base), _vp(x), _vp(out), C.c_long(base.numel()), x.numel() // base.numel(), C.c_float(a))
    return out


def batch_sum(x, per_shape, a=1.0, add=None):
    """out = a * sum_b x[b] (+ add), bf16."""
    out = torch.empty(per_shape, device=x.device, dtype=bf16)
    L.call("tris_batch_sum", _vp(x), _vp(add), _vp(out), C.c_long(out.numel()), x.numel() // out.numel(), C.c_float(a))
    return out


def relu_mask(g, y):
    dst = torch.empty_like(y)
    L.call("tris_relu_mask", _vp(g), _vp(y), _vp(dst), C.c_long(y.numel()))
    return dst


def head_fwd(R, logit_scale, T, focal_p, focal_l, train):
    B, Pn, Tp = R.shape
    dev = R.device
    maps = torch.empty((B, Pn), device=dev, dtype=f32)
    es = torch.empty((), device=dev, dtype=f32)
    cls = fg = mbar = am = None
    if train:
        cls = torch.empty((B, T), device=dev, dtype=f32)
        fg = torch.empty((B,), device=dev, dtype=f32)
        mbar = torch.empty((B, T), device=dev, dtype=f32)
        am = torch.empty((B, T), device=dev, dtype=torch.int32)
    L.call("tris_head_fwd", _vp(R), _vp(logit_scale), _vp(cls), _vp(fg), _vp(maps), _vp(mbar), _vp(am), _vp(es), B, Pn, T, Tp,
           C.c_float(focal_p), C.c_float(focal_l), int(train))
    return cls, fg, maps, mbar, am, es


def head_bwd(R, logit_scale, dcls, dfg, dmaps, mbar, am, dlogit_scale, T, focal_p, focal_l):
    """dlogit_scale: fp32 [B] per-image partial gradients of logit_scale (written, not accumulated)."""
    B, Pn, Tp = R.shape
    assert dlogit_scale is None or dlogit_scale.numel() == B
    D = torch.empty((B, Pn, Tp), device=R.device, dtype=bf16)
    L.call("tris_head_bwd", _vp(R), _vp(logit_scale), _vp(dcls), _vp(dfg), _vp(dmaps), _vp(mbar), _vp(am), _vp(D), _vp(dlogit_scale),
           B, Pn, T, Tp, C.c_float(focal_p), C.c_float(focal_l))
    return D


def upsample_fwd(maps, h, w, H, W, want_sig=True):
    B = maps.shape[0]
    relu = torch.empty((B, 1, H, W), device=maps.device, dtype=f32)
    sig = torch.empty((B, 1, H, W), device=maps.device, dtype=f32) if want_sig else None
    L.call("tris_upsample_fwd", _vp(maps), _vp(relu), _vp(sig), B, h, w, H, W)
    return relu, sig


def upsample_bwd(drelu, dsig, sig, h, w):
    B, _, H, W = sig.shape
    dmaps = torch.empty((B, h * w), device=sig.device, dtype=f32)
    L.call("tris_upsample_bwd", _vp(drelu), _vp(dsig), _vp(sig), _vp(dmaps), B, h, w, H, W)
    return dmaps


def mask_resize_fwd(sig, img, out_size=224, ps=32, want_fg=False):
    """-> (patches bf16 [B*(O/ps)^2, 3*ps*ps], fg f32 [B,3,O,O] | None).  sig None = plain patchify of img."""
    B, _, S, S2 = img.shape
    assert S == S2, "square inputs"
    g = out_size // ps
    patches = torch.empty((B * g * g, 3 * ps * ps), device=img.device, dtype=bf16)
    fg = torch.empty((B, 3, out_size, out_size), device=img.device, dtype=f32) if want_fg else None
    L.call("tris_mask_resize_fwd", _vp(sig), _vp(img), _vp(patches), _vp(fg), B, S, out_size, ps)
    return patches, fg


def mask_resize_bwd(dpatches, img, out_size=224, ps=32):
    B, _, S, _ = img.shape
    dcam = torch.empty((B, out_size, out_size), device=img.device, dtype=f32)
    dsig = torch.empty((B, 1, S, S), device=img.device, dtype=f32)
    L.call("tris_mask_resize_bwd", _vp(dpatches), _vp(img), _vp(dcam), _vp(dsig), B, S, out_size, ps, launches=2)
    return dsig


def stage1_loss_fwd(f, g, cls, K, w):
    B, D = f.shape
    out = torch.empty((4,), device=f.device, dtype=f32)
    L.call("tris_stage1_loss_fwd", _vp(f), _vp(g), _vp(cls), _vp(out), B, D, K, C.c_float(w[0]), C.c_float(w[1]), C.c_float(w[2]))
    return out


def stage1_loss_bwd(f, g, cls, dout, K, w):
    B, D = f.shape
    df = torch.empty_like(f)
    dcls = torch.empty_like(cls)
    L.call("tris_stage1_loss_bwd", _vp(f), _vp(g), _vp(cls), _vp(dout), _vp(df), _vp(dcls), B, D, K, C.c_float(w[0]), C.c_float(w[1]),
           C.c_float(w[2]))
    return df, dcls


def resize_bilinear_ac(x, oh, ow):
    """fp32 [N,C,H,W] -> [N,C,oh,ow], bilinear, align_corners=True."""
    n, c, h, w = x.shape
    out = torch.empty((n, c, oh, ow), device=x.device, dtype=f32)
    L.call("tris_resize_bilinear_ac", _vp(x.contiguous()), _vp(out), C.c_long(n * c), h, w, oh, ow)
    return out


def prms_select(f, g):
    """f, g bf16 [S, D] -> (best int32 [1] on the device, scores f32 [S]): PRMS choice without a host round trip."""
    S, D = f.shape
    scores = torch.empty((S,), device=f.device, dtype=f32)
    best = torch.empty((1,), device=f.device, dtype=torch.int32)
    L.call("tris_prms_select", _vp(f.contiguous()), _vp(g.contiguous()), _vp(scores), _vp(best), S, D)
    return best, scores


def cam_metrics(cams, target, sel=None, stats=None):
    """cams f32 [S, H, W] (or [H, W]), target int64 [H, W] -> (normalised map f32 [H, W], stats f32 [4] = I, U, hit, max)."""
    hw = target.numel()
    cams = cams.reshape(-1, hw).contiguous()
    out = torch.empty(target.shape, device=cams.device, dtype=f32)
    if stats is None:
        stats = torch.empty((4,), device=cams.device, dtype=f32)
    L.call("tris_cam_metrics", _vp(cams), _vp(sel), _vp(target.contiguous()), _vp(out), _vp(stats), hw)
    return out, stats
