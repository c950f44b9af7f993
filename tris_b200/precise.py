"""fp32 parity mode of the Stage-1 forward (``model.set_precision("fp32")``): forward-only evaluation of the same network
in full fp32 on the CUDA-core kernels of csrc/precise.cu, for checking response maps against the reference's fp32
forward at the north-star tolerance (1e-3 rel).  Same math / citations as resnet.py, transformer.py and head.py; weights
are the fp32 masters of the flat parameter store (no bf16 rounding anywhere).  BatchNorm in train mode uses batch
statistics (fp64 sums) but does NOT update the running statistics (verification mode, not a training mode).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib as L
from . import ops

f32 = torch.float32
_vp = lambda t: C.c_void_p(L.ptr(t))
MIX = 0.1


def sgemm(a, b, M, N, K, lda=None, ldb=None, out=None, ldc=None, bias=None, res=None, b_kn=False, act=L.ACT_NONE, alpha=1.0,
          batch=1, sa=0, sb=0, sc=0):
    """out[M,N] = act(alpha * a[M,K] . b + bias) + res   (b: [N,K] rows, or [K,N] rows with b_kn); fp32, optional batch."""
    lda = K if lda is None else lda
    ldb = (N if b_kn else K) if ldb is None else ldb
    ldc = N if ldc is None else ldc
    if out is None:
        out = torch.empty((batch * M, N) if batch > 1 else (M, N), device=a.device, dtype=f32)
        sc = M * N if batch > 1 else 0
    L.call("tris_sgemm", _vp(a), _vp(b), _vp(out), _vp(bias), _vp(res), M, N, K, lda, ldb, ldc, int(b_kn), act, C.c_float(alpha), batch,
           C.c_long(sa), C.c_long(sb), C.c_long(sc))
    return out


class PreciseStage1:
    def __init__(self, model):
        self.model = model
        self.eng = model.engine()
        self.st = self.eng.store
        self.bufs = dict(model.named_buffers())
        self.dev = self.st.device

    # ------------------------------------------------------------------ conv / BN helpers (NHWC fp32)
    def _w(self, key):
        return self.st.p(key)

    def conv1x1(self, x, key):
        n, h, w, c = x.shape
        wt = self._w(key)
        return sgemm(x.reshape(-1, c), wt.view(wt.shape[0], -1), n * h * w, wt.shape[0], c).view(n, h, w, -1)

    def conv3x3(self, x, key, stride=1, nchw=False):
        if nchw:
            n, c, h, w = x.shape
        else:
            n, h, w, c = x.shape
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        col = torch.empty((n * ho * wo, 9 * c), device=x.device, dtype=f32)
        L.call("tris_im2col3x3_f32", _vp(x.contiguous()), _vp(col), n, h, w, c, stride, int(nchw))
        wt = self._w(key)
        wp = wt.permute(0, 2, 3, 1).reshape(wt.shape[0], -1).contiguous()      # [co, (r,s,ci)]: layout only
        return sgemm(col, wp, n * ho * wo, wt.shape[0], 9 * c).view(n, ho, wo, -1)

    def _bn_args(self, y, key, train):
        c = y.shape[-1]
        stats = None
        if train:
            stats = torch.empty(2 * c, device=y.device, dtype=torch.float64)
            L.call("tris_colstats_f32", _vp(y), _vp(stats), C.c_long(y.numel() // c), c, launches=2)
        return [_vp(y), _vp(stats), _vp(self._w(key + ".weight")), _vp(self._w(key + ".bias")), _vp(self.bufs[key + ".running_mean"]),
                _vp(self.bufs[key + ".running_var"])], stats

    def bn(self, y, key, train, relu=True, y1=None, key1=None, res=None):
        c = y.shape[-1]
        a0, keep0 = self._bn_args(y, key, train)
        a1, keep1 = self._bn_args(y1, key1, train) if y1 is not None else ([_vp(None)] * 6, None)
        out = torch.empty_like(y)
        L.call("tris_bn_f32", *a0, *a1, _vp(res), _vp(out), C.c_long(y.numel() // c), c, int(relu), C.c_float(1e-5))
        return out

    def avgpool(self, x):
        n, h, w, c = x.shape
        out = torch.empty((n, h // 2, w // 2, c), device=x.device, dtype=f32)
        L.call("tris_avgpool2_f32", _vp(x), _vp(out), n, h, w, c)
        return out

    # ------------------------------------------------------------------ towers
    def resnet(self, img, train):
        p = self.eng.resnet.prefix
        x = self.bn(self.conv3x3(img.float(), p + "conv1.weight", stride=2, nchw=True), p + "bn1", train)
        x = self.bn(self.conv3x3(x, p + "conv2.weight"), p + "bn2", train)
        x = self.bn(self.conv3x3(x, p + "conv3.weight"), p + "bn3", train)
        x = self.avgpool(x)
        for blk in self.eng.resnet.blocks:
            q = blk.p
            o = self.bn(self.conv1x1(x, q + "conv1.weight"), q + "bn1", train)
            o = self.bn(self.conv3x3(o, q + "conv2.weight"), q + "bn2", train)
            if blk.stride > 1:
                o = self.avgpool(o)
            y3 = self.conv1x1(o, q + "conv3.weight")
            if blk.down:
                idt = self.avgpool(x) if blk.stride > 1 else x
                yd = self.conv1x1(idt, q + "downsample.0.weight")
                x = self.bn(y3, q + "bn3", train, y1=yd, key1=q + "downsample.1")
            else:
                x = self.bn(y3, q + "bn3", train, res=x)
        return x                                                            # [B, H/32, W/32, 2048]

    def _ln(self, x, key, w=None):
        w = w or self._w
        y = torch.empty_like(x)
        L.call("tris_layernorm_f32", _vp(x), _vp(w(key + ".weight")), _vp(w(key + ".bias")), _vp(y), x.shape[0], x.shape[1], C.c_float(1e-5))
        return y

    def _blocks(self, x, n, l, prefix, heads, causal, w=None):
        """12 ResidualAttentionBlocks (CLIP/clip/model.py:366-386) on fp32 rows x [n*l, d]."""
        w = w or self._w
        d = x.shape[1]
        for i in range(12):
            k = f"{prefix}transformer.resblocks.{i}."
            h = self._ln(x, k + "ln_1", w)
            qkv = sgemm(h, w(k + "attn.in_proj_weight"), n * l, 3 * d, d, bias=w(k + "attn.in_proj_bias"))
            a = torch.empty((n * l, d), device=self.dev, dtype=f32)
            L.call("tris_attn_f32", _vp(qkv), _vp(a), n, l, heads, int(causal))
            x1 = sgemm(a, w(k + "attn.out_proj.weight"), n * l, d, d, bias=w(k + "attn.out_proj.bias"), res=x)
            h2 = self._ln(x1, k + "ln_2", w)
            u = sgemm(h2, w(k + "mlp.c_fc.weight"), n * l, 4 * d, d, bias=w(k + "mlp.c_fc.bias"), act=L.ACT_QUICKGELU)
            x = sgemm(u, w(k + "mlp.c_proj.weight"), n * l, d, 4 * d, bias=w(k + "mlp.c_proj.bias"), res=x1)
        return x

    def text(self, ids, prefix=None, w=None):
        """CLIP.encode_text (model.py:552-564) -> hidden fp32 [N, E]."""
        p = self.eng.text.prefix if prefix is None else prefix
        w = w or self._w
        n, l = ids.shape
        ids = ids.to(torch.int32).contiguous()
        E, Ppos = w(p + "token_embedding.weight"), w(p + "positional_embedding")
        d = E.shape[1]
        x = torch.empty((n * l, d), device=self.dev, dtype=f32)
        eot = torch.empty((n,), device=self.dev, dtype=torch.int32)
        L.call("tris_embed_f32", _vp(ids), _vp(E), _vp(Ppos), _vp(x), _vp(eot), n, l, d)
        x = self._blocks(x, n, l, p, d // 64, True, w)
        xe = torch.empty((n, d), device=self.dev, dtype=f32)
        L.call("tris_gather_rows_f32", _vp(x), _vp(eot), _vp(xe), n, d)
        xn = self._ln(xe, p + "ln_final", w)
        proj = w(p + "text_projection")                                     # [512, E] = [K, N]
        return sgemm(xn, proj, n, proj.shape[1], d, b_kn=True)

    def vit(self, fg, w):
        """VisionTransformer.forward (model.py:419-448) on fg fp32 [N,3,224,224] with the aux model's fp32 weights."""
        n, c, s, _ = fg.shape
        ps, g = 32, s // 32
        patches = fg.reshape(n, c, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(n * g * g, c * ps * ps).contiguous()   # layout only
        wc = w("visual.conv1.weight")
        d = wc.shape[0]
        pe = sgemm(patches, wc.view(d, -1), n * g * g, d, c * ps * ps)
        t = g * g + 1
        tok = torch.empty((n * t, d), device=self.dev, dtype=f32)
        L.call("tris_vit_assemble_f32", _vp(pe), _vp(w("visual.class_embedding")), _vp(w("visual.positional_embedding")), _vp(tok), n, t, d)
        x = self._blocks(self._ln(tok, "visual.ln_pre", w), n, t, "visual.", d // 64, False, w)
        idx = (torch.arange(n, device=self.dev, dtype=torch.int32) * t).contiguous()
        xc = torch.empty((n, d), device=self.dev, dtype=f32)
        L.call("tris_gather_rows_f32", _vp(x), _vp(idx), _vp(xc), n, d)
        proj = w("visual.proj")
        return sgemm(self._ln(xc, "visual.ln_post", w), proj, n, proj.shape[1], d, b_kn=True)

    @torch.no_grad()
    def step_losses(self, aux, img, word_ids, neg_word_ids, w=(1.0, 5.0, 2.0)):
        """Forward of the whole training step in fp32 (train_stage1.py:320-364): -> (loss, l1, l4, l5) fp32 [4]."""
        aw = aux._engine().store.p
        cls, _, _, sig, _ = self.forward(img, word_ids, True)
        fg = ops.mask_resize_fwd(sig.contiguous(), img.float().contiguous(), 224, 32, want_fg=True)[1]
        f = self.vit(fg, aw)
        k = 0 if neg_word_ids is None else neg_word_ids.shape[1]
        ids = word_ids if k == 0 else torch.cat([word_ids, neg_word_ids.reshape(-1, word_ids.shape[1])], 0)
        g = self.text(ids, prefix="", w=aw)
        out = torch.empty(4, device=self.dev, dtype=f32)
        L.call("tris_stage1_loss_fwd_f32", _vp(f), _vp(g), _vp(cls), _vp(out), img.shape[0], f.shape[1], k, C.c_float(w[0]), C.c_float(w[1]),
               C.c_float(w[2]))
        return out

    # ------------------------------------------------------------------ head (model_stage1.py:61-119, attn.py:111-136)
    def _l2(self, x):
        y = torch.empty_like(x)
        L.call("tris_l2norm_f32", _vp(x), _vp(y), x.shape[0], x.shape[1])
        return y

    def _in(self, x, g, b, batch, relu, mix_scale=1.0, mix_add=None):
        out = torch.empty_like(x)
        L.call("tris_instnorm_f32", _vp(x), _vp(g), _vp(b), _vp(mix_add), _vp(out), batch, x.shape[0] // batch, x.shape[1],
               C.c_float(mix_scale), int(relu), C.c_float(1e-5))
        return out

    def _softmax(self, x, rows, n, scale):
        y = torch.empty_like(x)
        L.call("tris_softmax_f32", _vp(x), _vp(y), rows, n, n, C.c_float(scale))
        return y

    def head(self, c4, hidden, img_size, train):
        hd = self.eng.head
        Pf = hd._views("flat")
        B, h, w, cv = c4.shape
        Pn, T, Cc = h * w, hidden.shape[0], hd.C
        if T != B:
            raise ValueError(f"TRIS.forward pairs image i with sentence i: got {B} images, {T} sentences")
        Tp = (T + 7) // 8 * 8
        nv = self._l2(sgemm(c4.reshape(B * Pn, cv), Pf["Wv"], B * Pn, Cc, cv, bias=Pf["bv"]))
        nl = self._l2(sgemm(hidden, Pf["Wl"], T, Cc, hidden.shape[1], bias=Pf["bl"]))
        if hd.fuse:
            A3 = self._in(sgemm(nv, Pf["Wqkv"], B * Pn, 3 * Cc, Cc, bias=Pf["bqkv"]), Pf["gqkv"], Pf["beqkv"], B, True)
            At3 = sgemm(nl, Pf["Wt"], T, 3 * Cc, Cc, bias=Pf["bt"], act=L.ACT_RELU)
            qv, kv, vv = A3[:, :Cc], A3[:, Cc:2 * Cc], A3[:, 2 * Cc:]
            qt, kt, vt = At3[:, :Cc], At3[:, Cc:2 * Cc], At3[:, 2 * Cc:]
            sc = 1.0 / math.sqrt(Cc)
            PA = self._softmax(sgemm(qv, kt, B * Pn, T, Cc, lda=3 * Cc, ldb=3 * Cc), B * Pn, T, sc)                     # [BP, T]
            S2 = torch.empty((B * T, Pn), device=self.dev, dtype=f32)
            sgemm(qt, kv, T, Pn, Cc, lda=3 * Cc, ldb=3 * Cc, out=S2, ldc=Pn, batch=B, sa=0, sb=Pn * 3 * Cc, sc=T * Pn)
            PT = self._softmax(S2, B * T, Pn, sc)                                                                         # [B, T, P]
            nvp = sgemm(PA, vt, B * Pn, Cc, T, ldb=3 * Cc, b_kn=True)
            nlp = torch.empty((B * T, Cc), device=self.dev, dtype=f32)
            sgemm(PT, vv, T, Cc, Pn, ldb=3 * Cc, b_kn=True, out=nlp, ldc=Cc, batch=B, sa=T * Pn, sb=Pn * 3 * Cc, sc=T * Cc)
            vp = self._in(sgemm(nvp, Pf["Wo"], B * Pn, Cc, Cc, bias=Pf["bo"]), Pf["go"], Pf["beo"], B, False, MIX, nv)
            Ol = sgemm(nlp, Pf["Wto"], B * T, Cc, Cc, bias=Pf["bto"])
            lp = torch.empty_like(Ol)
            L.call("tris_bcast_mix_f32", _vp(nl), _vp(Ol), _vp(lp), C.c_long(T * Cc), B, C.c_float(MIX))
            lp_bs = T * Cc
        else:
            vp, lp, lp_bs = nv, nl, 0
        R = torch.zeros((B, Pn, Tp), device=self.dev, dtype=f32)
        sgemm(vp, lp, Pn, T, Cc, out=R, ldc=Tp, batch=B, sa=Pn * Cc, sb=lp_bs, sc=Pn * Tp)
        cls, fg, maps, _, _, es = ops.head_fwd(R, Pf["ls"], T, hd.focal_p, hd.focal_lambda, train)
        relu, sig = ops.upsample_fwd(maps, h, w, img_size[0], img_size[1], want_sig=train)
        return cls, fg, relu, sig, es

    @torch.no_grad()
    def forward(self, x, word_id, train):
        c4 = self.resnet(x, train)
        hidden = self.text(word_id)
        out = self.head(c4, hidden, tuple(x.shape[2:]), train)
        return out if train else out[2]
