"""Stage-1 head on libtris_sm100 kernels: vis/lan projection + L2-norm (K6), bilateral cross-modal attention + 0.1 mix
+ score (K7), response head and x32 bilinear maps (K8).  Forward AND hand-written backward.

Restates model/model_stage1.py:61-119 and model/attn.py:111-136 of the reference (Appendix A of SURVEY.md):
    vis = W_v c4 + b_v ; nv = vis/|vis|                    lan = W_l hidden + b_l ; nl = lan/|lan|
    [Qv|Kv|Vv] = relu(IN(nvc W_qkv^T))     (per image)     [Qt|Kt|Vt] = relu(nl W_t^T + b)   (shared by all images)
    PA = softmax_T(Qv Kt^T / sqrt(C))                      PT = softmax_P(Qt Kv^T / sqrt(C))
    v' = nv + 0.1 IN(W_o (PAc Vt) + b_o)                   l'_b = nl + 0.1 (W_to (PT_b Vv_b) + b_to)
    R_b = v'_b l'_b^T ; score = exp(logit_scale) R         -> cls_out / cls_fg / maps (head.cu)

Every dense contraction is a tcgen05 GEMM call (gemm.py; batched per image where an operand differs per image), the
text-side projections are computed ONCE for the T texts (the reference repeats them B times), normalisations /
softmaxes / reductions are the kernels of misc.cu and head.cu.  Activations bf16, statistics and scores fp32.
All [., T] operands are stored with a leading dimension Tp = T rounded up to 8 (16-byte TMA strides); padded columns are
never read (TMA extents stop at T) and written as zero.
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from . import gemm as G
from . import ops

bf16, f32 = torch.bfloat16, torch.float32
K2D, MN2D = L.OP_K2D, L.OP_MN2D
import os
# 1x1 conv + InstanceNorm [+ ReLU / 0.1-mix] as ONE batched GEMM with a tile-local normalisation epilogue (gemm.linear_in_fwd)
FUSE_IN = os.environ.get("TRIS_FUSE_IN", "1") != "0"
MIX = 0.1   # hard-coded in the reference (model_stage1.py:73-74); args.attn_multi only gates the fusion on/off


class Stage1Head:
    def __init__(self, eng, model):
        self.eng = eng
        self.st = eng.store
        self.fuse = hasattr(model, "attn_fusion") and model.args.attn_multi > 0
        self.focal_p, self.focal_lambda = float(model.args.FOCAL_P), float(model.args.FOCAL_LAMBDA)
        self.C = model.args.hidden_dim

    # ------------------------------------------------------------------ parameter views
    def _views(self, buf):
        """buf = 'shadow' (bf16 GEMM operands), 'flat' (fp32 masters) or 'grad' (fp32 gradients)."""
        st, C = self.st, self.C
        get = {"shadow": st.s, "flat": st.p, "grad": st.g}[buf]
        v = {"Wv": get("vis_project.weight").view(C, -1), "bv": get("vis_project.bias"),
             "Wl": get("lan_project.weight"), "bl": get("lan_project.bias"), "ls": get("logit_scale")}
        if self.fuse:
            a = "attn_fusion."
            cat = lambda fmt, shape: st.cat(buf, [a + fmt.format(i) for i in (1, 2, 3)], shape)
            v.update(Wqkv=cat("v_proj{}.0.weight", (3 * C, C)), bqkv=cat("v_proj{}.0.bias", (3 * C,)),
                     gqkv=cat("v_proj{}.1.weight", (3 * C,)), beqkv=cat("v_proj{}.1.bias", (3 * C,)),
                     Wt=cat("t_proj{}.0.weight", (3 * C, C)), bt=cat("t_proj{}.0.bias", (3 * C,)),
                     Wo=get(a + "v_output.0.weight").view(C, C), bo=get(a + "v_output.0.bias"),
                     go=get(a + "v_output.1.weight"), beo=get(a + "v_output.1.bias"),
                     Wto=get(a + "t_output.0.weight"), bto=get(a + "t_output.0.bias"))
        return v

    # ------------------------------------------------------------------ forward
    def _fwd(self, c4, hidden, img_size, training, save):
        W, Pf = self._views("shadow"), self._views("flat")
        B, h, w, cv = c4.shape
        Pn, T, C = h * w, hidden.shape[0], self.C
        if T != B:
            raise ValueError(f"TRIS.forward pairs image i with sentence i (score[i,:,i]): got {B} images, {T} sentences")
        Tp = (T + 7) // 8 * 8
        dev = c4.device
        X0 = c4.reshape(B * Pn, cv)
        vis = G.linear_fwd(X0, W["Wv"], Pf["bv"], out_dtype=f32)       # fp32: its pixel-varying part is what matters below
        nv, nv32, inv_v = ops.l2norm_fwd_f32(vis, want_f32=self.fuse)
        lan = G.linear_fwd(hidden, W["Wl"], Pf["bl"])
        nl, inv_l = ops.l2norm_fwd(lan)
        t = {"X0": X0, "hidden": hidden, "nv": nv, "inv_v": inv_v, "nl": nl, "inv_l": inv_l, "dims": (B, h, w, Pn, T, Tp)}
        if self.fuse:
            # the conv biases in front of an InstanceNorm cancel exactly (IN removes the per-channel pixel mean); they are
            # left out so that the bf16 rounding of Yv / Ov is relative to the pixel-varying part only
            # and the operand is centred over the pixels of each image in fp32 first (ops.center_pixels): exact, see misc.cu
            nvc = ops.center_pixels(nv32, B)
            fused_in = FUSE_IN and B >= 2 and Pn <= 128
            if fused_in:      # Q|K|V projection + InstanceNorm + ReLU in one kernel (image-aligned tiles)
                Yv, A3, mu3, is3 = G.linear_in_fwd(nvc, W["Wqkv"], B, Pf["gqkv"], Pf["beqkv"], relu=True)
            else:
                Yv = G.linear_fwd(nvc, W["Wqkv"])                                               # [BP, 3C]
                A3, mu3, is3 = ops.instnorm_fwd(Yv, Pf["gqkv"], Pf["beqkv"], B, relu=True)
            At3 = G.linear_fwd(nl, W["Wt"], Pf["bt"], act=L.ACT_RELU)                           # [T, 3C]
            qv, kv, vv = A3[:, :C], A3[:, C:2 * C], A3[:, 2 * C:]
            qt, kt, vt = At3[:, :C], At3[:, C:2 * C], At3[:, 2 * C:]
            S1 = torch.empty((B, Pn, Tp), device=dev, dtype=f32)
            S2T = torch.empty((B, Pn, Tp), device=dev, dtype=f32)
            G.gemm_ex(qv, kt, S1, B * Pn, T, C, lda=3 * C, ldb=3 * C, ldd=Tp)                   # Qv Kt^T
            G.gemm_ex(kv, qt, S2T, B * Pn, T, C, lda=3 * C, ldb=3 * C, ldd=Tp)                  # (Qt Kv^T)^T
            PA, PAc, PTt = ops.xattn_softmax_fwd(S1, S2T, T, 1.0 / math.sqrt(C))
            nvp = torch.empty((B * Pn, C), device=dev, dtype=bf16)
            G.gemm_ex(PAc, vt, nvp, B * Pn, C, T, b_mode=MN2D, lda=Tp, ldb=3 * C)               # (PA - mean_p PA) Vt
            nlp = torch.empty((B * T, C), device=dev, dtype=bf16)
            G.gemm_ex(PTt, vv, nlp, T, C, Pn, a_mode=MN2D, b_mode=MN2D, lda=Tp, ldb=3 * C, batch=B, a_bs=Pn * Tp,
                      b_bs=Pn * 3 * C, d_bs=T * C)                                              # PT_b Vv_b
            if fused_in:      # v_output + InstanceNorm + 0.1-mix + residual in one kernel
                Ov, vp, muo, iso = G.linear_in_fwd(nvp, W["Wo"], B, Pf["go"], Pf["beo"], relu=False, mix_scale=MIX, mix_add=nv)
            else:
                Ov = G.linear_fwd(nvp, W["Wo"])
                vp, muo, iso = ops.instnorm_fwd(Ov, Pf["go"], Pf["beo"], B, relu=False, mix_scale=MIX, mix_add=nv)
            Ol = G.linear_fwd(nlp, W["Wto"], Pf["bto"])
            lp = ops.bcast_mix(nl, Ol, MIX)                                                     # [B*T, C]
            lp_bs = T * C
            if save:
                t.update(nvc=nvc, Yv=Yv, A3=A3, mu3=mu3, is3=is3, At3=At3, PA=PA, PAc=PAc, PTt=PTt, nvp=nvp, nlp=nlp, Ov=Ov, muo=muo, iso=iso)
        else:
            vp, lp, lp_bs = nv, nl, 0
        R = torch.empty((B, Pn, Tp), device=dev, dtype=f32)
        G.gemm_ex(vp, lp, R, Pn, T, C, batch=B, a_bs=Pn * C, b_bs=lp_bs, d_bs=Pn * Tp, ldd=Tp)   # v'_b l'_b^T
        cls, fg, maps, mbar, am, es = ops.head_fwd(R, Pf["ls"], T, self.focal_p, self.focal_lambda, training)
        relu, sig = ops.upsample_fwd(maps, h, w, img_size[0], img_size[1], want_sig=training)
        if save:
            t.update(vp=vp, lp=lp, lp_bs=lp_bs, R=R, mbar=mbar, am=am, sig=sig)
        return (cls, fg, relu, sig, es), (t if save else None)

    def forward(self, c4, hidden, img_size, training):
        if training and torch.is_grad_enabled():
            return _HeadFn.apply(c4, hidden, self, tuple(img_size))
        with torch.no_grad():
            out, _ = self._fwd(c4, hidden, tuple(img_size), training, save=False)
        return out if training else (out[2],)

    # ------------------------------------------------------------------ backward
    def _wgrad(self, dy, x, gw):
        """gw[N,K] (fp32, accumulated) += dy[M,N]^T x[M,K]; split-K second stages are deferred to one launch at the end
        of the backward (self._q.flush())."""
        G.linear_wgrad(dy, x, out=gw, accumulate=True, queue=self._q)

    def _bwd(self, t, dcls, dfg, drelu, dsig, des):
        W, Pf, Gr = self._views("shadow"), self._views("flat"), self._views("grad")
        self._q = G.SplitKQueue()
        B, h, w, Pn, T, Tp = t["dims"]
        C = self.C
        dev = t["nv"].device
        dmaps = ops.upsample_bwd(drelu, dsig, t["sig"], h, w) if (drelu is not None or dsig is not None) else None
        dls = torch.empty((B,), device=dev, dtype=f32)
        D = ops.head_bwd(t["R"], Pf["ls"], dcls, dfg, dmaps, t["mbar"], t["am"], dls, T, self.focal_p, self.focal_lambda)
        Gr["ls"].add_(dls.sum())        # B per-image partials added in a fixed order (the kernel used to atomicAdd them)
        if des is not None:
            Gr["ls"].add_(des * Pf["ls"].exp())
        vp, lp, lp_bs = t["vp"], t["lp"], t["lp_bs"]
        dvp = torch.empty((B * Pn, C), device=dev, dtype=bf16)
        G.gemm_ex(D, lp, dvp, Pn, C, T, b_mode=MN2D, lda=Tp, ldb=C, batch=B, a_bs=Pn * Tp, b_bs=lp_bs, d_bs=Pn * C)
        dlp = torch.empty((B * T, C), device=dev, dtype=bf16)
        G.gemm_ex(D, vp, dlp, T, C, Pn, a_mode=MN2D, b_mode=MN2D, lda=Tp, ldb=C, batch=B, a_bs=Pn * Tp, b_bs=Pn * C, d_bs=T * C)
        nv, nl = t["nv"], t["nl"]
        if self.fuse:
            A3, At3, PA, PAc, PTt, nvp, nlp = t["A3"], t["At3"], t["PA"], t["PAc"], t["PTt"], t["nvp"], t["nlp"]
            qv, kv, vv = A3[:, :C], A3[:, C:2 * C], A3[:, 2 * C:]
            qt, kt, vt = At3[:, :C], At3[:, C:2 * C], At3[:, 2 * C:]
            # ---- v' = nv + 0.1 IN(Ov) ; Ov = nvp Wo^T + bo
            dOv = ops.instnorm_bwd(dvp, t["Ov"], Pf["go"], Pf["beo"], t["muo"], t["iso"], Gr["go"], Gr["beo"], B, relu=False,
                                   mix_scale=MIX, queue=self._q)
            self._wgrad(dOv, nvp, Gr["Wo"])
            ops.colsum(dOv, Gr["bo"], queue=self._q)
            dnvp = G.linear_dgrad(dOv, W["Wo"])
            # ---- l'_b = nl + 0.1 (nlp_b Wto^T + bto)
            dnl = ops.batch_sum(dlp, (T, C))
            dOl = ops.axpby(dlp, dlp, MIX, 0.0)
            self._wgrad(dOl, nlp, Gr["Wto"])
            ops.colsum(dOl, Gr["bto"], queue=self._q)
            dnlp = G.linear_dgrad(dOl, W["Wto"])
            # ---- nvp = PA Vt
            dPA = torch.empty((B, Pn, Tp), device=dev, dtype=f32)
            G.gemm_ex(dnvp, vt, dPA, B * Pn, T, C, ldb=3 * C, ldd=Tp)
            dAt3 = torch.zeros((T, 3 * C), device=dev, dtype=f32)
            self._wgrad_ex(PAc, Tp, dnvp, C, dAt3[:, 2 * C:], T, C, B * Pn)                      # dVt = PAc^T dnvp
            # ---- nlp_b = PT_b Vv_b
            dPTt = torch.empty((B, Pn, Tp), device=dev, dtype=f32)
            G.gemm_ex(vv, dnlp, dPTt, Pn, T, C, lda=3 * C, batch=B, a_bs=Pn * 3 * C, b_bs=T * C, d_bs=Pn * Tp, ldd=Tp)
            dA3 = torch.empty((B * Pn, 3 * C), device=dev, dtype=bf16)
            G.gemm_ex(PTt, dnlp, dA3[:, 2 * C:], Pn, C, T, b_mode=MN2D, lda=Tp, ldb=C, ldd=3 * C, batch=B, a_bs=Pn * Tp,
                      b_bs=T * C, d_bs=Pn * 3 * C)                                              # dVv_b = PT_b^T dnlp_b
            # ---- softmaxes
            dS1, dS2T = ops.xattn_softmax_bwd(PA, dPA, PTt, dPTt, T, 1.0 / math.sqrt(C))
            G.gemm_ex(dS1, kt, dA3[:, :C], B * Pn, C, T, b_mode=MN2D, lda=Tp, ldb=3 * C, ldd=3 * C)       # dQv = dS1 Kt
            G.gemm_ex(dS2T, qt, dA3[:, C:2 * C], B * Pn, C, T, b_mode=MN2D, lda=Tp, ldb=3 * C, ldd=3 * C)  # dKv = dS2^T Qt
            self._wgrad_ex(dS1, Tp, qv, 3 * C, dAt3[:, C:2 * C], T, C, B * Pn)                   # dKt = dS1^T Qv
            self._wgrad_ex(dS2T, Tp, kv, 3 * C, dAt3[:, :C], T, C, B * Pn)                       # dQt = dS2 Kv
            # ---- text projections (ReLU, shared by all images)
            dYt = ops.relu_mask(dAt3, At3)
            self._wgrad(dYt, nl, Gr["Wt"])
            ops.colsum(dYt, Gr["bt"], queue=self._q)
            dnl = G.linear_dgrad(dYt, W["Wt"], residual=dnl)
            # ---- visual projections (InstanceNorm + ReLU)
            dYv = ops.instnorm_bwd(dA3, t["Yv"], Pf["gqkv"], Pf["beqkv"], t["mu3"], t["is3"], Gr["gqkv"], Gr["beqkv"], B, relu=True, queue=self._q)
            self._wgrad(dYv, t["nvc"], Gr["Wqkv"])
            ops.colsum(dYv, Gr["bqkv"], queue=self._q)
            dnv = G.linear_dgrad(dYv, W["Wqkv"], residual=dvp)
        else:
            dnv = dvp
            dnl = ops.batch_sum(dlp, (T, C))
        dvis = ops.l2norm_bwd(dnv, nv, t["inv_v"])
        self._wgrad(dvis, t["X0"], Gr["Wv"])
        ops.colsum(dvis, Gr["bv"], queue=self._q)
        dc4 = G.linear_dgrad(dvis, W["Wv"]).view(B, h, w, -1)
        dlan = ops.l2norm_bwd(dnl, nl, t["inv_l"])
        self._wgrad(dlan, t["hidden"], Gr["Wl"])
        ops.colsum(dlan, Gr["bl"], queue=self._q)
        dhidden = G.linear_dgrad(dlan, W["Wl"])
        self._q.flush()
        return dc4, dhidden

    @staticmethod
    def _wgrad_ex(dy, ldy, x, ldx, out, n, k, m):
        """out[n,k] (fp32 view, pre-zeroed, row stride out.stride(0)) += dy[m,n]^T x[m,k] with explicit strides."""
        tiles = ((n + 127) // 128) * ((k + 127) // 128)
        sk = G._split_for(tiles, (m + 63) // 64)
        G.gemm_ex(dy, x, out, n, k, m, a_mode=MN2D, b_mode=MN2D, lda=ldy, ldb=ldx, ldd=out.stride(0), atomic=1, split_k=sk,
                  block_n=128)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, c4, hidden, head, img_size):
        out, tape = head._fwd(c4, hidden, img_size, True, save=True)
        ctx.head, ctx.tape, ctx.fid = head, tape, head.eng.fwd_id
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, dcls, dfg, drelu, dsig, des):
        head = ctx.head
        head.eng.begin_backward(ctx.fid)
        c = lambda g: None if g is None else g.contiguous()
        dc4, dhidden = head._bwd(ctx.tape, c(dcls), c(dfg), c(drelu), c(dsig), des)
        ctx.tape = None
        return dc4, dhidden, None, None
