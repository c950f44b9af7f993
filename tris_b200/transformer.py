"""CLIP transformer towers on hand-written kernels: text encoder (causal, L = max_query_len) and ViT-B/32 image
encoder, forward and backward (parameter gradients optional: the auxiliary CLIP of train_stage1.py:164-168 is frozen,
so its backward is data-gradient only -- SURVEY K10).

Restates CLIP/clip/model.py:366-386 (ResidualAttentionBlock), :552-564 (encode_text), :419-448 (VisionTransformer).
Linears are tcgen05 GEMMs with bias / QuickGELU / residual fused in the epilogue; the QuickGELU derivative is fused
into the c_proj data-gradient GEMM; LayerNorm, softmax-attention (L <= 64 fits one CTA) are transformer.cu kernels.
Activations bf16, statistics fp32.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from . import gemm as G
from . import ops

bf16, f32 = torch.bfloat16, torch.float32

# The residual stream between transformer blocks is kept in fp32 (forward values; gradients stay bf16): under torch.autocast the
# reference does the same -- fp32 embeddings + bf16 branch outputs promote to fp32 (CLIP/clip/model.py:383-386) -- and the tensors
# are small ([960, 512] / [2400, 768]).  Measured: the bf16 stream put 1.0 % noise on the sentence features, which moved the
# classification loss by up to -0.8 % (profiles/r2_loss_bias_split.txt).  TRIS_RESIDUAL_F32=0 restores the bf16 stream.
# The frozen ViT-B/32 (on the critical path of the step; its features only enter the two small CLIP-guided terms, which deviate
# by < 0.1 % either way) keeps the bf16 stream unless TRIS_VIT_RESIDUAL_F32=1.
STREAM_DTYPE = f32 if os.environ.get("TRIS_RESIDUAL_F32", "1") != "0" else bf16
VIT_STREAM_DTYPE = f32 if os.environ.get("TRIS_VIT_RESIDUAL_F32", "0") == "1" else bf16


class TransformerStack:
    def __init__(self, store, prefix: str, width: int, heads: int, layers: int, causal: bool):
        self.st, self.prefix, self.width, self.heads, self.layers, self.causal = store, prefix, width, heads, layers, causal

    def _k(self, i, name):
        return f"{self.prefix}.resblocks.{i}.{name}"

    def forward(self, x, n, l, save: bool):
        st = self.st
        tape = [] if save else None
        for i in range(self.layers):
            k = lambda nm: self._k(i, nm)
            h, m1, r1 = ops.layernorm_fwd(x, st.p(k("ln_1.weight")), st.p(k("ln_1.bias")), save)
            qkv = G.linear_fwd(h, st.s(k("attn.in_proj_weight")), st.p(k("attn.in_proj_bias")))
            a = ops.attn_fwd(qkv, n, l, self.heads, self.causal)
            x1 = G.linear_fwd(a, st.s(k("attn.out_proj.weight")), st.p(k("attn.out_proj.bias")), residual=x)
            h2, m2, r2 = ops.layernorm_fwd(x1, st.p(k("ln_2.weight")), st.p(k("ln_2.bias")), save)
            pre = torch.empty((x.shape[0], 4 * self.width), device=x.device, dtype=bf16) if save else None
            u = G.linear_fwd(h2, st.s(k("mlp.c_fc.weight")), st.p(k("mlp.c_fc.bias")), act=L.ACT_QUICKGELU, d_pre=pre)
            x2 = G.linear_fwd(u, st.s(k("mlp.c_proj.weight")), st.p(k("mlp.c_proj.bias")), residual=x1)
            if save:
                tape.append((x, m1, r1, h, qkv, a, x1, m2, r2, h2, pre, u))
            x = x2
        return x, tape

    def backward(self, tape, dx, n, l, param_grads: bool, queue=None):
        """queue: gemm.SplitKQueue collecting the split-K second stages of the weight gradients (the caller flushes it)."""
        st = self.st
        for i in reversed(range(self.layers)):
            k = lambda nm: self._k(i, nm)
            x, m1, r1, h, qkv, a, x1, m2, r2, h2, pre, u = tape[i]
            g = (lambda nm: st.g(k(nm))) if param_grads else (lambda nm: None)
            if param_grads:
                G.linear_wgrad(dx, u, out=g("mlp.c_proj.weight"), accumulate=True, queue=queue)
                ops.colsum(dx, g("mlp.c_proj.bias"), queue=queue)
            dpre = G.linear_dgrad(dx, st.s(k("mlp.c_proj.weight")), dact_src=pre, act=L.ACT_QUICKGELU)
            if param_grads:
                G.linear_wgrad(dpre, h2, out=g("mlp.c_fc.weight"), accumulate=True, queue=queue)
                ops.colsum(dpre, g("mlp.c_fc.bias"), queue=queue)
            dh2 = G.linear_dgrad(dpre, st.s(k("mlp.c_fc.weight")))
            dx1 = ops.layernorm_bwd(dh2, x1, st.p(k("ln_2.weight")), m2, r2, add=dx, dgamma=g("ln_2.weight"), dbeta=g("ln_2.bias"), queue=queue)
            if param_grads:
                G.linear_wgrad(dx1, a, out=g("attn.out_proj.weight"), accumulate=True, queue=queue)
                ops.colsum(dx1, g("attn.out_proj.bias"), queue=queue)
            da = G.linear_dgrad(dx1, st.s(k("attn.out_proj.weight")))
            dqkv = ops.attn_bwd(qkv, da, n, l, self.heads, self.causal)
            if param_grads:
                G.linear_wgrad(dqkv, h, out=g("attn.in_proj_weight"), accumulate=True, queue=queue)
                ops.colsum(dqkv, g("attn.in_proj_bias"), queue=queue)
            dh = G.linear_dgrad(dqkv, st.s(k("attn.in_proj_weight")))
            dx = ops.layernorm_bwd(dh, x, st.p(k("ln_1.weight")), m1, r1, add=dx1, dgamma=g("ln_1.weight"), dbeta=g("ln_1.bias"), queue=queue)
        return dx


class TextTower:
    """ids [N, L] int32 -> hidden [N, E] bf16 (EOT-pooled, projected)."""

    def __init__(self, store, prefix: str, width=512, heads=8, layers=12):
        self.st, self.prefix, self.width = store, prefix, width
        self.stack = TransformerStack(store, prefix + "transformer", width, heads, layers, causal=True)

    def forward(self, ids, save: bool, full_sequence: bool = False):
        st, p = self.st, self.prefix
        n, l = ids.shape
        ids = ids.to(torch.int32).contiguous()
        x0, eot = ops.embed_fwd(ids, st.p(p + "token_embedding.weight"), st.p(p + "positional_embedding"), out_dtype=STREAM_DTYPE)
        x, tape = self.stack.forward(x0, n, l, save)
        xe = ops.gather_rows(x, eot)
        xn, m, r = ops.layernorm_fwd(xe, st.p(p + "ln_final.weight"), st.p(p + "ln_final.bias"), save)
        hidden = G.linear_dgrad(xn, st.s(p + "text_projection"))          # [N,W] @ [W,E]
        seq = None
        if full_sequence:
            seq = ops.layernorm_fwd(x, st.p(p + "ln_final.weight"), st.p(p + "ln_final.bias"), False)[0].view(n, l, -1)
        return hidden, ((ids, eot, tape, xe, xn, m, r, x.shape[0]) if save else None), seq

    def backward(self, rec, dhidden):
        st, p = self.st, self.prefix
        ids, eot, tape, xe, xn, m, r, rows = rec
        n, l = ids.shape
        q = G.SplitKQueue()
        G.linear_wgrad(xn, dhidden, out=st.g(p + "text_projection"), accumulate=True, queue=q)
        dxn = G.linear_fwd(dhidden, st.s(p + "text_projection"))         # [N,E] @ [W,E]^T
        dxe = ops.layernorm_bwd(dxn, xe, st.p(p + "ln_final.weight"), m, r, dgamma=st.g(p + "ln_final.weight"),
                                dbeta=st.g(p + "ln_final.bias"), queue=q)
        dx = ops.scatter_rows(dxe, eot, rows)
        dx0 = self.stack.backward(tape, dx, n, l, True, queue=q)
        q.flush()          # one launch: all split-K weight-gradient partials of the tower, added in split order
        ops.embed_bwd(ids, dx0, st.g(p + "token_embedding.weight"), st.g(p + "positional_embedding"))


class VitTower:
    """patches [N*G, 3*ps*ps] bf16 (k = c*ps*ps + py*ps + px) -> features [N, E] bf16.  Frozen: dgrad-only backward."""

    def __init__(self, store, prefix: str = "visual.", width=768, heads=12, layers=12, grid=7):
        self.st, self.prefix, self.width, self.tokens = store, prefix, width, grid * grid + 1
        self.stack = TransformerStack(store, prefix + "transformer", width, heads, layers, causal=False)
        dev = store.device
        self._idx_cache = {}

    def _idx(self, n, dev):
        if n not in self._idx_cache:
            t = self.tokens
            cls = torch.arange(n, device=dev, dtype=torch.int32) * t
            allr = torch.arange(n * t, device=dev, dtype=torch.int32)
            self._idx_cache[n] = (cls, allr[allr % t != 0].contiguous())
        return self._idx_cache[n]

    def forward(self, patches, n, save: bool):
        st, p = self.st, self.prefix
        w = st.s(p + "conv1.weight")
        pe = G.linear_fwd(patches, w.view(w.shape[0], -1))
        tok = ops.vit_assemble(pe, st.p(p + "class_embedding"), st.p(p + "positional_embedding"), n)
        x0, m0, r0 = ops.layernorm_fwd(tok, st.p(p + "ln_pre.weight"), st.p(p + "ln_pre.bias"), save, out_dtype=VIT_STREAM_DTYPE)
        x, tape = self.stack.forward(x0, n, self.tokens, save)
        cls_idx, patch_idx = self._idx(n, patches.device)
        xc = ops.gather_rows(x, cls_idx)
        xn, m, r = ops.layernorm_fwd(xc, st.p(p + "ln_post.weight"), st.p(p + "ln_post.bias"), save)
        feat = G.linear_dgrad(xn, st.s(p + "proj"))                      # [N,W] @ [W,E]
        return feat, ((tok, m0, r0, tape, xc, m, r, n) if save else None)

    def backward(self, rec, dfeat):
        st, p = self.st, self.prefix
        tok, m0, r0, tape, xc, m, r, n = rec
        cls_idx, patch_idx = self._idx(n, dfeat.device)
        dxn = G.linear_fwd(dfeat, st.s(p + "proj"))
        dxc = ops.layernorm_bwd(dxn, xc, st.p(p + "ln_post.weight"), m, r)
        dx = ops.scatter_rows(dxc, cls_idx, n * self.tokens)
        dx0 = self.stack.backward(tape, dx, n, self.tokens, False)
        dtok = ops.layernorm_bwd(dx0, tok, st.p(p + "ln_pre.weight"), m0, r0)
        dpe = ops.gather_rows(dtok, patch_idx)
        w = st.s(p + "conv1.weight")
        return G.linear_dgrad(dpe, w.view(w.shape[0], -1))                # [N*G, 3*ps*ps]
