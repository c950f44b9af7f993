"""Stage-1 head: vis/lan projection + L2-norm (K6), bilateral cross-modal attention + mix + score (K7), response head
(K8).  Restates model/model_stage1.py:61-119 and model/attn.py:111-136.

ROUND-1 STATUS: the dense projections run on the tcgen05 GEMM (gemm.py); the per-image attention core, InstanceNorm
and the classification reductions are still composed from torch CUDA ops inside ``_HeadFn`` (recompute-in-backward),
to be replaced by the fused kernels listed in DESIGN.md's coverage table.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

bf16, f32 = torch.bfloat16, torch.float32


def _inorm(x, w, b):
    mu = x.mean(dim=1, keepdim=True)
    var = x.var(dim=1, unbiased=False, keepdim=True)
    return (x - mu) * torch.rsqrt(var + 1e-5) * w + b


def _mm(a, w, b=None):
    """bf16 tensor-core matmul with fp32 result: a[..., K] @ w[N, K]^T (+ b)."""
    y = F.linear(a.to(bf16), w.to(bf16)).float()
    return y if b is None else y + b


class Stage1Head:
    KEYS = ["vis_project.weight", "vis_project.bias", "lan_project.weight", "lan_project.bias", "logit_scale"]

    def __init__(self, eng, model):
        self.eng = eng
        self.fuse = hasattr(model, "attn_fusion") and model.args.attn_multi > 0
        self.keys = list(self.KEYS)
        if self.fuse:
            self.keys += [k for k in eng.store.offsets if k.startswith("attn_fusion.")]
        self.focal_p, self.focal_lambda = float(model.args.FOCAL_P), float(model.args.FOCAL_LAMBDA)

    def params(self):
        return {k: self.eng.store.params[k] for k in self.keys}

    # --------------------------------------------------------------- math (differentiable torch ops; interim)
    def score(self, P, c4, hidden):
        B, h, w, cv = c4.shape
        C = P["vis_project.weight"].shape[0]
        vis = _mm(c4.reshape(B, h * w, cv), P["vis_project.weight"].reshape(C, cv), P["vis_project.bias"])
        lan = _mm(hidden, P["lan_project.weight"], P["lan_project.bias"])
        nv = vis / vis.norm(dim=-1, keepdim=True)
        nl = (lan / lan.norm(dim=-1, keepdim=True)).unsqueeze(0).expand(B, -1, -1)
        if self.fuse:
            a = "attn_fusion."
            vp = [F.relu(_inorm(_mm(nv, P[f"{a}v_proj{i}.0.weight"].reshape(C, C), P[f"{a}v_proj{i}.0.bias"]),
                                P[f"{a}v_proj{i}.1.weight"], P[f"{a}v_proj{i}.1.bias"])) for i in (1, 2, 3)]
            tp = [F.relu(_mm(nl[0], P[f"{a}t_proj{i}.0.weight"], P[f"{a}t_proj{i}.0.bias"])) for i in (1, 2, 3)]
            qv, kv, vv = vp
            qt, kt, vt = tp
            av = torch.softmax(_mm(qv, kt) / math.sqrt(C), dim=2)                       # [B,P,T]
            at = torch.softmax(torch.einsum("tc,bpc->btp", qt.to(bf16), kv.to(bf16)).float() / math.sqrt(C), dim=2)
            new_vis = torch.einsum("bpt,tc->bpc", av.to(bf16), vt.to(bf16)).float()
            new_lan = torch.bmm(at.to(bf16), vv.to(bf16)).float()
            new_vis = _inorm(_mm(new_vis, P[a + "v_output.0.weight"].reshape(C, C), P[a + "v_output.0.bias"]),
                             P[a + "v_output.1.weight"], P[a + "v_output.1.bias"])
            new_lan = _mm(new_lan, P[a + "t_output.0.weight"], P[a + "t_output.0.bias"])
            nv = new_vis * 0.1 + nv
            nl = new_lan * 0.1 + nl
        return P["logit_scale"].exp() * torch.bmm(nv, nl.transpose(1, 2))                 # fp32 [B,P,T]

    def outputs(self, score, hw, img_size, training):
        B, Pn, T = score.shape
        res = {}
        if training:
            feat = torch.cat([torch.ones_like(score[:, :, :1]), score], dim=2).transpose(1, 2)
            masks = torch.softmax(feat, dim=1)
            m = masks.mean(-1)
            cls = feat.mean(-1) + feat.max(dim=-1).values + torch.pow(1 - m, self.focal_p) * torch.log(self.focal_lambda + m)
            res["cls_out"] = cls[:, 1:]
            res["cls_fg"] = torch.diagonal(m[:, 1:], dim1=0, dim2=1)
        idx = torch.arange(B, device=score.device)
        maps = score[idx, :, idx].reshape(B, 1, *hw)
        seg = F.interpolate(maps, size=img_size, mode="bilinear", align_corners=False)
        res["maps"] = maps
        res["relu"] = F.relu(seg)
        if training:
            res["sig"] = torch.sigmoid(seg)
        return res

    def _run(self, c4, hidden, img_size, training):
        score = self.score(self.params(), c4, hidden)
        o = self.outputs(score, c4.shape[1:3], img_size, training)
        if training:
            return o["cls_out"], o["cls_fg"], o["relu"], o["sig"], self.eng.store.params["logit_scale"].exp()
        return (o["relu"],)

    def forward(self, c4, hidden, img_size, training):
        if training and torch.is_grad_enabled():
            return _HeadFn.apply(c4, hidden, self, img_size)
        with torch.no_grad():
            return self._run(c4, hidden, img_size, training)


class _HeadFn(torch.autograd.Function):
    """Interim: forward without a graph, backward = recompute under autograd and add the parameter gradients into
    the flat gradient buffer (keeps ``param.grad`` the store's views)."""

    @staticmethod
    def forward(ctx, c4, hidden, head, img_size):
        ctx.head, ctx.img_size, ctx.fid = head, img_size, head.eng.fwd_id
        ctx.save_for_backward(c4, hidden)
        with torch.no_grad():
            return head._run(c4, hidden, img_size, True)

    @staticmethod
    def backward(ctx, *douts):
        head = ctx.head
        eng = head.eng
        eng.begin_backward(ctx.fid)
        c4, hidden = ctx.saved_tensors
        c4 = c4.detach().requires_grad_(True)
        hidden = hidden.detach().requires_grad_(True)
        with torch.enable_grad():
            outs = head._run(c4, hidden, ctx.img_size, True)
        P = head.params()
        keys = list(P.keys())
        pairs = [(o, g) for o, g in zip(outs, douts) if g is not None and o.requires_grad]
        grads = torch.autograd.grad([o for o, _ in pairs], [c4, hidden] + [P[k] for k in keys], [g for _, g in pairs],
                                    allow_unused=True)
        for k, g in zip(keys, grads[2:]):
            if g is not None:
                eng.store.g(k).add_(g)
        return grads[0], grads[1], None, None
