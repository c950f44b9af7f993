"""Functional fp32 restatement of the TRIS Stage-1 hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

Every function works on a flat ``state_dict``-style mapping (no nn.Module) and
cites the reference lines it restates (paths relative to the reference root,
fawnliu/TRIS @ c6666b3).  Arithmetic bottoms out in torch fp32 CPU ops, the same
third-party library the reference itself uses (torch; reference pins 1.8.0,
README >= 1.13.1, this image 2.11.0 -- SURVEY 8c).

Pinned by tests/test_oracle.py against tests/golden/*.npz, which were produced
by the unmodified reference modules (tests/golden/make_golden.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------
# CLIP building blocks
# ----------------------------------------------------------------------------
def layer_norm(x, sd: SD, p: str):
    """CLIP/clip/model.py:352-358 -- fp32 LayerNorm, eps 1e-5."""
    return F.layer_norm(x.float(), (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def residual_attention_block(x, sd: SD, p: str, heads: int, causal: bool):
    """CLIP/clip/model.py:366-386 (+ nn.MultiheadAttention semantics).
    x: [N, L, D] batch-first (the reference runs sequence-first; same math)."""
    n, l, d = x.shape
    hd = d // heads
    h = layer_norm(x, sd, p + ".ln_1")
    qkv = h @ sd[p + ".attn.in_proj_weight"].t() + sd[p + ".attn.in_proj_bias"]
    q, k, v = qkv.split(d, dim=-1)
    q = q.reshape(n, l, heads, hd).transpose(1, 2)
    k = k.reshape(n, l, heads, hd).transpose(1, 2)
    v = v.reshape(n, l, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if causal:  # model.py:537-543: -inf strictly above the diagonal
        s = s + torch.full((l, l), float("-inf"), dtype=s.dtype, device=s.device).triu_(1)
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(n, l, d)
    x = x + a @ sd[p + ".attn.out_proj.weight"].t() + sd[p + ".attn.out_proj.bias"]
    u = layer_norm(x, sd, p + ".ln_2") @ sd[p + ".mlp.c_fc.weight"].t() + sd[p + ".mlp.c_fc.bias"]
    u = u * torch.sigmoid(1.702 * u)  # QuickGELU, model.py:361-363
    return x + u @ sd[p + ".mlp.c_proj.weight"].t() + sd[p + ".mlp.c_proj.bias"]


def encode_text(sd: SD, ids, prefix: str = "", heads: int = 8, layers: int = 12):
    """CLIP.encode_text, CLIP/clip/model.py:552-564 -> (x [N,L,W], hidden [N,E])."""
    ids = ids.long()
    x = sd[prefix + "token_embedding.weight"][ids] + sd[prefix + "positional_embedding"][: ids.shape[1]]
    for i in range(layers):
        x = residual_attention_block(x, sd, f"{prefix}transformer.resblocks.{i}", heads, causal=True)
    x = layer_norm(x, sd, prefix + "ln_final")
    eot = ids.argmax(dim=-1)
    hidden = x[torch.arange(x.shape[0]), eot] @ sd[prefix + "text_projection"]
    return x, hidden


def batch_norm(x, sd: SD, p: str, train: bool, new_stats: Optional[SD]):
    """nn.BatchNorm2d (eps 1e-5, momentum 0.1): batch stats + running update in train
    mode, running stats in eval.  Running-stat updates are written to
    ``new_stats`` instead of mutating ``sd``."""
    if not train:
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                            False, 0.1, 1e-5)
    rm, rv = sd[p + ".running_mean"].clone(), sd[p + ".running_var"].clone()
    y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, 0.1, 1e-5)
    if new_stats is not None:
        new_stats[p + ".running_mean"], new_stats[p + ".running_var"] = rm, rv
        new_stats[p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
    return y


def bottleneck(x, sd: SD, p: str, stride: int, train: bool, new_stats):
    """CLIP/clip/model.py:10-55."""
    o = F.relu(batch_norm(F.conv2d(x, sd[p + ".conv1.weight"]), sd, p + ".bn1", train, new_stats))
    o = F.relu(batch_norm(F.conv2d(o, sd[p + ".conv2.weight"], padding=1), sd, p + ".bn2", train, new_stats))
    if stride > 1:
        o = F.avg_pool2d(o, stride)
    o = batch_norm(F.conv2d(o, sd[p + ".conv3.weight"]), sd, p + ".bn3", train, new_stats)
    if (p + ".downsample.0.weight") in sd:
        idt = F.avg_pool2d(x, stride) if stride > 1 else x
        idt = batch_norm(F.conv2d(idt, sd[p + ".downsample.0.weight"]), sd, p + ".downsample.1", train, new_stats)
    else:
        idt = x
    return F.relu(o + idt)


def resnet_tower(sd: SD, x, prefix: str = "visual.", layers=(3, 4, 6, 3), train: bool = False, new_stats=None):
    """ModifiedResNet.forward, CLIP/clip/model.py:254-279, WITHOUT the attention pool
    (its output is discarded by model_stage1.py:59).  Returns (c1,c2,c3,c4)."""
    for i in (1, 2, 3):
        x = F.conv2d(x, sd[f"{prefix}conv{i}.weight"], stride=2 if i == 1 else 1, padding=1)
        x = F.relu(batch_norm(x, sd, f"{prefix}bn{i}", train, new_stats))
    x = F.avg_pool2d(x, 2)
    outs = []
    for li, blocks in enumerate(layers, start=1):
        for b in range(blocks):
            x = bottleneck(x, sd, f"{prefix}layer{li}.{b}", 2 if (b == 0 and li > 1) else 1, train, new_stats)
        outs.append(x)
    return tuple(outs)


def attention_pool(sd: SD, c4, prefix: str = "visual.attnpool.", heads: int = 32):
    """AttentionPool2d.forward, CLIP/clip/model.py:70-104 (dead compute in Stage-1; kept for
    encode_image API parity).  Returns (global [B,E], local [B,E,H,W])."""
    b, c, h, w = c4.shape
    x = c4.reshape(b, c, h * w).permute(0, 2, 1)
    x = torch.cat([x.mean(dim=1, keepdim=True), x], dim=1)
    pe = sd[prefix + "positional_embedding"]
    sp = int(round((pe.shape[0] - 1) ** 0.5))
    spatial = F.interpolate(pe[1:].reshape(1, sp, sp, c).permute(0, 3, 1, 2), size=(h, w), mode="bilinear")
    pos = torch.cat([pe[:1], spatial.reshape(c, h * w).t()], dim=0)
    x = x + pos[None]
    hd = c // heads
    q = (x @ sd[prefix + "q_proj.weight"].t() + sd[prefix + "q_proj.bias"]).reshape(b, -1, heads, hd).transpose(1, 2)
    k = (x @ sd[prefix + "k_proj.weight"].t() + sd[prefix + "k_proj.bias"]).reshape(b, -1, heads, hd).transpose(1, 2)
    v = (x @ sd[prefix + "v_proj.weight"].t() + sd[prefix + "v_proj.bias"]).reshape(b, -1, heads, hd).transpose(1, 2)
    a = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1) @ v
    a = a.transpose(1, 2).reshape(b, -1, c)
    o = a @ sd[prefix + "c_proj.weight"].t() + sd[prefix + "c_proj.bias"]
    return o[:, 0], o[:, 1:].transpose(1, 2).reshape(b, -1, h, w)


def vit_tower(sd: SD, x, prefix: str = "visual.", heads: int = 12, layers: int = 12):
    """VisionTransformer.forward, CLIP/clip/model.py:419-448 (ViT-B/32) -> [N, 512]."""
    w = sd[prefix + "conv1.weight"]
    x = F.conv2d(x, w, stride=w.shape[-1])
    n, d = x.shape[:2]
    x = x.reshape(n, d, -1).permute(0, 2, 1)
    cls = sd[prefix + "class_embedding"].expand(n, 1, d)
    x = torch.cat([cls, x], dim=1) + sd[prefix + "positional_embedding"]
    x = layer_norm(x, sd, prefix + "ln_pre")
    for i in range(layers):
        x = residual_attention_block(x, sd, f"{prefix}transformer.resblocks.{i}", heads, causal=False)
    x = layer_norm(x[:, 0, :], sd, prefix + "ln_post")
    return x @ sd[prefix + "proj"]


# ----------------------------------------------------------------------------
# TRIS Stage-1
# ----------------------------------------------------------------------------
def instance_norm(x_bpc, w, b):
    """nn.InstanceNorm2d(affine=True): per (sample, channel) stats over the pixels,
    biased variance, eps 1e-5, identical in train/eval.  x: [B, P, C]."""
    mu = x_bpc.mean(dim=1, keepdim=True)
    var = x_bpc.var(dim=1, unbiased=False, keepdim=True)
    return (x_bpc - mu) / torch.sqrt(var + 1e-5) * w + b


def bilateral_prompt(sd: SD, vis_bpc, lan_btc, p: str = "attn_fusion."):
    """model/attn.py:111-136 on channel-last views.  vis [B,P,C] (pixels), lan [B,T,C].
    Returns (new_vis [B,P,C], new_lan [B,T,C])."""
    c = vis_bpc.shape[-1]

    def vproj(i):
        w = sd[f"{p}v_proj{i}.0.weight"].reshape(c, -1)
        y = vis_bpc @ w.t() + sd[f"{p}v_proj{i}.0.bias"]
        return F.relu(instance_norm(y, sd[f"{p}v_proj{i}.1.weight"], sd[f"{p}v_proj{i}.1.bias"]))

    def tproj(i):
        return F.relu(lan_btc @ sd[f"{p}t_proj{i}.0.weight"].t() + sd[f"{p}t_proj{i}.0.bias"])

    qv, kv, vv = vproj(1), vproj(2), vproj(3)
    qt, kt, vt = tproj(1), tproj(2), tproj(3)
    av = torch.softmax(qv @ kt.transpose(1, 2) / math.sqrt(c), dim=2)      # [B,P,T]  attn.py:122
    at = torch.softmax(qt @ kv.transpose(1, 2) / math.sqrt(c), dim=2)      # [B,T,P]  attn.py:125
    new_vis = av @ vt                                                     # [B,P,C]
    new_lan = at @ vv                                                     # [B,T,C]
    wo = sd[p + "v_output.0.weight"].reshape(c, -1)
    new_vis = instance_norm(new_vis @ wo.t() + sd[p + "v_output.0.bias"],
                            sd[p + "v_output.1.weight"], sd[p + "v_output.1.bias"])
    new_lan = new_lan @ sd[p + "t_output.0.weight"].t() + sd[p + "t_output.0.bias"]
    return new_vis, new_lan


def tris_score(sd: SD, c4, hidden, attn_multi: float = 0.1):
    """model/model_stage1.py:61-78 -> score [B, P, T] (already times exp(logit_scale))."""
    b = c4.shape[0]
    lan = hidden @ sd["lan_project.weight"].t() + sd["lan_project.bias"]
    wv = sd["vis_project.weight"].reshape(sd["vis_project.weight"].shape[0], -1)
    vis = c4.flatten(2).transpose(1, 2) @ wv.t() + sd["vis_project.bias"]          # [B,P,C]
    lan = lan.unsqueeze(0).repeat(b, 1, 1)                                          # [B,T,C]
    nv = vis / vis.norm(dim=-1, keepdim=True)
    nl = lan / lan.norm(dim=-1, keepdim=True)
    if attn_multi > 0:
        new_vis, new_lan = bilateral_prompt(sd, nv, nl)
        nv = new_vis * 0.1 + nv
        nl = new_lan * 0.1 + nl
    return sd["logit_scale"].exp() * torch.bmm(nv, nl.transpose(1, 2))


def tris_head(score, hw, img_size, train: bool, focal_p: float = 3.0, focal_lambda: float = 0.01):
    """model/model_stage1.py:80-119.  score [B,P,T]."""
    b, p, t = score.shape
    h_, w_ = hw
    outs = {}
    if train:
        feat = torch.cat([torch.ones_like(score[:, :, :1]), score], dim=2).transpose(1, 2)   # [B,T+1,P]
        masks = torch.softmax(feat, dim=1)
        cls1 = feat.mean(-1) + feat.max(dim=-1).values
        m = masks.mean(-1)
        cls2 = torch.pow(1 - m, focal_p) * torch.log(focal_lambda + m)
        outs["cls_out"] = cls1[:, 1:] + cls2[:, 1:]
        outs["cls_fg"] = torch.diagonal(m[:, 1:], dim1=0, dim2=1).clone()
    idx = torch.arange(b)
    maps = score[idx, :, idx].reshape(b, 1, h_, w_)
    seg = F.interpolate(maps, size=img_size, mode="bilinear", align_corners=False)       # model/utils.py:5-10
    outs["maps10"] = maps
    outs["relu"] = F.relu(seg)
    outs["sig"] = torch.sigmoid(seg)
    return outs


def tris_forward(sd: SD, x, word_id, train: bool, new_stats=None, attn_multi=0.1, focal_p=3.0, focal_lambda=0.01):
    """TRIS.forward, model/model_stage1.py:54-119.
    train -> (cls_out, cls_fg, relu_map, sig_map, exp(logit_scale)); eval -> relu_map."""
    _, hidden = encode_text(sd, word_id, prefix="backbone.")
    c4 = resnet_tower(sd, x, prefix="backbone.visual.", train=train, new_stats=new_stats)[-1]
    score = tris_score(sd, c4, hidden, attn_multi)
    o = tris_head(score, c4.shape[2:], x.shape[2:], train, focal_p, focal_lambda)
    if train:
        return o["cls_out"], o["cls_fg"], o["relu"], o["sig"], sd["logit_scale"].exp()
    return o["relu"]


# ----------------------------------------------------------------------------
# Training-step glue (train_stage1.py:320-364)
# ----------------------------------------------------------------------------
def mask_and_resize(sig_out, img, size: int = 224):
    """train_stage1.py:327-339: fg = bilinear_ac(sig->224) * bilinear_ac(img->224)."""
    if img.shape[2] != size:
        cam = F.interpolate(sig_out, (size, size), mode="bilinear", align_corners=True)
        im = F.interpolate(img, (size, size), mode="bilinear", align_corners=True)
    else:
        cam, im = sig_out, img
    return cam * im


def _unit(x):
    return x / x.norm(dim=-1, keepdim=True)


def stage1_losses(cls_out, sig_out, img, word_ids, neg_word_ids, aux: SD, w1=1.0, w4=5.0, w5=2.0):
    """train_stage1.py:320-364 -> dict(loss, l1, l4, l5, fg).  aux = ViT-B/32 CLIP state dict."""
    b = img.shape[0]
    fg = mask_and_resize(sig_out, img)
    f = _unit(vit_tower(aux, fg))                                         # clip_forward :263-278
    g = _unit(encode_text(aux, word_ids)[1])
    cos = (f * g).sum(-1)
    l1 = -torch.log(cos.clamp(0.0001, 0.9999)).mean()                      # MaxLoss :280-284
    if neg_word_ids is not None:
        k = neg_word_ids.shape[1]
        nfeat = _unit(encode_text(aux, neg_word_ids.reshape(b * k, -1))[1]).reshape(b, k, -1)
        nscore = torch.einsum("bc,bkc->bk", f, nfeat)
        l5 = (-torch.log(1 - nscore)).mean(dim=1).sum() / b                # :345-353
    else:
        l5 = torch.zeros((), dtype=img.dtype, device=img.device)
    l4 = F.multilabel_soft_margin_loss(cls_out, torch.eye(b, dtype=cls_out.dtype, device=cls_out.device))  # :354
    return {"loss": l1 * w1 + l4 * w4 + l5 * w5, "l1": l1, "l4": l4, "l5": l5, "fg": fg}


NO_GRAD_KEYS = ("backbone.visual.attnpool.", "backbone.logit_scale")


def trainable_keys(sd: SD):
    """Keys that receive a gradient in the Stage-1 step: everything float
    except attnpool.*, backbone.logit_scale and BN running stats."""
    out = []
    for k, v in sd.items():
        if not v.is_floating_point() or "running_" in k:
            continue
        if any(k.startswith(n) for n in NO_GRAD_KEYS):
            continue
        out.append(k)
    return out


def train_step(sd: SD, aux: SD, img, word_ids, neg_word_ids, w=(1.0, 5.0, 2.0)):
    """forward + the three losses + backward.  Returns (losses, grads{key}, new_stats, fwd outputs)."""
    keys = trainable_keys(sd)
    leaf = dict(sd)
    for k in keys:
        leaf[k] = sd[k].detach().clone().requires_grad_(True)
    new_stats: SD = {}
    cls_out, cls_fg, relu_map, sig_map, ls = tris_forward(leaf, img, word_ids, True, new_stats)
    with torch.no_grad():
        aux = {k: v.detach() for k, v in aux.items()}
    losses = stage1_losses(cls_out, sig_map, img, word_ids, neg_word_ids, aux, *w)
    grads = torch.autograd.grad(losses["loss"], [leaf[k] for k in keys], allow_unused=True)
    gd = {k: (g if g is not None else torch.zeros_like(sd[k])) for k, g in zip(keys, grads)}
    fwd = {"cls_out": cls_out.detach(), "cls_fg": cls_fg.detach(), "relu": relu_map.detach(), "sig": sig_map.detach()}
    return {k: v.detach() for k, v in losses.items()}, gd, new_stats, fwd


# ----------------------------------------------------------------------------
# Optimizer (train_stage1.py:133-144, 368-372)
# ----------------------------------------------------------------------------
def poly_lr(it: int, max_it: int, power: float = 0.9) -> float:
    """LambdaLR factor (1 - it/max_it)^0.9, train_stage1.py:143-144."""
    return (1.0 - it / max_it) ** power


def adamw_step(p, g, m, v, step: int, lr: float, wd: float = 0.01, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.AdamW single-tensor math (decoupled weight decay), in place on clones."""
    p = p * (1 - lr * wd)
    m = m * b1 + g * (1 - b1)
    v = v * b2 + g * g * (1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    p = p - (lr / bc1) * m / (v.sqrt() / math.sqrt(bc2) + eps)
    return p, m, v


def param_group_of(key: str) -> int:
    """0 = backbone group (lr*lr_multi), 1 = new modules (model_stage1.py:44-52), -1 = none
    (TRIS.logit_scale is in neither group)."""
    if key.startswith("backbone."):
        return 0
    if key.startswith(("vis_project.", "lan_project.", "attn_fusion.")):
        return 1
    return -1
