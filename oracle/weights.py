"""Deterministic random weights + synthetic batches for the TRIS Stage-1 path.

Test infrastructure (see oracle/__init__.py).  The key set / shapes reproduce
the reference ``state_dict`` so the same dict loads (strict) into
  * the unmodified reference ``model.model_stage1.TRIS``  (golden generation),
  * ``oracle.tris_oracle`` (functional restatement), and
  * ``tris_b200.model_stage1.TRIS`` (the product).

Shapes follow reference CLIP/clip/model.py:451-506 (CLIP ctor), :194-252
(ModifiedResNet), :58-68 (AttentionPool2d), :366-378 (ResidualAttentionBlock),
:400-417 (VisionTransformer) and model/model_stage1.py:36-42, model/attn.py:69-109.
Init scales follow model.py:508-535 except that every BatchNorm is re-randomised
(reference zero-inits bn3.weight, model.py:520-523, which would hide the whole
residual branch from a parity test -- SURVEY F9).
"""
from __future__ import annotations

from collections import OrderedDict

import torch

RN50_LAYERS = (3, 4, 6, 3)
VOCAB = 49408
SOT, EOT = 49406, 49407


def _g(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _normal(g, shape, std):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _uniform(g, shape, lo, hi):
    return torch.rand(shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def _bn(sd, prefix, c, g):
    sd[prefix + ".weight"] = _uniform(g, (c,), 0.5, 1.5)
    sd[prefix + ".bias"] = _normal(g, (c,), 0.1)
    sd[prefix + ".running_mean"] = _normal(g, (c,), 0.1)
    sd[prefix + ".running_var"] = _uniform(g, (c,), 0.5, 1.5)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.int64)


def _conv(sd, key, cout, cin, k, g):
    fan_in = cin * k * k
    sd[key] = _normal(g, (cout, cin, k, k), (2.0 / fan_in) ** 0.5 * 0.7)


def _linear(sd, prefix, cout, cin, g, std=None, bias=True):
    std = std if std is not None else cin ** -0.5
    sd[prefix + ".weight"] = _normal(g, (cout, cin), std)
    if bias:
        sd[prefix + ".bias"] = _normal(g, (cout,), 0.02)


def _transformer(sd, prefix, width, layers, g):
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    attn_std = width ** -0.5
    fc_std = (2 * width) ** -0.5
    for i in range(layers):
        p = f"{prefix}.resblocks.{i}"
        sd[p + ".attn.in_proj_weight"] = _normal(g, (3 * width, width), attn_std)
        sd[p + ".attn.in_proj_bias"] = _normal(g, (3 * width,), 0.02)
        _linear(sd, p + ".attn.out_proj", width, width, g, proj_std)
        sd[p + ".ln_1.weight"] = _uniform(g, (width,), 0.8, 1.2)
        sd[p + ".ln_1.bias"] = _normal(g, (width,), 0.05)
        _linear(sd, p + ".mlp.c_fc", 4 * width, width, g, fc_std)
        _linear(sd, p + ".mlp.c_proj", width, 4 * width, g, proj_std)
        sd[p + ".ln_2.weight"] = _uniform(g, (width,), 0.8, 1.2)
        sd[p + ".ln_2.bias"] = _normal(g, (width,), 0.05)


def _text_side(sd, prefix, width, embed_dim, g, context_length=77, cos_bias=False):
    """positional_embedding, text_projection, logit_scale, transformer, token_embedding,
    ln_final -- registration order of CLIP.__init__ (model.py:488-504)."""
    sd[prefix + "positional_embedding"] = _normal(g, (context_length, width), 0.01)
    tp = _normal(g, (width, embed_dim), width ** -0.5)
    if cos_bias:
        tp = tp + 0.005
    sd[prefix + "text_projection"] = tp
    sd[prefix + "logit_scale"] = torch.tensor(2.6592600, dtype=torch.float32)  # ln(1/0.07)


def make_rn50_clip_state_dict(seed: int = 0, prefix: str = "", embed_dim: int = 1024) -> "OrderedDict[str, torch.Tensor]":
    """CLIP-RN50 (embed 1024, width 64, layers (3,4,6,3), text width 512 / 8 heads / 12 layers)."""
    g = _g(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _text_side(sd, prefix, 512, embed_dim, g)
    v = prefix + "visual."
    _conv(sd, v + "conv1.weight", 32, 3, 3, g); _bn(sd, v + "bn1", 32, g)
    _conv(sd, v + "conv2.weight", 32, 32, 3, g); _bn(sd, v + "bn2", 32, g)
    _conv(sd, v + "conv3.weight", 64, 32, 3, g); _bn(sd, v + "bn3", 64, g)
    inplanes = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), RN50_LAYERS), start=1):
        for b in range(blocks):
            p = f"{v}layer{li}.{b}."
            stride = 2 if (b == 0 and li > 1) else 1
            _conv(sd, p + "conv1.weight", planes, inplanes, 1, g); _bn(sd, p + "bn1", planes, g)
            _conv(sd, p + "conv2.weight", planes, planes, 3, g); _bn(sd, p + "bn2", planes, g)
            _conv(sd, p + "conv3.weight", planes * 4, planes, 1, g); _bn(sd, p + "bn3", planes * 4, g)
            # residual-branch gain: small but non-zero.  gamma ~ U(0.5,1.5) on every branch makes a random-init
            # train-mode BN ResNet chaotic (perturbations grow ~1.3x per block: bf16 storage noise of 1% becomes
            # 50% at c4 even in the fp32 reference), which says nothing about kernel correctness; the reference's
            # own init uses gamma = 0 here (model.py:520-523), trained checkpoints sit in between.
            sd[p + "bn3.weight"] = _uniform(g, (planes * 4,), 0.1, 0.3)
            if stride > 1 or inplanes != planes * 4:
                _conv(sd, p + "downsample.0.weight", planes * 4, inplanes, 1, g)
                _bn(sd, p + "downsample.1", planes * 4, g)
            inplanes = planes * 4
    a = v + "attnpool."
    sd[a + "positional_embedding"] = _normal(g, (7 * 7 + 1, 2048), 2048 ** -0.5)
    for nm in ("k_proj", "q_proj", "v_proj"):
        _linear(sd, a + nm, 2048, 2048, g)
    _linear(sd, a + "c_proj", embed_dim, 2048, g)
    _transformer(sd, prefix + "transformer", 512, 12, g)
    sd[prefix + "token_embedding.weight"] = _normal(g, (VOCAB, 512), 0.02)
    sd[prefix + "ln_final.weight"] = _uniform(g, (512,), 0.8, 1.2)
    sd[prefix + "ln_final.bias"] = _normal(g, (512,), 0.05)
    return sd


def make_tris_state_dict(seed: int = 0, hidden_dim: int = 1024) -> "OrderedDict[str, torch.Tensor]":
    """Full Stage-1 TRIS state_dict (518 entries)."""
    g = _g(seed + 1000)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    sd["logit_scale"] = torch.tensor(2.6592600, dtype=torch.float32)
    sd.update(make_rn50_clip_state_dict(seed, prefix="backbone."))
    sd["vis_project.weight"] = _normal(g, (hidden_dim, 2048, 1, 1), 2048 ** -0.5)
    sd["vis_project.bias"] = _normal(g, (hidden_dim,), 0.02)
    _linear(sd, "lan_project", hidden_dim, 1024, g)
    for nm in ("v_proj1", "v_proj2", "v_proj3"):
        sd[f"attn_fusion.{nm}.0.weight"] = _normal(g, (hidden_dim, hidden_dim, 1, 1), 1.0)
        sd[f"attn_fusion.{nm}.0.bias"] = _normal(g, (hidden_dim,), 0.02)
        sd[f"attn_fusion.{nm}.1.weight"] = _uniform(g, (hidden_dim,), 0.5, 1.5)
        sd[f"attn_fusion.{nm}.1.bias"] = _normal(g, (hidden_dim,), 0.3)
    for nm in ("t_proj1", "t_proj2", "t_proj3"):
        _linear(sd, f"attn_fusion.{nm}.0", hidden_dim, hidden_dim, g, std=1.0)
    sd["attn_fusion.v_output.0.weight"] = _normal(g, (hidden_dim, hidden_dim, 1, 1), hidden_dim ** -0.5)
    sd["attn_fusion.v_output.0.bias"] = _normal(g, (hidden_dim,), 0.02)
    sd["attn_fusion.v_output.1.weight"] = _uniform(g, (hidden_dim,), 0.5, 1.5)
    sd["attn_fusion.v_output.1.bias"] = _normal(g, (hidden_dim,), 0.1)
    _linear(sd, "attn_fusion.t_output.0", hidden_dim, hidden_dim, g)
    return sd


def make_vitb32_clip_state_dict(seed: int = 7, cos_bias: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Aux CLIP ViT-B/32 (embed 512, width 768, patch 32, 12 layers; text 512/8/12).

    ``cos_bias`` gives both feature heads a shared positive direction so that
    cos(image, text) lands inside the clamp window (1e-4, 0.9999) of
    ``MaxLoss`` (train_stage1.py:280-284); with pure zero-mean random weights the
    fg loss saturates at -log(1e-4) and has zero gradient.
    """
    g = _g(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _text_side(sd, "", 512, 512, g, cos_bias=cos_bias)
    v = "visual."
    scale = 768 ** -0.5
    sd[v + "class_embedding"] = _normal(g, (768,), scale)
    sd[v + "positional_embedding"] = _normal(g, (50, 768), scale)
    pr = _normal(g, (768, 512), scale)
    if cos_bias:
        pr = pr + 0.004
    sd[v + "proj"] = pr
    sd[v + "conv1.weight"] = _normal(g, (768, 3, 32, 32), (3 * 32 * 32) ** -0.5)
    sd[v + "ln_pre.weight"] = _uniform(g, (768,), 0.8, 1.2)
    sd[v + "ln_pre.bias"] = _normal(g, (768,), 0.05)
    _transformer(sd, v + "transformer", 768, 12, g)
    sd[v + "ln_post.weight"] = _uniform(g, (768,), 0.8, 1.2)
    sd[v + "ln_post.bias"] = _normal(g, (768,), 0.05) + (0.5 if cos_bias else 0.0)
    _transformer(sd, "transformer", 512, 12, g)
    sd["token_embedding.weight"] = _normal(g, (VOCAB, 512), 0.02)
    sd["ln_final.weight"] = _uniform(g, (512,), 0.8, 1.2)
    sd["ln_final.bias"] = _normal(g, (512,), 0.05) + (0.5 if cos_bias else 0.0)
    return sd


def synthetic_batch(batch: int, size: int = 320, max_len: int = 20, negatives: int = 3, seed: int = 1234):
    """RefCOCOg-shaped synthetic batch: img ~ N(0,1) f32 [B,3,S,S];
    word_ids int32 [B,L] = SOT, n~U{3..17} tokens, EOT, zero pad; neg_word_ids [B,negs,L]."""
    g = _g(seed)
    img = torch.randn((batch, 3, size, size), generator=g, dtype=torch.float32)

    def sentences(n):
        ids = torch.zeros((n, max_len), dtype=torch.int32)
        lens = torch.randint(3, min(17, max_len - 2) + 1, (n,), generator=g)
        body = torch.randint(1, SOT, (n, max_len), generator=g, dtype=torch.int32)
        for i in range(n):
            k = int(lens[i])
            ids[i, 0] = SOT
            ids[i, 1:1 + k] = body[i, :k]
            ids[i, 1 + k] = EOT
        return ids

    word_ids = sentences(batch)
    neg = sentences(batch * negatives).reshape(batch, negatives, max_len) if negatives > 0 else None
    return img, word_ids, neg
