"""Parameter / buffer inventory of the Stage-1 networks (key, shape, kind, init) and generic module-tree builder.

The product does not reproduce the reference's nn.Module classes; it only has to expose the SAME ``state_dict``
(518 entries for TRIS) so reference checkpoints load and the reference's checkpoint helpers
(utils/util.py:50-107) keep working.  The tree of ``Node`` modules below is generated from this inventory.

Key / shape sources: CLIP/clip/model.py:451-506 (CLIP), :194-252 (ModifiedResNet), :10-40 (Bottleneck), :58-68
(AttentionPool2d), :366-378 (ResidualAttentionBlock), :400-417 (VisionTransformer); model/model_stage1.py:36-42;
model/attn.py:69-109.  Initialisers follow model.py:508-535 and the torch defaults of the layer types used there.
"""
from __future__ import annotations

import math
from typing import Callable, Iterator, List, Tuple

import torch
from torch import nn

Entry = Tuple[str, tuple, str, Callable]  # key, shape, "param" | "buffer" | "ibuffer", init(tensor)

RN_LAYERS = {"RN50": (3, 4, 6, 3), "RN101": (3, 4, 23, 3)}
RN_EMBED = {"RN50": 1024, "RN101": 512}
VOCAB = 49408


def _normal(std):
    return lambda t: t.normal_(0.0, std)


def _uniform_fan(fan_in):
    b = 1.0 / math.sqrt(fan_in)
    return lambda t: t.uniform_(-b, b)


_ones = lambda t: t.fill_(1.0)
_zeros = lambda t: t.zero_()


def _bn(p, c, zero_gamma=False) -> List[Entry]:
    return [(p + ".weight", (c,), "param", _zeros if zero_gamma else _ones), (p + ".bias", (c,), "param", _zeros),
            (p + ".running_mean", (c,), "buffer", _zeros), (p + ".running_var", (c,), "buffer", _ones),
            (p + ".num_batches_tracked", (), "ibuffer", _zeros)]


def _conv(key, co, ci, k) -> List[Entry]:
    return [(key, (co, ci, k, k), "param", _uniform_fan(ci * k * k))]


def _linear(p, co, ci, w_init=None) -> List[Entry]:
    return [(p + ".weight", (co, ci), "param", w_init or _uniform_fan(ci)), (p + ".bias", (co,), "param", _uniform_fan(ci))]


def _ln(p, c) -> List[Entry]:
    return [(p + ".weight", (c,), "param", _ones), (p + ".bias", (c,), "param", _zeros)]


def _transformer(p, width, layers) -> List[Entry]:
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    attn_std = width ** -0.5
    fc_std = (2 * width) ** -0.5
    out: List[Entry] = []
    for i in range(layers):
        q = f"{p}.resblocks.{i}"
        out += [(q + ".attn.in_proj_weight", (3 * width, width), "param", _normal(attn_std)),
                (q + ".attn.in_proj_bias", (3 * width,), "param", _zeros)]
        out += [(q + ".attn.out_proj.weight", (width, width), "param", _normal(proj_std)),
                (q + ".attn.out_proj.bias", (width,), "param", _zeros)]
        out += _ln(q + ".ln_1", width)
        out += [(q + ".mlp.c_fc.weight", (4 * width, width), "param", _normal(fc_std)),
                (q + ".mlp.c_fc.bias", (4 * width,), "param", _uniform_fan(width))]
        out += [(q + ".mlp.c_proj.weight", (width, 4 * width), "param", _normal(proj_std)),
                (q + ".mlp.c_proj.bias", (width,), "param", _uniform_fan(4 * width))]
        out += _ln(q + ".ln_2", width)
    return out


def _text_head(embed_dim, width=512, context=77) -> List[Entry]:
    return [("positional_embedding", (context, width), "param", _normal(0.01)),
            ("text_projection", (width, embed_dim), "param", _normal(width ** -0.5)),
            ("logit_scale", (), "param", lambda t: t.fill_(math.log(1 / 0.07)))]


def _text_tail(width=512) -> List[Entry]:
    return _transformer("transformer", width, 12) + [("token_embedding.weight", (VOCAB, width), "param", _normal(0.02))] + \
        _ln("ln_final", width)


def clip_resnet_spec(name: str = "RN50") -> List[Entry]:
    layers, embed = RN_LAYERS[name], RN_EMBED[name]
    out = _text_head(embed)
    v = "visual."
    out += _conv(v + "conv1.weight", 32, 3, 3) + _bn(v + "bn1", 32)
    out += _conv(v + "conv2.weight", 32, 32, 3) + _bn(v + "bn2", 32)
    out += _conv(v + "conv3.weight", 64, 32, 3) + _bn(v + "bn3", 64)
    inpl = 64
    for li, (planes, blocks) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for b in range(blocks):
            p = f"{v}layer{li}.{b}."
            stride = 2 if (b == 0 and li > 1) else 1
            out += _conv(p + "conv1.weight", planes, inpl, 1) + _bn(p + "bn1", planes)
            out += _conv(p + "conv2.weight", planes, planes, 3) + _bn(p + "bn2", planes)
            out += _conv(p + "conv3.weight", planes * 4, planes, 1) + _bn(p + "bn3", planes * 4, zero_gamma=True)
            if stride > 1 or inpl != planes * 4:
                out += _conv(p + "downsample.0.weight", planes * 4, inpl, 1) + _bn(p + "downsample.1", planes * 4)
            inpl = planes * 4
    a = v + "attnpool."
    std = 2048 ** -0.5
    out += [(a + "positional_embedding", (50, 2048), "param", _normal(std))]
    for nm in ("k_proj", "q_proj", "v_proj"):
        out += _linear(a + nm, 2048, 2048, _normal(std))
    out += _linear(a + "c_proj", embed, 2048, _normal(std))
    return out + _text_tail()


def clip_vit_spec(width=768, patch=32, layers=12, embed=512, res=224) -> List[Entry]:
    out = _text_head(embed)
    v = "visual."
    scale = width ** -0.5
    out += [(v + "class_embedding", (width,), "param", _normal(scale)),
            (v + "positional_embedding", ((res // patch) ** 2 + 1, width), "param", _normal(scale)),
            (v + "proj", (width, embed), "param", _normal(scale))]
    out += _conv(v + "conv1.weight", width, 3, patch)
    out += _ln(v + "ln_pre", width) + _transformer(v + "transformer", width, layers) + _ln(v + "ln_post", width)
    return out + _text_tail()


def tris_head_spec(hidden=1024, textdim=1024, vis_ch=2048) -> List[Entry]:
    out: List[Entry] = []
    out += [("vis_project.weight", (hidden, vis_ch, 1, 1), "param", _uniform_fan(vis_ch)),
            ("vis_project.bias", (hidden,), "param", _uniform_fan(vis_ch))]
    out += _linear("lan_project", hidden, textdim)
    for nm in ("v_proj1", "v_proj2", "v_proj3"):
        out += [(f"attn_fusion.{nm}.0.weight", (hidden, hidden, 1, 1), "param", _uniform_fan(hidden)),
                (f"attn_fusion.{nm}.0.bias", (hidden,), "param", _uniform_fan(hidden)),
                (f"attn_fusion.{nm}.1.weight", (hidden,), "param", _ones), (f"attn_fusion.{nm}.1.bias", (hidden,), "param", _zeros)]
    for nm in ("t_proj1", "t_proj2", "t_proj3"):
        out += _linear(f"attn_fusion.{nm}.0", hidden, hidden)
    out += [("attn_fusion.v_output.0.weight", (hidden, hidden, 1, 1), "param", _uniform_fan(hidden)),
            ("attn_fusion.v_output.0.bias", (hidden,), "param", _uniform_fan(hidden)),
            ("attn_fusion.v_output.1.weight", (hidden,), "param", _ones), ("attn_fusion.v_output.1.bias", (hidden,), "param", _zeros)]
    out += _linear("attn_fusion.t_output.0", hidden, hidden)
    return out


class Node(nn.Module):
    """Anonymous container: parameters/buffers/children are attached by ``build_tree`` from dotted keys."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("tris_b200 container nodes are not callable; use the owning model's forward")


def build_tree(root: nn.Module, entries: List[Entry], prefix: str = "") -> None:
    """Register every (key, shape, kind, init) under ``root`` creating nested Nodes; numeric path components become
    children of a Node exactly like nn.Sequential / ModuleList children, so the state_dict keys match."""
    for key, shape, kind, init in entries:
        parts = (prefix + key).split(".")
        mod = root
        for comp in parts[:-1]:
            nxt = mod._modules.get(comp)
            if nxt is None:
                nxt = Node()
                mod.add_module(comp, nxt)
            mod = nxt
        if kind == "param":
            t = torch.empty(shape, dtype=torch.float32)
            init(t)
            mod.register_parameter(parts[-1], nn.Parameter(t))
        else:
            t = torch.zeros(shape, dtype=torch.int64 if kind == "ibuffer" else torch.float32)
            init(t)
            mod.register_buffer(parts[-1], t)


def walk(entries: List[Entry], prefix: str = "") -> Iterator[Tuple[str, tuple, str]]:
    for key, shape, kind, _ in entries:
        yield prefix + key, shape, kind
