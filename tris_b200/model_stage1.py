"""Drop-in for the reference's ``model.model_stage1.TRIS`` (model/model_stage1.py:14-123).

Same constructor (``TRIS(args)``), same ``state_dict`` keys/shapes (518 entries), same ``trainable_parameters()``
grouping, same ``forward(x, word_id)`` return convention: train -> (cls_out[B,B], cls_fg[B], relu(seg)[B,1,H,W],
sigmoid(seg)[B,1,H,W], exp(logit_scale)), eval -> relu(seg).  All device arithmetic goes through libtris_sm100.so;
constructing the model works on CPU (for checkpoints), running it requires an sm_100a GPU.
"""
from __future__ import annotations

import math

import torch
from torch import nn

from . import _lib as L
from . import clip_model as clip
from . import spec as S


class TRIS(nn.Module):
    def __init__(self, args=None):
        super().__init__()
        self.args = args
        self.bert_model = args.bert_tokenizer
        if args.backbone == "clip-RN50":
            last_vis_channel, self.textdim = 2048, 1024
        elif args.backbone == "clip-RN101":
            last_vis_channel, self.textdim = 2048, 512
        else:
            raise ValueError(f"Stage-1 TRIS supports clip-RN50 / clip-RN101 only (got {args.backbone}); SURVEY F4")
        kind = args.backbone.split("-")[-1]
        # random-init backbone only on explicit request (args.synthetic_weights / TRIS_ALLOW_RANDOM_INIT=1): the reference's
        # clip.load downloads the OpenAI weights or raises (CLIP/clip/clip.py:122-127)
        clip_model, _ = clip.load(kind, device="cpu", jit=False, txt_length=args.max_query_len,
                                  allow_random_init=getattr(args, "synthetic_weights", None))
        self.backbone = clip_model.float()
        entries = S.tris_head_spec(args.hidden_dim, self.textdim, last_vis_channel)
        if not args.attn_multi > 0:
            entries = [e for e in entries if not e[0].startswith("attn_fusion.")]
        S.build_tree(self, entries)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self._eng = None
        self.precision = "bf16"
        self.train()

    def set_precision(self, precision: str):
        """"bf16" (default): tcgen05 bf16 path, forward + backward.  "fp32": forward-only parity mode in full fp32
        (tris_b200/precise.py) for checking response maps against the reference at 1e-3."""
        if precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        self.precision = precision
        return self

    def trainable_parameters(self):
        new = [self.vis_project, self.lan_project]
        if hasattr(self, "attn_fusion"):
            new.append(self.attn_fusion)
        else:
            print("no attn fusion")
        return list(self.backbone.parameters()), list(nn.ModuleList(new).parameters())

    # ------------------------------------------------------------------
    def engine(self):
        dev = self.logit_scale.device
        if dev.type != "cuda":
            raise L.TrisLibError("tris_b200.TRIS.forward needs a CUDA sm_100a device (no CPU fallback); move the model "
                                 "with .cuda() first")
        if self._eng is None or self._eng.store.device != dev or not self._eng.store.still_bound():
            from .engine import Stage1Engine
            self._eng = Stage1Engine(self, dev)
        return self._eng

    def forward(self, x, word_id):
        if word_id.dim() != 2 or word_id.shape[1] != self.args.max_query_len:
            raise ValueError(f"word_id must be [B, max_query_len={self.args.max_query_len}], got {tuple(word_id.shape)}")
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] % 32 or x.shape[3] % 32:
            raise ValueError(f"x must be [B,3,H,W] with H,W multiples of 32, got {tuple(x.shape)}")
        eng = self.engine()
        if self.precision == "fp32":
            if self.training and torch.is_grad_enabled():
                raise L.TrisLibError("precision='fp32' is a forward-only parity mode: call it under torch.no_grad()")
            from .precise import PreciseStage1
            return PreciseStage1(self).forward(x, word_id, self.training)
        c4, hidden = eng.towers(x, word_id, self.training)
        out = eng.head.forward(c4, hidden, tuple(x.shape[2:]), self.training)
        return out if self.training else out[0]


    # ------------------------------------------------------------------ inference helpers (validate.py / demo.py)
    @torch.no_grad()
    def image_features(self, x):
        """RN50 tower once per image (eval BN): -> c4 bf16 NHWC.  validate.py re-runs the whole network for every
        sentence of the same image (validate.py:173-179); the drivers here cache this instead."""
        eng = self.engine()
        eng.ensure_fresh()
        return eng.resnet.forward(x.float(), train=False)[0]

    @torch.no_grad()
    def respond(self, c4, word_id, img_size):
        """Response maps relu(seg) [B,1,H,W] for sentences word_id [B,L] paired with cached image features c4 [B,...]."""
        eng = self.engine()
        hidden = eng.text.forward(word_id, save=False)[0]
        return eng.head.forward(c4, hidden, img_size, False)[0]


def focal_loss(x, p=1, c=0.1):
    return torch.pow(1 - x, p) * torch.log(c + x)
