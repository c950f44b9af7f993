"""Execution engines: bind a module tree (spec.build_tree) to a ParamStore + kernel towers, and expose them to
torch.autograd as tower-level Functions with hand-written backward passes.

Gradient convention: tower backward passes write parameter gradients straight into ``store.grad`` (views published
as ``param.grad``); autograd only carries activation gradients between towers.
"""
from __future__ import annotations

import os

import torch

from . import _lib as L
from . import ops
from .resnet import ResNetTower
from .store import ParamStore
from .transformer import TextTower, VitTower

bf16, f32 = torch.bfloat16, torch.float32


OVERLAP = os.environ.get("TRIS_OVERLAP", "1") != "0"   # side-stream overlap of the small text towers with the image tower


class _EngineBase:
    _side = None

    def side_stream(self):
        if self._side is None:
            self._side = torch.cuda.Stream()
        return self._side

    def __init__(self, module, device, group_of, adjacent=()):
        L.require_device()
        self.module = module
        self.store = ParamStore(module, device, group_of, adjacent)
        self.anchor = torch.zeros(1, device=device, requires_grad=True)
        self.fwd_id = 0
        self.last_bwd_id = -1
        self._seen_versions = None
        self.derived_stale = True
        self.extra_grad_keys = []

    # ---- shadow maintenance
    def _versions(self):
        return sum(p._version for p in self.store.params.values())

    def ensure_fresh(self, force=False):
        """bf16 shadow + derived operands (packed 3x3 / stem weights, padded BN vectors) follow the fp32 masters.
        The shadow is re-cast only when a master changed behind it (torch-side update: param._version moved, or a load);
        the fused AdamW rewrites the shadow itself and only flags ``derived_stale``.  ``force`` (every training forward)
        re-derives the packed operands unconditionally so that a captured CUDA graph re-derives them on every replay."""
        v = self._versions()
        stale = v != self._seen_versions or not self.store.shadow_valid
        if stale or force or self.derived_stale:
            with torch.no_grad():
                if stale:
                    self.store.refresh_shadow()
                self.refresh_derived()
            self._seen_versions = self._versions()
            self.derived_stale = False

    def refresh_derived(self):
        pass

    # ---- gradient buffer protocol (see ParamStore)
    def begin_backward(self, fwd_id):
        if self.last_bwd_id == fwd_id:
            return
        self.last_bwd_id = fwd_id
        st = self.store
        if st.params[st.trainable[0]].grad is None:
            st.grad.zero_()
        st.publish_grads()
        for k in self.extra_grad_keys:
            if st.params[k].grad is None:
                st.g(k).zero_()
                st.params[k].grad = st.g(k)


# ============================================================================================ aux CLIP (ViT-B/32 or RN)
class _VitFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, patches, eng, n):
        need = patches.requires_grad
        feat, rec = eng.vit.forward(patches, n, save=need)
        ctx.eng, ctx.rec = eng, rec
        return feat

    @staticmethod
    def backward(ctx, dfeat):
        return ctx.eng.vit.backward(ctx.rec, dfeat.contiguous()), None, None


class _PatchifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, ps):
        ctx.shape, ctx.ps = img.shape, ps
        return ops.mask_resize_fwd(None, img.float().contiguous(), out_size=img.shape[2], ps=ps)[0]

    @staticmethod
    def backward(ctx, dp):
        n, c, s, _ = ctx.shape
        g, ps = s // ctx.ps, ctx.ps
        return dp.float().view(n, g, g, c, ps, ps).permute(0, 3, 1, 4, 2, 5).reshape(n, c, s, s), None   # layout only


def patchify(img, ps=32):
    """[N,3,S,S] fp32 -> bf16 [N*(S/ps)^2, 3*ps*ps] with k = c*ps*ps + py*ps + px (conv1.weight.reshape(W,-1) order).
    Differentiable w.r.t. img (the reference's training loop back-propagates through encode_image, train_stage1.py:340)."""
    return _PatchifyFn.apply(img, ps)


class _MaskResizeFn(torch.autograd.Function):
    """fg = bilinear_ac(sig -> 224) * bilinear_ac(img -> 224), emitted as ViT-B/32 patches (train_stage1.py:327-339)."""

    @staticmethod
    def forward(ctx, sig, img, size, ps):
        img = img.float().contiguous()
        patches, _ = ops.mask_resize_fwd(sig.contiguous(), img, size, ps)
        ctx.img, ctx.size, ctx.ps = img, size, ps
        return patches

    @staticmethod
    def backward(ctx, dpatches):
        return ops.mask_resize_bwd(dpatches.contiguous(), ctx.img, ctx.size, ctx.ps), None, None, None


def masked_patches(sig, img, size=224, ps=32):
    return _MaskResizeFn.apply(sig, img, size, ps)


class _LossFn(torch.autograd.Function):
    """(f [B,512] bf16, g [B*(1+K),512] bf16, cls [B,B] f32) -> (loss, l1, l4, l5); train_stage1.py:340-364."""

    @staticmethod
    def forward(ctx, f, g, cls, k, w):
        f, g, cls = f.contiguous(), g.contiguous(), cls.contiguous()
        out = ops.stage1_loss_fwd(f, g, cls, k, w)
        ctx.save_for_backward(f, g, cls)
        ctx.k, ctx.w = k, w
        ctx.set_materialize_grads(False)
        return out[0], out[1], out[2], out[3]

    @staticmethod
    def backward(ctx, d0, d1, d4, d5):
        f, g, cls = ctx.saved_tensors
        dout = torch.zeros(4, device=f.device, dtype=f32)
        for i, d in enumerate((d0, d1, d4, d5)):
            if d is not None:
                dout[i].copy_(d)
        df, dcls = ops.stage1_loss_bwd(f, g, cls, dout, ctx.k, ctx.w)
        return df, None, dcls, None, None


def stage1_loss(f, g, cls, k, w):
    return _LossFn.apply(f, g, cls, k, w)


class ClipEngine(_EngineBase):
    """Standalone CLIPModel (the frozen auxiliary scorer): no gradient buffers are published."""

    def __init__(self, module, device):
        super().__init__(module, device, group_of=lambda k: 2)
        self.text = TextTower(self.store, "")
        if module.is_vit:
            self.vit = VitTower(self.store)
        else:
            from .spec import RN_LAYERS
            self.resnet = ResNetTower(self.store, module, "visual.", RN_LAYERS[module.kind])

    def refresh_derived(self):
        if not self.module.is_vit:
            self.resnet.refresh()

    def encode_patches(self, patches, n):
        """bf16 patches [n*49, 3072] -> bf16 features [n, 512] (autograd edge to the patches: dgrad-only backward)."""
        self.ensure_fresh()
        return _VitFn.apply(patches, self, n)

    def encode_image(self, image):
        self.ensure_fresh()
        if self.module.is_vit:
            return self.encode_patches(patchify(image), image.shape[0]).float()
        with torch.no_grad():
            c4, _ = self.resnet.forward(image.float(), train=False)
        return c4.permute(0, 3, 1, 2).float()

    def encode_text(self, text):
        self.ensure_fresh()
        with torch.no_grad():
            hidden, _, seq = self.text.forward(text, save=False, full_sequence=True)
        return seq.float(), hidden.float()

    def encode_text_hidden(self, text):
        self.ensure_fresh()
        with torch.no_grad():
            return self.text.forward(text, save=False)[0]


# ============================================================================================ TRIS Stage-1
def _tris_group(key: str) -> int:
    if key.startswith("backbone."):
        if key.startswith("backbone.visual.attnpool.") or key == "backbone.logit_scale":
            return 2
        return 0
    if key.startswith(("vis_project.", "lan_project.", "attn_fusion.")):
        return 1
    return 2


_ADJ = [[f"attn_fusion.v_proj{i}.0.weight" for i in (1, 2, 3)], [f"attn_fusion.v_proj{i}.0.bias" for i in (1, 2, 3)],
        [f"attn_fusion.v_proj{i}.1.weight" for i in (1, 2, 3)], [f"attn_fusion.v_proj{i}.1.bias" for i in (1, 2, 3)],
        [f"attn_fusion.t_proj{i}.0.weight" for i in (1, 2, 3)], [f"attn_fusion.t_proj{i}.0.bias" for i in (1, 2, 3)]]


class _ResNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, img, eng):
        c4, tape = eng.resnet.forward(img, train=True)
        ctx.eng, ctx.tape, ctx.fid = eng, tape, eng.fwd_id
        return c4

    @staticmethod
    def backward(ctx, dc4):
        ctx.eng.begin_backward(ctx.fid)
        ctx.eng.resnet.backward(ctx.tape, dc4.contiguous())
        ctx.tape = None
        return None, None, None


class _TextFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, ids, eng):
        hidden, rec, _ = eng.text.forward(ids, save=True)
        ctx.eng, ctx.rec, ctx.fid = eng, rec, eng.fwd_id
        return hidden

    @staticmethod
    def backward(ctx, dh):
        ctx.eng.begin_backward(ctx.fid)
        ctx.eng.text.backward(ctx.rec, dh.contiguous())
        ctx.rec = None
        # head + text-tower gradients (everything behind the image tower in the flat buffer) are final on this stream: a
        # data-parallel trainer starts their all-reduce here, under the RN50 backward (train_step.Stage1Trainer)
        hook = getattr(ctx.eng, "on_text_grads_ready", None)
        if hook is not None:
            hook()
        return None, None, None


class Stage1Engine(_EngineBase):
    def __init__(self, model, device):
        adj = _ADJ if hasattr(model, "attn_fusion") else ()
        super().__init__(model, device, _tris_group, adj)
        from .spec import RN_LAYERS
        self.resnet = ResNetTower(self.store, model, "backbone.visual.", RN_LAYERS[model.backbone.kind])
        self.text = TextTower(self.store, "backbone.")
        self.extra_grad_keys = ["logit_scale"]
        from .head import Stage1Head
        self.head = Stage1Head(self, model)

    def refresh_derived(self):
        self.resnet.refresh()

    def towers(self, x, word_id, train: bool):
        """-> (c4 bf16 NHWC, hidden bf16 [T, E]) with autograd edges when training under grad mode."""
        self.ensure_fresh(force=train)
        x = x.float()
        if train and torch.is_grad_enabled():
            self.fwd_id += 1
            if OVERLAP:
                # the text tower (M = B*L = 960 rows: latency-bound kernels on <= 64 SMs) runs on a side stream next to
                # the image tower; autograd replays its backward on the same side stream next to the RN50 backward
                main = torch.cuda.current_stream()
                side = self.side_stream()
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    hidden = _TextFn.apply(self.anchor, word_id, self)
                c4 = _ResNetFn.apply(self.anchor, x, self)
                main.wait_stream(side)
                hidden.record_stream(main)
            else:
                c4 = _ResNetFn.apply(self.anchor, x, self)
                hidden = _TextFn.apply(self.anchor, word_id, self)
        else:
            with torch.no_grad():
                c4, _ = self.resnet.forward(x, train=train)
                hidden = self.text.forward(word_id, save=False)[0]
        return c4, hidden
