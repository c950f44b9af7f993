"""Flat parameter / gradient / bf16-shadow storage for a module tree.

All float parameters of a model become views into ONE fp32 buffer (``flat``), their gradients views into one fp32
buffer (``grad``) and their GEMM operands views into one bf16 buffer (``shadow``).  That gives
  * a single NCCL all-reduce per step over ``grad[:n_train]`` (north-star: one gradient all-reduce),
  * a single fused AdamW launch over ``flat[:n_train]`` (train_stage1.py:133-144, 368-372),
  * one fp32->bf16 conversion launch for every tensor-core operand.
Layout order: [group 0 = backbone (lr * lr_multi)] [group 1 = new modules (lr)] [no-grad rest]; inside a group,
``adjacent`` name lists are laid out back to back so that concatenated weight matrices are free views.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch
from torch import nn

from . import ops


class ParamStore:
    def __init__(self, module: nn.Module, device, group_of: Callable[[str], int], adjacent: Sequence[Sequence[str]] = ()):
        named = [(k, p) for k, p in module.named_parameters() if p.is_floating_point()]
        order_hint: Dict[str, float] = {}
        for i, (k, _) in enumerate(named):
            order_hint[k] = float(i)
        for names in adjacent:  # pull each adjacency list together at the position of its first member
            base = order_hint[names[0]]
            for j, k in enumerate(names):
                order_hint[k] = base + j * 1e-4
        groups: List[List[tuple]] = [[], [], []]
        for k, p in named:
            g = group_of(k)
            groups[g if g in (0, 1) else 2].append((order_hint[k], k, p))
        self.offsets: Dict[str, int] = {}
        self.shapes: Dict[str, torch.Size] = {}
        self.group_bounds: List[int] = [0]
        off = 0
        ordered = []
        for g in groups:
            for _, k, p in sorted(g, key=lambda t: t[0]):
                self.offsets[k] = off
                self.shapes[k] = p.shape
                n = p.numel()
                off += (n + 7) // 8 * 8  # keep every tensor 32-byte (fp32) / 16-byte (bf16) aligned
                ordered.append((k, p))
            self.group_bounds.append(off)
        self.total = off
        self.n_train = self.group_bounds[2]
        self.device = torch.device(device)
        self.flat = torch.zeros(self.total, device=self.device, dtype=torch.float32)
        self.grad = torch.zeros(self.total if self.n_train > 0 else 8, device=self.device, dtype=torch.float32)
        self.shadow = torch.zeros(self.total, device=self.device, dtype=torch.bfloat16)
        self.params: Dict[str, nn.Parameter] = {}
        with torch.no_grad():
            for k, p in ordered:
                v = self.p(k)
                v.copy_(p.detach().to(self.device))
                p.data = v
                p.grad = None
                self.params[k] = p
        self.trainable = [k for k, _ in ordered if self.offsets[k] < self.n_train]
        self._probe = ordered[0][0]
        self.shadow_valid = False
        self._versions = None

    # ---- views
    def _view(self, buf, k):
        o = self.offsets[k]
        shp = self.shapes[k]
        return buf[o:o + shp.numel()].view(shp)

    def p(self, k):
        return self._view(self.flat, k)

    def g(self, k):
        if self.n_train == 0:
            return None
        return self._view(self.grad, k)

    def s(self, k):
        return self._view(self.shadow, k)

    def cat(self, buf_name: str, names: Sequence[str], shape):
        """View over several back-to-back tensors (declared ``adjacent`` at construction)."""
        buf = getattr(self, buf_name)
        o = self.offsets[names[0]]
        n = 0
        for k in names:
            assert self.offsets[k] == o + n, f"{k} is not adjacent in the flat layout"
            n += (self.shapes[k].numel() + 7) // 8 * 8
            assert self.shapes[k].numel() % 8 == 0
        return buf[o:o + n].view(shape)

    # ---- maintenance
    def still_bound(self) -> bool:
        k = self._probe
        return self.params[k].data_ptr() == self.p(k).data_ptr()

    def refresh_shadow(self):
        ops.f32_to_bf16(self.flat, self.shadow)
        self.shadow_valid = True

    def zero_grad(self):
        self.grad.zero_()

    def begin_backward_target(self):
        """Called at the start of a training forward: if the caller dropped .grad (zero_grad(set_to_none=True)) the
        flat gradient buffer still holds the previous step and must be cleared."""
        if self.params[self.trainable[0]].grad is None:
            self.zero_grad()

    def publish_grads(self):
        for k in self.trainable:
            p = self.params[k]
            if p.grad is None:
                p.grad = self.g(k)
