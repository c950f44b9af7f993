"""Drive the UNMODIFIED reference (baseline/_ref or /root/reference) on synthetic Stage-1 batches.

* `train_epoch_reference_loop` calls the reference's own `train_one_epoch` (train_stage1.py:286-411) -- it hard-codes
  `.cuda()`, so it needs a GPU; any object with the `TRIS` / aux-CLIP surface can be passed as `model` / `clip_model`
  (this is how tests/test_dropin_gpu.py proves that tris_b200.TRIS is a drop-in).
* `cpu_train_step` restates the loop body (train_stage1.py:320-372) on CPU tensors using the reference's own
  `clip_forward` / `MaxLoss` -- the CPU arm of bench.py.
* `make_val_loader` feeds the reference's `validate` / `validate_same_sentence` (validate.py:131-249, 253-387).

Test / benchmark infrastructure only: nothing under tris_b200/ imports this module.
"""
from __future__ import annotations

import logging
import os
import sys
import types

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline import ref_loader  # noqa: E402


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass


def load(size=320, max_len=20, negs=3, batch=48, epochs=15):
    """-> (ns, args): reference namespace (TRIS, clip, train_stage1 / validate modules) + its argparse Namespace."""
    ns = ref_loader.load_reference_once()
    args = ref_loader.reference_args(ns, size, max_len, negs, batch)
    args.epoch = epochs
    ns.T.writer = _NullWriter()
    ns.T.logger = logging.getLogger("tris_reference")
    return ns, args


def build_models(ns, args, device="cpu", aux_half=None, seed=0):
    """Reference TRIS (fp32, as model_stage1.py:31 forces) + aux ViT-B/32 with oracle.weights' deterministic weights.
    On CUDA the aux model keeps fp16 weights exactly as the reference's clip.load leaves it (CLIP/clip/model.py:642;
    only floated on CPU, clip.py:143-144)."""
    from oracle import weights as W
    model = ns.TRIS(args)
    model.load_state_dict(W.make_tris_state_dict(seed), strict=True)
    aux, _ = ns.fake_load("ViT-B/32", txt_length=args.max_query_len)
    aux.load_state_dict(W.make_vitb32_clip_state_dict(7, cos_bias=True), strict=True)
    aux.eval()
    dev = torch.device(device)
    model = model.to(dev)
    aux = aux.to(dev)
    if aux_half is None:
        aux_half = dev.type == "cuda"
    if aux_half:
        ns.convert_weights(aux)
    return model, aux


def make_optimizer(model, args, max_iter):
    """train_stage1.py:133-144: AdamW with two lr groups + per-step poly-0.9 LambdaLR."""
    backbone, new = model.trainable_parameters()
    opt = torch.optim.AdamW([{"params": backbone, "lr": args.lr * args.lr_multi},
                             {"params": new, "lr": args.lr}], lr=args.lr, weight_decay=args.weight_decay)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda x: (1 - x / max_iter) ** 0.9)
    return opt, sched


def make_train_loader(n_batches, batch, size=320, max_len=20, negs=3, seed=1234):
    """List of (samples, targets) with the tensor contract of ReferDataset in train mode (ReferDataset.py:190-252)."""
    from tris_b200.synthetic import synthetic_batch
    out = []
    for i in range(n_batches):
        img, ids, neg = synthetic_batch(batch, size, max_len, negs, seed=seed + i)
        samples = {"img": img, "word_ids": ids.long().unsqueeze(1), "word_masks": (ids != 0).long().unsqueeze(1)}
        if neg is not None:
            samples["neg_word_ids"] = neg.long()
        targets = {"target": torch.zeros((batch, 1, size, size), dtype=torch.int64), "boxes": torch.zeros((batch, 4)),
                   "sentences": ["synthetic"] * batch}
        out.append((samples, targets))
    return out


def train_epoch_reference_loop(ns, args, model, aux, loader, opt, sched, iteration=0):
    """The reference's own hot loop, unmodified (needs CUDA: it calls .cuda() on every tensor)."""
    return ns.T.train_one_epoch(loader, model, opt, 0, 0, args, iteration=iteration, clip_model=aux, lr_scheduler=sched)


def cpu_train_step(ns, args, model, aux, opt, sched, img, word_ids, neg):
    """train_stage1.py:320-372 restated for CPU tensors (the reference loop hard-codes .cuda()); the loss glue is the
    reference's own clip_forward / MaxLoss."""
    T = ns.T
    B = img.shape[0]
    labels = torch.eye(B)
    cls, _, _, sig_out, _ = model(img, word_ids)
    cam_224 = F.interpolate(sig_out, (224, 224), mode="bilinear", align_corners=True)
    img_224 = F.interpolate(img, (224, 224), mode="bilinear", align_corners=True)
    fg = torch.stack([cam_224[i] * img_224[i] for i in range(B)], dim=0)
    fg_loss = T.MaxLoss(T.clip_forward(aux, fg, word_ids))
    cbs = torch.tensor(0.0, requires_grad=True)
    if args.negative_samples > 0:
        image_features = aux.encode_image(fg)
        for i_ in range(B):
            _, tf = aux.encode_text(neg[i_])
            f = image_features[i_].reshape(1, -1)
            f = f / f.norm(dim=-1, keepdim=True)
            tf = tf / tf.norm(dim=-1, keepdim=True)
            cbs = cbs + (-(torch.log(1 - torch.matmul(f, tf.transpose(0, 1)))).mean())
        cbs = cbs / B
    cls_loss = F.multilabel_soft_margin_loss(cls, labels)
    loss = fg_loss * args.w1 + cls_loss * args.w4 + cbs * args.w5
    opt.zero_grad()
    loss.backward()
    opt.step()
    sched.step()
    return {"loss": loss.detach(), "l1": fg_loss.detach(), "l4": cls_loss.detach(), "l5": cbs.detach()}


def make_val_loader(n_refs, size=320, max_len=20, sentences=2, orig=(480, 640), seed=0):
    """List of (samples, targets) like the eval-mode ReferDataset at batch_size 1 (all sentences of a ref stacked on
    the last dim: word_ids [1,1,L,S]).  Same refs as validate.py::synthetic_refs of this repo."""
    from tris_b200.synthetic import synthetic_batch
    out = []
    for idx in range(n_refs):
        img, _, _ = synthetic_batch(1, size, max_len, 0, seed=77000 + seed + idx)
        _, ids, _ = synthetic_batch(sentences, 32, max_len, 0, seed=99000 + seed + idx)
        g = torch.Generator().manual_seed(seed + idx)
        target = torch.zeros((1, *orig), dtype=torch.int64)
        y0 = int(torch.randint(0, orig[0] // 2, (1,), generator=g))
        x0 = int(torch.randint(0, orig[1] // 2, (1,), generator=g))
        target[:, y0:y0 + orig[0] // 3, x0:x0 + orig[1] // 3] = 1
        wid = ids.long().t().reshape(1, 1, max_len, sentences).contiguous()
        samples = {"img": img, "word_ids": wid, "word_masks": (wid != 0).long()}
        targets = {"target": target, "img_path": torch.tensor([idx]), "sentences": ["synthetic"],
                   "boxes": torch.tensor([[float(x0), float(y0), float(x0 + orig[1] // 3), float(y0 + orig[0] // 3)]])}
        out.append((samples, targets))
    return out
