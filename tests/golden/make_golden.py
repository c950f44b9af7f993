"""Generate tests/golden/stage1_golden.npz from the UNMODIFIED reference on CPU.

Run in the build container only:   python tests/golden/make_golden.py
The reference modules (model.model_stage1.TRIS, CLIP.clip.model.CLIP, train_stage1.clip_forward /
MaxLoss) are imported from /root/reference with the stub recipe of SURVEY Appendix D; weights and
inputs come from oracle.weights so the fixture can be re-derived anywhere.

The training-step glue of train_stage1.py:320-364 hard-codes .cuda(); it is driven here line by line
on CPU tensors using the reference's own clip_forward / MaxLoss functions.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from oracle import weights as W  # noqa: E402

B, SIZE, L, NEG = 3, 320, 20, 3
SUB = 8  # spatial subsampling stride for stored maps


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load_reference()
    args = ref_loader.reference_args(ns, SIZE, L, NEG, B)
    model = ns.TRIS(args)
    sd = W.make_tris_state_dict(seed=0)
    missing = model.load_state_dict(sd, strict=True)
    print("load_state_dict(strict):", missing)
    aux, _ = ns.fake_load("ViT-B/32", txt_length=L)
    aux_sd = W.make_vitb32_clip_state_dict(seed=7, cos_bias=True)
    aux.load_state_dict(aux_sd, strict=True)
    aux.eval()

    img, word_ids, neg = W.synthetic_batch(B, SIZE, L, NEG, seed=1234)
    out = {}

    # ---------------- eval forward, batch 1 (config 1 / 4 shape) ----------------
    model.eval()
    t0 = time.time()
    with torch.no_grad():
        ev = model(img[:1], word_ids[:1])
        c1, c2, c3, c4, _ = model.backbone.encode_image(img[:1])
        _, hid = model.backbone.encode_text(word_ids[:1])
    print(f"eval fwd {time.time() - t0:.1f}s", ev.shape)
    out["eval_relu_sub"] = ev[:, :, ::SUB, ::SUB].numpy()
    out["eval_relu_sum"] = np.array([ev.double().sum().item(), (ev.double() ** 2).sum().item()])
    out["eval_c4_sub"] = c4[0, ::64].numpy()
    out["eval_c1_stats"] = np.array([c1.mean().item(), c1.std().item(), c2.mean().item(), c2.std().item(),
                                     c3.mean().item(), c3.std().item(), c4.mean().item(), c4.std().item()])
    out["eval_hidden"] = hid.numpy()

    # ---------------- train forward + step ----------------
    model.train()
    t0 = time.time()
    cls, cls_fg, relu_map, sig_out, ls = model(img, word_ids)
    print(f"train fwd {time.time() - t0:.1f}s")
    # train_stage1.py:327-339
    cam_224 = F.interpolate(sig_out, (224, 224), mode="bilinear", align_corners=True)
    img_224 = F.interpolate(img, (224, 224), mode="bilinear", align_corners=True)
    fg = torch.stack([cam_224[i] * img_224[i] for i in range(B)], dim=0)
    # :340
    sim = ns.clip_forward(aux, fg, word_ids)
    fg_loss = ns.MaxLoss(sim) if ns.MaxLoss is not None else -(torch.log(sim.clamp(0.0001, 0.9999))).mean()
    # :342-353
    image_features = aux.encode_image(fg)
    cbs = torch.tensor(0.0, requires_grad=True)
    for i_ in range(B):
        _, tf = aux.encode_text(neg[i_])
        f = image_features[i_].reshape(1, -1)
        f = f / f.norm(dim=-1, keepdim=True)
        tf = tf / tf.norm(dim=-1, keepdim=True)
        cbs = cbs + (-(torch.log(1 - torch.matmul(f, tf.transpose(0, 1)))).mean())
    cbs = cbs / B
    cls_loss = F.multilabel_soft_margin_loss(cls, torch.eye(B))          # :354
    loss = fg_loss * args.w1 + cls_loss * args.w4 + cbs * args.w5           # :364
    t0 = time.time()
    model.zero_grad()
    loss.backward()
    print(f"bwd {time.time() - t0:.1f}s  loss={loss.item():.6f} l1={fg_loss.item():.6f} "
          f"l4={cls_loss.item():.6f} l5={cbs.item():.6f}")

    out["cls_out"] = cls.detach().numpy()
    out["cls_fg"] = cls_fg.detach().numpy()
    out["relu_sub"] = relu_map.detach()[:, :, ::SUB, ::SUB].numpy()
    out["sig_sub"] = sig_out.detach()[:, :, ::SUB, ::SUB].numpy()
    out["relu_sum"] = np.array([relu_map.double().sum().item(), (relu_map.double() ** 2).sum().item()])
    out["sig_sum"] = np.array([sig_out.double().sum().item(), (sig_out.double() ** 2).sum().item()])
    out["logit_scale_exp"] = np.array(ls.item())
    out["fg_sub"] = fg.detach()[:, :, ::SUB, ::SUB].numpy()
    out["cos_pos"] = sim.detach().reshape(-1).numpy()
    out["losses"] = np.array([loss.item(), fg_loss.item(), cls_loss.item(), cbs.item()])

    names, norms = [], []
    for k, p in model.named_parameters():
        names.append(k)
        norms.append(float(p.grad.double().norm()) if p.grad is not None else -1.0)
    out["grad_names"] = np.array(names)
    out["grad_norms"] = np.array(norms)
    pd = dict(model.named_parameters())
    for k in ["backbone.visual.conv1.weight", "backbone.visual.layer4.2.conv2.weight", "vis_project.weight",
              "attn_fusion.v_proj2.0.weight", "attn_fusion.t_output.0.weight", "lan_project.weight",
              "backbone.transformer.resblocks.0.attn.in_proj_weight", "backbone.text_projection",
              "backbone.positional_embedding", "logit_scale"]:
        g = pd[k].grad
        out["grad::" + k] = g.reshape(-1)[:: max(1, g.numel() // 256)][:256].numpy().copy()
    bsd = model.state_dict()
    for k in ["backbone.visual.bn1.running_mean", "backbone.visual.bn1.running_var",
              "backbone.visual.layer4.2.bn3.running_mean", "backbone.visual.layer4.2.bn3.running_var",
              "backbone.visual.layer2.0.downsample.1.running_var"]:
        out["stat::" + k] = bsd[k].numpy().copy()
    out["meta"] = np.array([B, SIZE, L, NEG, SUB, 0, 7, 1234])  # ..., tris seed, aux seed, data seed

    path = os.path.join(HERE, "stage1_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path) / 1e3, "KB")


if __name__ == "__main__":
    main()
