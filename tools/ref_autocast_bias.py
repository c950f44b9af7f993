"""Loss of one Stage-1 step at the benchmark batch (48) for five seeds, on the GPU box:
   (a) unmodified reference, true fp32 (TF32 off)        -- the parity target
   (b) unmodified reference under torch.autocast(bf16)   -- what bf16 does to the REFERENCE's own code
   (c) unmodified reference with TF32 convs/matmuls
   (d) tris_b200, default bf16 path
Prints |x - a| / |a| per seed.  Evidence for the bf16 loss tolerance discussion."""
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import torch
import torch.nn.functional as F

from baseline import ref_step as RS
from oracle import weights as W
from tris_b200 import clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.synthetic import synthetic_batch
from tris_b200.train_step import stage1_losses

B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
SEEDS = (1234, 4321, 7, 99, 2024)
ns, args = RS.load(batch=B)


def ref_loss(model, aux, img, ids, neg):
    T = ns.T
    cls, _, _, sig_out, _ = model(img, ids)
    cam = F.interpolate(sig_out, (224, 224), mode="bilinear", align_corners=True)
    i224 = F.interpolate(img, (224, 224), mode="bilinear", align_corners=True)
    fg = cam * i224
    l1 = T.MaxLoss(T.clip_forward(aux, fg, ids))
    f = aux.encode_image(fg)
    f = f / f.norm(dim=-1, keepdim=True)
    _, t = aux.encode_text(neg.reshape(-1, neg.shape[-1]))
    t = (t / t.norm(dim=-1, keepdim=True)).reshape(B, -1, t.shape[-1])
    l5 = (-(torch.log(1 - torch.einsum("bc,bkc->bk", f.float(), t.float())))).mean()
    l4 = F.multilabel_soft_margin_loss(cls.float(), torch.eye(B, device=cls.device))
    return float(l1 * args.w1 + l4 * args.w4 + l5 * args.w5), float(l1), float(l4), float(l5)


rows = []
model, aux = RS.build_models(ns, args, "cuda", aux_half=False)
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
args.synthetic_weights = True
ours = TRIS(args)
ours.load_state_dict(W.make_tris_state_dict(0), strict=True)
ours = ours.cuda().train()
oaux = clip_model.CLIPModel("ViT-B/32", txt_length=20)
oaux.load_state_dict(W.make_vitb32_clip_state_dict(7, cos_bias=True), strict=True)
oaux = oaux.cuda().eval()
osd0 = {k: v.clone() for k, v in ours.state_dict().items()}
for seed in SEEDS:
    img, ids, neg = (t.cuda() for t in synthetic_batch(B, 320, 20, 3, seed=seed))
    ids_l, neg_l = ids.long(), neg.long()
    out = {}
    for mode in ("fp32", "bf16-autocast", "tf32"):
        model.load_state_dict(sd0)
        model.train()
        torch.backends.cudnn.allow_tf32 = mode == "tf32"
        torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16-autocast")):
            out[mode] = ref_loss(model, aux, img, ids_l, neg_l)
    ours.load_state_dict(osd0)
    ours.train()
    with torch.no_grad():
        lo = stage1_losses(ours, oaux, img, ids, neg)
    out["tris_b200"] = tuple(float(lo[k]) for k in ("loss", "l1", "l4", "l5"))
    a = out["fp32"][0]
    print(f"seed {seed:5d} B={B}: ref fp32 {a:.5f} | " + " | ".join(
        f"{k} {v[0]:.5f} ({abs(v[0] - a) / abs(a) * 100:.3f} %)" for k, v in out.items() if k != "fp32"), flush=True)
    print("            terms (loss,l1,l4,l5):", {k: tuple(round(x, 4) for x in v) for k, v in out.items()}, flush=True)
