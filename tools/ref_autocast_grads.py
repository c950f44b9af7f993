"""What does bf16 do to the REFERENCE's own gradients?  Unmodified reference (baseline/_ref) on the GPU: gradients of the Stage-1
loss under torch.autocast(bfloat16) vs true fp32, per-parameter cosine -- the yardstick for tools/grad_fidelity.py."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
warnings.simplefilter("ignore")
import torch
import torch.nn.functional as F
from baseline import ref_step as RS
from tris_b200.synthetic import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4321
ns, args = RS.load(batch=B)
model, aux = RS.build_models(ns, args, "cuda", aux_half=False)
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def loss_of(img, ids, neg):
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
    return l1 * args.w1 + l4 * args.w4 + l5 * args.w5


img, ids, neg = (t.cuda() for t in synthetic_batch(B, 320, 20, 3, seed=seed))
ids, neg = ids.long(), neg.long()
G = {}
for mode in ("fp32", "bf16"):
    model.load_state_dict(sd0)
    model.train()
    model.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        loss = loss_of(img, ids, neg)
    loss.backward()
    G[mode] = ({k: p.grad.detach().double().reshape(-1).clone() for k, p in model.named_parameters() if p.grad is not None}, loss.item())
rows, dot, ng, nr = [], 0.0, 0.0, 0.0
for k, r in G["fp32"][0].items():
    g = G["bf16"][0].get(k)
    if g is None or r.norm() < 1e-7:
        continue
    rows.append(((g @ r / (g.norm() * r.norm() + 1e-30)).item(), k, (g.norm() / r.norm()).item(), r.numel()))
    dot += (g @ r).item(); ng += (g @ g).item(); nr += (r @ r).item()
rows.sort()
print(f"reference autocast(bf16) vs reference fp32, B={B} seed={seed}: loss {G['bf16'][1]:.4f} vs {G['fp32'][1]:.4f}; "
      f"whole gradient: cosine {dot / (ng * nr) ** 0.5:.5f}, norm ratio {(ng / nr) ** 0.5:.4f}, {len(rows)} tensors")
for name, pre in (("image tower", "backbone.visual."), ("text tower", "backbone.t"), ("fusion head", ("vis_project", "lan_project", "attn_fusion"))):
    sel = sorted(x[0] for x in rows if x[1].startswith(pre))
    if sel:
        print(f"  {name:12s}: {len(sel):3d} tensors, cosine min {sel[0]:.4f} / 5th pct {sel[len(sel) // 20]:.4f} / median {sel[len(sel) // 2]:.4f}")
print("  ten lowest:", "; ".join(f"{k} cos {c:.3f} ratio {q:.3f} (n={n})" for c, k, q, n in rows[:10]))
print("  stem conv1:", [f"{c:.4f}" for c, k, q, n in rows if k == "backbone.visual.conv1.weight"], flush=True)
