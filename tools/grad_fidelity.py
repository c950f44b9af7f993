"""bf16 backward vs the fp32 oracle's autograd (evaluated on the GPU): cosine and norm ratio of every parameter gradient.
Usage: python tools/grad_fidelity.py B:seed [...]"""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
os.environ.setdefault("TRIS_ALLOW_RANDOM_INIT", "1")
from oracle import tris_oracle as O, weights as W
from tris_b200 import clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.train_step import stage1_losses
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                          attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0)
aux_sd = W.make_vitb32_clip_state_dict(7, cos_bias=True)
model = TRIS(args); model.load_state_dict(sd, strict=True); model = model.cuda().train()
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
aux.load_state_dict(aux_sd, strict=True)
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
sdc = {k: v.cuda() for k, v in sd.items()}
auxc = {k: v.cuda() for k, v in aux_sd.items()}
for a in (sys.argv[1:] or ["8:4321"]):
    B, seed = (int(v) for v in a.split(":"))
    img, ids, negs = (t.cuda() for t in W.synthetic_batch(B, 320, 20, 3, seed))
    model.load_state_dict(sd0)
    model.zero_grad(set_to_none=True)
    losses = stage1_losses(model, aux, img, ids, negs)
    losses["loss"].backward()
    ref, grads, _, _ = O.train_step(sdc, auxc, img, ids, negs)
    pd = dict(model.named_parameters())
    rows, dot, ng, nr = [], 0.0, 0.0, 0.0
    for k, r in grads.items():
        g = pd[k].grad
        if g is None or r.norm() < 1e-7:
            continue
        g, r = g.double().reshape(-1), r.double().reshape(-1)
        c = (g @ r / (g.norm() * r.norm() + 1e-30)).item()
        rows.append((c, k, (g.norm() / r.norm()).item(), r.numel()))
        dot += (g @ r).item(); ng += (g @ g).item(); nr += (r @ r).item()
    rows.sort()
    grp = lambda pre: [x for x in rows if x[1].startswith(pre)]
    print(f"B={B} seed={seed}: loss {losses['loss'].item():.4f} vs {ref['loss'].item():.4f}; "
          f"whole gradient: cosine {dot / (ng * nr) ** 0.5:.5f}, norm ratio {(ng / nr) ** 0.5:.4f}, {len(rows)} tensors")
    for name, pre in (("image tower", "backbone.visual."), ("text tower", "backbone.t"), ("fusion head", ("vis_project", "lan_project", "attn_fusion"))):
        sel = [x for x in rows if x[1].startswith(pre)]
        if sel:
            cs = sorted(x[0] for x in sel)
            print(f"  {name:12s}: {len(sel):3d} tensors, cosine min {cs[0]:.4f} / 5th pct {cs[len(cs) // 20]:.4f} / median {cs[len(cs) // 2]:.4f}")
    print("  ten lowest:", "; ".join(f"{k} cos {c:.3f} ratio {q:.3f} (n={n})" for c, k, q, n in rows[:10]))
    print("  stem conv1:", [f"{c:.4f}" for c, k, q, n in rows if k == "backbone.visual.conv1.weight"], flush=True)
