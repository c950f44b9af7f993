"""Per-parameter gradient error of the composed head vs the oracle (debug aid)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import tris_oracle as O
from oracle import weights as W
from tris_b200.model_stage1 import TRIS
bf16 = torch.bfloat16

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
def frob(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()
def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)

ap = argparse.ArgumentParser(); ap.add_argument("--B", type=int, default=4); ap.add_argument("--nofuse", action="store_true")
ap.add_argument("--which", default="both")
a = ap.parse_args()
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                          attn_multi=0.0 if a.nofuse else 0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0)
g = torch.Generator().manual_seed(3)
for k in sd:
    if k.startswith("attn_fusion.") and (k.endswith(".1.weight") or k.endswith(".1.bias")):
        sd[k] = sd[k] + 0.2 * torch.randn(sd[k].shape, generator=g)
m = TRIS(args)
m.load_state_dict(sd, strict=False)
m = m.cuda().train(); eng = m.engine(); eng.ensure_fresh(True)
B, h = a.B, 10
c4 = (rnd(B, h, h, 2048, seed=20).abs() * 0.5).to(bf16)
hidden = rnd(B, 1024, seed=21, scale=0.3, dtype=bf16)
keys = ["vis_project.weight", "vis_project.bias", "lan_project.weight", "lan_project.bias", "logit_scale"]
if not a.nofuse:
    keys += [k for k in sd if k.startswith("attn_fusion.")]
leaf = {k: sd[k].cuda() for k in keys}
for k in keys:
    leaf[k] = leaf[k].to(bf16).float().requires_grad_(True) if leaf[k].dim() > 1 else leaf[k].requires_grad_(True)
c4r = c4.float().permute(0, 3, 1, 2).requires_grad_(True); hr = hidden.float().requires_grad_(True)
score = O.tris_score(leaf, c4r, hr, attn_multi=args.attn_multi)
o = O.tris_head(score, (h, h), (320, 320), True)
eng.fwd_id += 1; eng.store.zero_grad()
c4g, hg = c4.clone().requires_grad_(True), hidden.clone().requires_grad_(True)
cls, fg, relu, sig, es = eng.head.forward(c4g, hg, (320, 320), True)
dcls, dsig = rnd(B, B, seed=22), rnd(B, 1, 320, 320, seed=23, scale=0.01)
if a.which == "cls": dsig = dsig * 0
if a.which == "sig": dcls = dcls * 0
obj = (o["cls_out"] * dcls).sum() + (o["sig"] * dsig).sum()
gr = torch.autograd.grad(obj, [c4r, hr] + [leaf[k] for k in keys], allow_unused=True)
torch.autograd.backward([cls, sig], [dcls, dsig])
print("fwd", rel(cls, o["cls_out"]), rel(sig, o["sig"]))
print("dc4 max", rel(c4g.grad.float().permute(0, 3, 1, 2), gr[0]), "frob", frob(c4g.grad.float().permute(0, 3, 1, 2), gr[0]))
print("dhid max", rel(hg.grad, gr[1]), "frob", frob(hg.grad, gr[1]))
for k, gg in zip(keys, gr[2:]):
    got = eng.store.g(k)
    if gg is None: print(k, "None"); continue
    print(f"{k:40s} max {rel(got.reshape(gg.shape), gg):.4f} frob {frob(got.reshape(gg.shape), gg):.4f} |ref| {gg.norm().item():.4g}")
