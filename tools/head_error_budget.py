"""Which bf16 rounding point of the head dominates the gradient error?  Emulates the product's storage roundings one at a
time inside the fp32 oracle math (torch on CUDA) and prints the resulting parameter-gradient error."""
import math, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
warnings.simplefilter("ignore")
from oracle import weights as W
bf16 = torch.bfloat16
dev = "cuda"

def frob(a, b): return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()
def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.randn(*shape, generator=g, device=dev) * scale

class RF(torch.autograd.Function):      # round forward value
    @staticmethod
    def forward(ctx, x): return x.to(bf16).float()
    @staticmethod
    def backward(ctx, g): return g
class RB(torch.autograd.Function):      # round backward gradient
    @staticmethod
    def forward(ctx, x): return x.view_as(x)
    @staticmethod
    def backward(ctx, g): return g.to(bf16).float()

ON = set()
def f(name, x): return RF.apply(x) if ("f:" + name) in ON or "f:*" in ON else x
def b(name, x): return RB.apply(x) if ("b:" + name) in ON or "b:*" in ON else x
def fb(name, x): return b(name, f(name, x))

def inorm(x, w, bb):
    mu = x.mean(1, keepdim=True); var = x.var(1, unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + 1e-5) * w + bb

def head(P, c4, hidden, B, h):
    C = 1024
    vis = fb("vis", c4.flatten(2).transpose(1, 2) @ P["vis_project.weight"].reshape(C, -1).t() + P["vis_project.bias"])
    lan = fb("lan", hidden @ P["lan_project.weight"].t() + P["lan_project.bias"])
    nv = fb("nv", vis / vis.norm(dim=-1, keepdim=True)); nl = fb("nl", lan / lan.norm(dim=-1, keepdim=True))
    a = "attn_fusion."
    vp = []
    for i in (1, 2, 3):
        y = fb(f"Yv{i}", nv @ P[f"{a}v_proj{i}.0.weight"].reshape(C, C).t())
        vp.append(fb(f"A3{i}", F.relu(inorm(y, P[f"{a}v_proj{i}.1.weight"], P[f"{a}v_proj{i}.1.bias"]))))
    tp = [fb(f"At3{i}", F.relu(nl @ P[f"{a}t_proj{i}.0.weight"].t() + P[f"{a}t_proj{i}.0.bias"])) for i in (1, 2, 3)]
    qv, kv, vv = vp; qt, kt, vt = tp
    PA = torch.softmax(qv @ kt.t() / 32, dim=2)
    PT = torch.softmax((kv @ qt.t()) / 32, dim=1)          # [B,P,T] softmax over P
    PAc = fb("PAc", PA - PA.mean(1, keepdim=True)); PT = fb("PT", PT)
    nvp = fb("nvp", PAc @ vt)
    nlp = fb("nlp", PT.transpose(1, 2) @ vv)
    Ov = fb("Ov", nvp @ P[a + "v_output.0.weight"].reshape(C, C).t())
    vpr = fb("vp", nv + 0.1 * inorm(Ov, P[a + "v_output.1.weight"], P[a + "v_output.1.bias"]))
    Ol = fb("Ol", nlp @ P[a + "t_output.0.weight"].t() + P[a + "t_output.0.bias"])
    lpr = fb("lp", nl.unsqueeze(0) + 0.1 * Ol)
    R = vpr @ lpr.transpose(1, 2)
    return P["logit_scale"].exp() * R

def run(sd, B=4, h=10):
    from oracle import tris_oracle as O
    c4 = (rnd(B, 2048, h, h, seed=20).abs() * 0.5).to(bf16).float()
    hidden = rnd(B, 1024, seed=21, scale=0.3).to(bf16).float()
    keys = [k for k in sd if k.startswith(("attn_fusion.", "vis_project", "lan_project")) or k == "logit_scale"]
    P = {k: (sd[k].to(dev).to(bf16).float() if sd[k].dim() > 1 else sd[k].to(dev)).requires_grad_(True) for k in keys}
    dcls, dsig = rnd(B, B, seed=22), rnd(B, 1, 320, 320, seed=23, scale=0.01)
    score = head(P, c4, hidden, B, h)
    o = O.tris_head(score, (h, h), (320, 320), True)
    obj = (o["cls_out"] * dcls).sum() + (o["sig"] * dsig).sum()
    return dict(zip(keys, torch.autograd.grad(obj, [P[k] for k in keys], allow_unused=True)))

sd = W.make_tris_state_dict(0)
g = torch.Generator().manual_seed(3)
for k in sd:
    if k.startswith("attn_fusion.") and (k.endswith(".1.weight") or k.endswith(".1.bias")):
        sd[k] = sd[k] + 0.2 * torch.randn(sd[k].shape, generator=g)
ref = run(sd)
WATCH = ["vis_project.weight", "attn_fusion.v_proj1.0.weight", "attn_fusion.v_proj3.0.weight", "attn_fusion.t_proj2.0.weight",
         "attn_fusion.t_proj3.0.weight", "attn_fusion.v_output.0.weight"]
names = ["vis", "nv", "nl", "Yv1", "Yv3", "A31", "A32", "A33", "At31", "At32", "At33", "PAc", "PT", "nvp", "nlp", "Ov", "vp", "Ol", "lp"]
print("point".ljust(10), " ".join(w.replace("attn_fusion.","").replace(".0.weight","").replace(".weight","")[-9:].rjust(9) for w in WATCH))
for mode in ("f", "b"):
    for n in names + ["*"]:
        ON.clear(); ON.add(f"{mode}:{n}")
        got = run(sd)
        print(f"{mode}:{n}".ljust(10), " ".join(f"{frob(got[w], ref[w]):9.4f}" for w in WATCH))
ON.clear(); ON.update({"f:*", "b:*"})
got = run(sd)
print("all".ljust(10), " ".join(f"{frob(got[w], ref[w]):9.4f}" for w in WATCH))
