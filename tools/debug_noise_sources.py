"""How much of the bf16 image tower's c4 error is due to (a) rounding the input image to bf16, (b) rounding the weights to bf16,
(c) bf16 activation storage?  (a), (b) evaluated with the fp32 oracle tower on modified inputs; (c) = the rest."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import tris_oracle as O, weights as W
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4321
sd = {k: v.cuda() for k, v in W.make_tris_state_dict(0).items()}
img, ids, _ = W.synthetic_batch(B, 320, 20, 3, seed)
img = img.cuda()
pre = "backbone.visual."
tower = lambda s, x: O.resnet_tower(s, x, prefix=pre, train=True, new_stats={})[-1]
err = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
rb = lambda t: t.to(torch.bfloat16).float()
with torch.no_grad():
    ref = tower(sd, img)
    print(f"B={B} seed={seed}  image: mean {img.mean().item():.3f} std {img.std().item():.3f}, "
          f"neighbour-difference rms / rms {((img[..., 1:] - img[..., :-1]).pow(2).mean().sqrt() / img.pow(2).mean().sqrt()).item():.3f}")
    print(f"(a) input image rounded to bf16          : c4 rel err {err(tower(sd, rb(img)), ref):.4f}")
    sdw = {k: (rb(v) if (k.startswith(pre) and v.dim() == 4) else v) for k, v in sd.items()}
    print(f"(b) conv weights rounded to bf16          : c4 rel err {err(tower(sdw, img), ref):.4f}")
    print(f"(a)+(b)                                   : c4 rel err {err(tower(sdw, rb(img)), ref):.4f}")
    for stage in ("conv1", "conv2", "conv3", "layer1", "layer2", "layer3", "layer4"):
        sds = {k: (rb(v) if (k.startswith(pre + stage) and v.dim() == 4) else v) for k, v in sd.items()}
        print(f"    weights of {stage:7s} only rounded        : c4 rel err {err(tower(sds, img), ref):.4f}")
