"""GPU-box debug helper: stage-by-stage comparison of the product towers with the CPU oracle (train mode)."""
import argparse, sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
warnings.simplefilter("ignore")
from oracle import tris_oracle as O, weights as W
from tris_b200.model_stage1 import TRIS

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item(), ((a - b).norm() / (b.norm() + 1e-12)).item()

args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
train = "--eval" not in sys.argv
B = 3
sd = W.make_tris_state_dict(0)
img, ids, negs = W.synthetic_batch(B, 320, 20, 3, 1234)
m = TRIS(args); m.load_state_dict(sd); m = m.cuda(); m.train(train)
eng = m.engine(); eng.ensure_fresh(True)
with torch.no_grad():
    c4, tape = eng.resnet.forward(img.cuda(), train=train)
    hidden = eng.text.forward(ids.cuda(), save=False)[0]
# oracle with intermediates
torch.set_num_threads(16)
x = img
p = "backbone.visual."
ns = {}
outs = {}
with torch.no_grad():
    for i in (1, 2, 3):
        x = F.conv2d(x, sd[f"{p}conv{i}.weight"], stride=2 if i == 1 else 1, padding=1)
        outs[f"stem_y{i}"] = x
        x = F.relu(O.batch_norm(x, sd, f"{p}bn{i}", train, ns))
        outs[f"stem_a{i}"] = x
    x = F.avg_pool2d(x, 2)
    outs["stem"] = x
    for li, blocks in enumerate((3, 4, 6, 3), start=1):
        for b in range(blocks):
            x = O.bottleneck(x, sd, f"{p}layer{li}.{b}", 2 if (b == 0 and li > 1) else 1, train, ns)
            outs[f"layer{li}.{b}"] = x
    _, hid = O.encode_text(sd, ids, prefix="backbone.")
nhwc = lambda t: t.permute(0, 2, 3, 1)
if train:
    col, y1, a1, y2, a2, y3 = tape["stem"]
    print("stem y1", rel(y1[..., :32], nhwc(outs["stem_y1"])), "pad max", y1[..., 32:].abs().max().item())
    print("stem a1", rel(a1[..., :32], nhwc(outs["stem_a1"])), "pad max", a1[..., 32:].abs().max().item())
    print("stem y2", rel(y2[..., :32], nhwc(outs["stem_y2"])))
    print("stem a2", rel(a2[..., :32], nhwc(outs["stem_a2"])))
    print("stem y3", rel(y3, nhwc(outs["stem_y3"])))
    for blk in eng.resnet.blocks:
        rec = tape[blk.p]
        name = blk.p[len(p):-1]
        print(name, "in", rel(rec[0], nhwc(outs["stem"] if name == "layer1.0" else outs[prev])), "out", rel(rec[8], nhwc(outs[name])))
        prev = name
print("c4", rel(c4, nhwc(outs["layer4.2"])))
print("hidden", rel(hidden, hid))
