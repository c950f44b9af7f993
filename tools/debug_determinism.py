"""GPU-box debug helper: run the towers twice on identical inputs and report run-to-run differences."""
import argparse, sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import weights as W
from tris_b200.model_stage1 import TRIS

def d(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()

args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0)
img, ids, negs = W.synthetic_batch(3, 320, 20, 3, 1234)
m = TRIS(args); m.load_state_dict(sd); m = m.cuda().train()
eng = m.engine(); eng.ensure_fresh(True)
img, ids = img.cuda(), ids.cuda()
runs = []
with torch.no_grad():
    for r in range(3):
        m.load_state_dict(sd)
        eng.ensure_fresh(True)
        c4, tape = eng.resnet.forward(img, train=True)
        hidden = eng.text.forward(ids, save=False)[0]
        outs = eng.head._run(c4, hidden, (320, 320), True)
        torch.cuda.synchronize()
        runs.append((c4.clone(), hidden.clone(), [o.clone() for o in outs], {k: [t.clone() if t is not None else None for t in v] for k, v in tape.items()}))
for r in (1, 2):
    print("run", r, "vs 0: c4", d(runs[r][0], runs[0][0]), "hidden", d(runs[r][1], runs[0][1]),
          "cls", d(runs[r][2][0], runs[0][2][0]), "sig", d(runs[r][2][3], runs[0][2][3]))
    for k in runs[0][3]:
        diffs = [d(a, b) if a is not None else 0 for a, b in zip(runs[r][3][k], runs[0][3][k])]
        if max(diffs) > 0:
            print("   ", k, ["%.2e" % x for x in diffs])
            break
# head only, same inputs
with torch.no_grad():
    o1 = eng.head._run(runs[0][0], runs[0][1], (320, 320), True)
    o2 = eng.head._run(runs[0][0], runs[0][1], (320, 320), True)
print("head determinism cls", d(o1[0], o2[0]), "sig", d(o1[3], o2[3]))
