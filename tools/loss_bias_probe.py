"""bf16 loss vs fp32 oracle (GPU-evaluated) for a few (batch, seed) pairs: bisect helper for precision regressions.
Usage: python tools/loss_bias_probe.py [B:seed ...]   (env knobs select the kernel variants)"""
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
model = TRIS(args); model.load_state_dict(W.make_tris_state_dict(0), strict=True); model = model.cuda().train()
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
aux.load_state_dict(W.make_vitb32_clip_state_dict(7, cos_bias=True), strict=True)
sd0 = {k: v.clone() for k, v in model.state_dict().items()}
pairs = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(4, 99), (48, 4321), (48, 1234)]
out = []
for B, seed in pairs:
    img, ids, negs = (t.cuda() for t in W.synthetic_batch(B, 320, 20, 3, seed))
    with torch.no_grad():
        sdc = {k: v.clone() for k, v in sd0.items()}
        auxc = {k: v.detach().clone() for k, v in aux.state_dict().items()}
        cls_out, _, _, sig_map, _ = O.tris_forward(sdc, img, ids, True, {})
        ref = O.stage1_losses(cls_out, sig_map, img, ids, negs, auxc)
        got = stage1_losses(model, aux, img, ids, negs)
    model.load_state_dict(sd0)
    out.append(f"{B}:{seed} " + " ".join(f"{k} {100 * (got[k].item() / ref[k].item() - 1):+.3f}%" for k in ("loss", "l1", "l4", "l5")))
knobs = {k: v for k, v in os.environ.items() if k.startswith("TRIS_") and k != "TRIS_ALLOW_RANDOM_INIT"}
print(knobs, " | ".join(out), flush=True)
