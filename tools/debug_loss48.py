"""Loss of the product vs the fp32 oracle (run on CUDA) at the benchmark batch size."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import tris_oracle as O
from oracle import weights as W
from tris_b200 import clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.train_step import stage1_losses
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0); aux_sd = W.make_vitb32_clip_state_dict(7, cos_bias=True)
m = TRIS(args); m.load_state_dict(sd); m = m.cuda().train()
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True); aux.load_state_dict(aux_sd, strict=True)
img, ids, negs = W.synthetic_batch(B, 320, 20, 3, 1234)
sdc = {k: v.cuda() for k, v in sd.items()}; auxc = {k: v.cuda() for k, v in aux_sd.items()}
with torch.no_grad():
    cls_out, cls_fg, relu_map, sig_map, ls = O.tris_forward(sdc, img.cuda(), ids.cuda(), True, {})
    ref = O.stage1_losses(cls_out, sig_map, img.cuda(), ids.cuda(), negs.cuda(), auxc)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    for i in range(3):
        m.load_state_dict(sd0)
        got = stage1_losses(m, aux, img.cuda(), ids.cuda(), negs.cuda())
        print(B, "loss", got["loss"].item(), ref["loss"].item(), "rel", abs(got["loss"].item() / ref["loss"].item() - 1),
              "l1", got["l1"].item(), ref["l1"].item(), "l4", got["l4"].item(), ref["l4"].item(), "l5", got["l5"].item(), ref["l5"].item())
