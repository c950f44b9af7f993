"""Where does the cls_out bias come from?  Product towers/head vs oracle (fp32 on CUDA), golden config (B=3)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import tris_oracle as O
from oracle import weights as W
from tris_b200.model_stage1 import TRIS
bf16 = torch.bfloat16
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0)
m = TRIS(args); m.load_state_dict(sd); m = m.cuda().train(); eng = m.engine(); eng.ensure_fresh(True)
img, ids, negs = W.synthetic_batch(3, 320, 20, 3, 1234)
sdc = {k: v.cuda() for k, v in sd.items()}
with torch.no_grad():
    _, hid_ref = O.encode_text(sdc, ids.cuda(), prefix="backbone.")
    c4_ref = O.resnet_tower(sdc, img.cuda(), prefix="backbone.visual.", train=True, new_stats={})[-1]
    score_ref = O.tris_score(sdc, c4_ref, hid_ref)
    o_ref = O.tris_head(score_ref, (10, 10), (320, 320), True)
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    c4, hidden = eng.towers(img.cuda(), ids.cuda(), True)
    m.load_state_dict(sd0)
    print("c4 rel", ((c4.float().permute(0,3,1,2) - c4_ref).norm() / c4_ref.norm()).item(), "hidden rel", ((hidden.float() - hid_ref).norm() / hid_ref.norm()).item())
    out = eng.head.forward(c4, hidden, (320, 320), True)
    print("cls product towers+head:\n", out[0].cpu().numpy(), "\nref:\n", o_ref["cls_out"].cpu().numpy())
    out2 = eng.head.forward(c4_ref.permute(0,2,3,1).contiguous().to(bf16), hid_ref.to(bf16), (320, 320), True)
    print("cls oracle towers + product head:\n", out2[0].cpu().numpy())
    sc2 = O.tris_score(sdc, c4.float().permute(0,3,1,2), hidden.float())
    print("cls product towers + oracle head:\n", O.tris_head(sc2, (10,10), (320,320), True)["cls_out"].cpu().numpy())
