"""Which side makes the bf16 classification term (l4) deviate from fp32: the towers or the fusion head?
l4 from (product towers, product head) / (product towers, fp32 oracle head) / (fp32 oracle towers rounded once to bf16, product head)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
warnings.simplefilter("ignore")
os.environ.setdefault("TRIS_ALLOW_RANDOM_INIT", "1")
from oracle import tris_oracle as O
from oracle import weights as W
from tris_b200.model_stage1 import TRIS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
bf16 = torch.bfloat16
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
sd = W.make_tris_state_dict(0)
m = TRIS(args); m.load_state_dict(sd); m = m.cuda().train(); eng = m.engine(); eng.ensure_fresh(True)
sdc = {k: v.cuda() for k, v in sd.items()}
sd0 = {k: v.clone() for k, v in m.state_dict().items()}
l4 = lambda cls: F.multilabel_soft_margin_loss(cls.float(), torch.eye(cls.shape[0], device=cls.device)).item()
for a in (sys.argv[1:] or ["48:4321", "48:1234", "48:7", "4:99"]):
    B, seed = (int(v) for v in a.split(":"))
    img, ids, negs = W.synthetic_batch(B, 320, 20, 3, seed)
    with torch.no_grad():
        _, hid_ref = O.encode_text(sdc, ids.cuda(), prefix="backbone.")
        c4_ref = O.resnet_tower(sdc, img.cuda(), prefix="backbone.visual.", train=True, new_stats={})[-1]
        ref = l4(O.tris_head(O.tris_score(sdc, c4_ref, hid_ref), (10, 10), (320, 320), True)["cls_out"])
        c4, hidden = eng.towers(img.cuda(), ids.cuda(), True)
        m.load_state_dict(sd0)
        c4n = ((c4.float().permute(0, 3, 1, 2) - c4_ref).norm() / c4_ref.norm()).item()
        hn = ((hidden.float() - hid_ref).norm() / hid_ref.norm()).item()
        pp = l4(eng.head.forward(c4, hidden, (320, 320), True)[0])
        op = l4(eng.head.forward(c4_ref.permute(0, 2, 3, 1).contiguous().to(bf16), hid_ref.to(bf16), (320, 320), True)[0])
        po = l4(O.tris_head(O.tris_score(sdc, c4.float().permute(0, 3, 1, 2), hidden.float()), (10, 10), (320, 320), True)["cls_out"])
        pto = l4(O.tris_head(O.tris_score(sdc, c4.float().permute(0, 3, 1, 2), hid_ref), (10, 10), (320, 320), True)["cls_out"])
    pc = lambda v: f"{100 * (v / ref - 1):+.3f}%"
    print(f"{a}: c4 noise {c4n:.4f} hidden noise {hn:.5f} | l4 fp32 {ref:.4f} | product {pc(pp)} | product towers + fp32 head {pc(po)} "
          f"(image tower only {pc(pto)}) | fp32 towers + product head {pc(op)}", flush=True)
