"""bf16 tower vs fp32 parity-mode tower: relative error after the stem and after every bottleneck (train-mode BN)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from oracle import weights as W
from tris_b200.model_stage1 import TRIS
from tris_b200.precise import PreciseStage1
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
m = TRIS(args); m.load_state_dict(W.make_tris_state_dict(0)); m = m.cuda().train()
eng = m.engine(); eng.ensure_fresh(True)
img, ids, _ = W.synthetic_batch(B, 320, 20, 3, 4321)
img = img.cuda()
rn, pr = eng.resnet, PreciseStage1(m)
def frob(a, b):
    a, b = a.double(), b.double()
    proj = (a * b).sum() / (b * b).sum()          # least-squares gain of the bf16 tensor on the fp32 one (1 = unbiased)
    return f"{((a - b).norm() / b.norm()).item():.4f}  gain {proj.item():.5f}  mean ratio {(a.mean() / b.mean()).item():.5f}"
with torch.no_grad():
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    so = [0]; rn.stats_buf.zero_()
    def stats(c):
        s = rn.stats_buf[so[0]: so[0] + 2 * c]; so[0] += 2 * c; return s
    x, _ = (rn._stem_fwd_pair if B % 2 == 0 else rn._stem_fwd_padded)(img, True, stats)
    p = rn.prefix
    xr = pr.bn(pr.conv3x3(img, p + "conv1.weight", stride=2, nchw=True), p + "bn1", True)
    xr = pr.bn(pr.conv3x3(xr, p + "conv2.weight"), p + "bn2", True)
    xr = pr.avgpool(pr.bn(pr.conv3x3(xr, p + "conv3.weight"), p + "bn3", True))
    print(f"stem        {frob(x.float(), xr)}")
    for blk in rn.blocks:
        x, _ = rn._block_fwd(blk, x, True, stats)
        q = blk.p
        o = pr.bn(pr.conv1x1(xr, q + "conv1.weight"), q + "bn1", True)
        o = pr.bn(pr.conv3x3(o, q + "conv2.weight"), q + "bn2", True)
        if blk.stride > 1: o = pr.avgpool(o)
        y3 = pr.conv1x1(o, q + "conv3.weight")
        if blk.down:
            idt = pr.avgpool(xr) if blk.stride > 1 else xr
            xr = pr.bn(y3, q + "bn3", True, y1=pr.conv1x1(idt, q + "downsample.0.weight"), key1=q + "downsample.1")
        else:
            xr = pr.bn(y3, q + "bn3", True, res=xr)
        print(f"{q[len(p):]:12s}{frob(x.float(), xr)}")
    m.load_state_dict(sd0)
