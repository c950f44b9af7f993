"""Where does the bf16 image tower's noise at c4 come from?  Inject the fp32 parity-mode activation (rounded once to bf16) at the
start of each stage, run the bf16 blocks from there on, report the relative error of c4 (train-mode BN)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
os.environ.setdefault("TRIS_ALLOW_RANDOM_INIT", "1")
from oracle import weights as W
from tris_b200.model_stage1 import TRIS
from tris_b200.precise import PreciseStage1
from tris_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4321
args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
m = TRIS(args); m.load_state_dict(W.make_tris_state_dict(0)); m = m.cuda().train()
eng = m.engine(); eng.ensure_fresh(True)
img, ids, _ = W.synthetic_batch(B, 320, 20, 3, seed)
img = img.cuda()
rn, pr = eng.resnet, PreciseStage1(m)
rn._join_packs()
err = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
with torch.no_grad():
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    p = rn.prefix
    xr = pr.bn(pr.conv3x3(img, p + "conv1.weight", stride=2, nchw=True), p + "bn1", True)
    xr = pr.bn(pr.conv3x3(xr, p + "conv2.weight"), p + "bn2", True)
    xr = pr.avgpool(pr.bn(pr.conv3x3(xr, p + "conv3.weight"), p + "bn3", True))
    refs = [xr]
    for blk in rn.blocks:
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
        refs.append(xr)
    c4r = refs[-1]

    def stats(c):
        return torch.empty(2 * c * ops.STAT_PARTS, device="cuda")
    full, _ = rn.forward(img, True)
    m.load_state_dict(sd0)
    print(f"B={B} seed={seed}: whole bf16 tower: c4 rel err {err(full.float(), c4r):.4f}")
    names = [b.p[len(p):] for b in rn.blocks]
    for start in range(len(rn.blocks)):
        if start and not rn.blocks[start].down:
            continue
        x = refs[start].to(torch.bfloat16).contiguous()
        for blk in rn.blocks[start:]:
            x, _ = rn._block_fwd(blk, x, True, stats)
        m.load_state_dict(sd0)
        print(f"  fp32 input injected in front of {names[start]:10s}: c4 rel err {err(x.float(), c4r):.4f}")
    # single stage in bf16, fp32 before and after: error at the stage's own output
    for start in range(len(rn.blocks)):
        if start and not rn.blocks[start].down:
            continue
        end = start + 1
        while end < len(rn.blocks) and not rn.blocks[end].down:
            end += 1
        x = refs[start].to(torch.bfloat16).contiguous()
        for blk in rn.blocks[start:end]:
            x, _ = rn._block_fwd(blk, x, True, stats)
        m.load_state_dict(sd0)
        print(f"  stage starting at {names[start]:10s} alone ({end - start} blocks): rel err at its output {err(x.float(), refs[end]):.4f}")
