"""Warm per-section device times of one Stage-1 training step (CUDA events; launches queued behind a spin kernel so the
events see back-to-back kernel execution; single stream, no overlap)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from bench import make_args
from tris_b200 import clip_model, ops
from tris_b200.model_stage1 import TRIS
from tris_b200.synthetic import synthetic_batch
from tris_b200.train_step import Stage1Trainer
bf16 = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
model = TRIS(make_args()).cuda().train()
with torch.no_grad():
    for k, p in model.named_parameters():
        if k.endswith("bn3.weight") and "layer" in k:
            p.uniform_(0.1, 0.3)
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
tr = Stage1Trainer(model, aux, max_iter=1000)
img, ids, negs = (t.cuda() for t in synthetic_batch(B, 320, 20, 3, 1))
eng, aeng = model.engine(), aux._engine()
eng.ensure_fresh(True); aeng.ensure_fresh(True)
ids_all = torch.cat([ids, negs.reshape(-1, 20)], 0)
sections, marks = [], []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
def run():
    marks.clear()
    with torch.no_grad():
        eng.store.zero_grad()
        mark("start")
        c4, tape = eng.resnet.forward(img, train=True); mark("RN50 forward (53 convs + BN)")
        hidden, trec, _ = eng.text.forward(ids, save=True); mark("TRIS text tower forward (M=960)")
        out, htape = eng.head._fwd(c4, hidden, (320, 320), True, save=True); mark("head forward (K6+K7+K8)")
        patches, _ = ops.mask_resize_fwd(out[3], img, 224, 32); mark("mask-resize forward (K9)")
        f, vrec = aeng.vit.forward(patches, B, save=True); mark("aux ViT-B/32 forward (M=2400)")
        g = aeng.text.forward(ids_all, save=False)[0]; mark("aux text tower forward (M=3840)")
        loss = ops.stage1_loss_fwd(f, g, out[0], 3, (1.0, 5.0, 2.0))
        dout = torch.tensor([1.0, 0, 0, 0], device="cuda")
        df, dcls = ops.stage1_loss_bwd(f, g, out[0], dout, 3, (1.0, 5.0, 2.0)); mark("loss forward+backward (K12)")
        dp = aeng.vit.backward(vrec, df); mark("aux ViT backward (dgrad only)")
        dsig = ops.mask_resize_bwd(dp, img, 224, 32); mark("mask-resize backward")
        dc4, dh = eng.head._bwd(htape, dcls, None, None, dsig, None); mark("head backward")
        eng.resnet.backward(tape, dc4.contiguous()); mark("RN50 backward (dgrad + wgrad + BN)")
        eng.text.backward(trec, dh.contiguous()); mark("TRIS text tower backward")
        tr.optimizer_step(); mark("fused AdamW (98.8 M params)")
import tris_b200.engine as E
E.OVERLAP = False
for _ in range(2):
    run()
torch.cuda.synchronize()
acc = {}
for it in range(3):
    torch.cuda._sleep(int(2.0e8))
    run()
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
        acc.setdefault(n1, []).append(e0.elapsed_time(e1))
tot = 0.0
for n, v in acc.items():
    m = sorted(v)[len(v) // 2]
    tot += m
    print(f"{m:8.3f} ms  {n}")
print(f"{tot:8.3f} ms  total (single stream, no overlap), batch {B}")
