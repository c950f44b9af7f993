"""Run one eager Stage-1 training step between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import argparse, os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from bench import make_args
from tris_b200 import clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.synthetic import synthetic_batch
from tris_b200.train_step import Stage1Trainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=48)
ap.add_argument("--warm", type=int, default=2)
a = ap.parse_args()
model = TRIS(make_args()).cuda().train()
with torch.no_grad():
    for k, p in model.named_parameters():
        if k.endswith("bn3.weight") and "layer" in k:
            p.uniform_(0.1, 0.3)
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
tr = Stage1Trainer(model, aux, max_iter=1000)
batch = tuple(t.cuda() for t in synthetic_batch(a.batch, 320, 20, 3, 1))
for _ in range(a.warm):
    tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(*batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
