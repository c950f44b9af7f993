"""Per-call CUDA-event timing of every tris_gemm launch in one eager training step (sorted by time)."""
import argparse, os, sys, warnings, json, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from bench import make_args
from tris_b200 import _lib as L, clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.synthetic import synthetic_batch
from tris_b200.train_step import Stage1Trainer

model = TRIS(make_args()).cuda().train()
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
tr = Stage1Trainer(model, aux, max_iter=1000)
batch = tuple(t.cuda() for t in synthetic_batch(48, 320, 20, 3, 1))
for _ in range(2):
    tr.step(*batch)
torch.cuda.synchronize()
calls = []
orig = L.gemm_raw
def timed(desc):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); orig(desc); e.record()
    taps = desc.taps if (desc.wgrad and desc.taps > 1) else 1
    calls.append((s, e, dict(M=desc.M, N=desc.N, K=desc.K, a=desc.a_mode, b=desc.b_mode, taps=desc.taps, wgrad=desc.wgrad,
                             bn=desc.block_n, split=desc.split_k, h=desc.img_h, th=desc.tile_h, tw=desc.tile_w,
                             stats=bool(desc.stats), f32=desc.out_dtype, atomic=desc.atomic), 2.0 * desc.M * desc.N * desc.K * taps))
L.gemm_raw = timed
tr.step(*batch)
torch.cuda.synchronize()
agg = collections.OrderedDict()
for s, e, d, fl in calls:
    key = json.dumps(d)
    ms = s.elapsed_time(e)
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += ms; a[2] += fl
tot = sum(a[1] for a in agg.values())
print(f"total gemm ms {tot:.3f}, calls {len(calls)}, TFLOP/s {sum(a[2] for a in agg.values())/tot/1e9:.1f}")
for key, (n, ms, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{ms:8.3f} ms x{n:<3d} {fl/ms/1e9:7.1f} TF/s  {key}")
