"""Kernel timeline of ONE CUDA-graph replay of the training step (torch.profiler / CUPTI activity records): per kernel
start, duration, stream.  Diagnostic only (numbers under a profiler are never bench values): shows what is on the critical
path, where streams idle, and the per-kernel-name totals inside the real overlapped step."""
import json, os, sys, warnings, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
import bench
from tris_b200 import _lib as L, clip_model
from tris_b200.model_stage1 import TRIS
from tris_b200.synthetic import synthetic_batch
from tris_b200.train_step import Stage1Trainer
L.require_device()
torch.manual_seed(1234)
B = 48
model = TRIS(bench.make_args()).cuda().train()
with torch.no_grad():
    for k, p in model.named_parameters():
        if k.endswith("bn3.weight") and "layer" in k:
            p.uniform_(0.1, 0.3)
aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
trainer = Stage1Trainer(model, aux, max_iter=100000)
dev = tuple(t.cuda() for t in synthetic_batch(B, 320, 20, 3, seed=1234))
trainer.step(*dev); trainer.capture(*dev, warmup=1)
for _ in range(5): trainer.step(*dev)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        trainer.step(*dev)
    torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_trace.json"
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# split into replays by the adamw kernel
ends = [i for i, e in enumerate(ev) if "adamw" in e["name"]]
lo = ends[-2] + 1 if len(ends) >= 2 else 0
step = [e for e in ev[lo: ends[-1] + 1] if "tick" not in e["name"]]
t0 = step[0]["ts"]
with open(out.replace(".json", "_timeline.txt"), "w") as f:
    for e in step:
        f.write(f"{e['ts'] - t0:10.1f} {e['dur']:8.1f} s{e['args'].get('stream', -1):<4} {e['name'][:90]}\n")
tot = collections.defaultdict(lambda: [0, 0.0])
for e in step:
    k = e["name"].split("(")[0][-70:]
    tot[k][0] += 1; tot[k][1] += e["dur"]
span = step[-1]["ts"] + step[-1]["dur"] - t0
print(f"step span {span / 1e3:.3f} ms, {len(step)} kernels, sum of durations {sum(e['dur'] for e in step) / 1e3:.3f} ms")
# busy time: union of intervals
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in step)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
for s_, e_ in iv[1:]:
    if s_ > cur_e:
        busy += cur_e - cur_s; cur_s, cur_e = s_, e_
    else:
        cur_e = max(cur_e, e_)
busy += cur_e - cur_s
print(f"GPU busy (any kernel running) {busy / 1e3:.3f} ms, idle inside the step {(span - busy) / 1e3:.3f} ms")
for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{d / 1e3:8.3f} ms x{n:<4} {k}")
os.remove(out)
