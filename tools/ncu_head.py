"""Every launch of the cross-modal fusion head (K6 + K7 + K8: vis/lan projection, cross-modal attention, response head),
forward and backward, at the benchmark batch (48 images x 100 pixels, 48 sentences) between cudaProfilerStart/Stop:
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_head python tools/ncu_head.py"""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
os.environ.setdefault("TRIS_ALLOW_RANDOM_INIT", "1")
from bench import make_args
from tris_b200.model_stage1 import TRIS
bf16 = torch.bfloat16
B = int(sys.argv[1]) if len(sys.argv) > 1 else 48
m = TRIS(make_args()).cuda().train()
eng = m.engine()
eng.ensure_fresh(True)
g = torch.Generator(device="cuda").manual_seed(0)
c4 = (torch.randn(B, 10, 10, 2048, generator=g, device="cuda").abs() * 0.5).to(bf16)
hidden = (torch.randn(B, 1024, generator=g, device="cuda") * 0.3).to(bf16)
dcls, dfg, dsig = torch.randn(B, B, device="cuda"), torch.randn(B, device="cuda"), torch.randn(B, 1, 320, 320, device="cuda") * 0.01


def run():
    eng.fwd_id += 1
    eng.store.zero_grad()
    c4g, hg = c4.clone().requires_grad_(True), hidden.clone().requires_grad_(True)
    cls, fg, relu, sig, es = eng.head.forward(c4g, hg, (320, 320), True)
    torch.autograd.backward([cls, fg, sig], [dcls, dfg, dsig])


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
