"""Isolated timing of the stem 3x3 convs: current 64-padded form vs image-pair-packed form."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
def run(name, fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    print(f"{name:50s} {s.elapsed_time(e) / n * 1e3:8.1f} us")
for tag, n_, ci, co in [("cur  conv2/3 48x160x160 64->64", 48, 64, 64), ("pack conv2 24x160x160 64->64", 24, 64, 64), ("pack conv3 24x160x160 64->128", 24, 64, 128),
                        ("pack4 conv2 12x160x160 128->128", 12, 128, 128)]:
    x = rnd(n_, 160, 160, ci); wp = rnd(co, 9 * ci); dy = rnd(n_, 160, 160, co)
    out = torch.empty(n_, 160, 160, co, device="cuda", dtype=bf16); st = torch.zeros(148 * 2 * co, device="cuda")
    run(tag + " fwd+stats", lambda: G.conv3x3_fwd(x, wp, stats=st, out=out))
    dx = torch.empty_like(x)
    run(tag + " dgrad", lambda: G.conv3x3_dgrad(dy, wp, ci, out=dx))
    gw = torch.zeros(co, 9 * ci, device="cuda")
    run(tag + " wgrad", lambda: G.conv3x3_wgrad(dy, x, out=gw))
