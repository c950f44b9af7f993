"""Back-to-back (CUDA graph) timing of every distinct RN50 conv shape at batch 48, 320x320: fwd(+stats) / dgrad / wgrad,
against the per-layer roofline max(flops / bf16 peak, compulsory bytes / HBM peak)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
PEAK_F, PEAK_B = 1388e12, 6.55e12
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
def run(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
B = 48
# (name, count, k, Cin, Cout, H)
layers = [("stem 64p->64p 3x3", 2, 3, 64, 64, 160),
          ("l1 64->64 1x1", 1, 1, 64, 64, 80), ("l1 64->64 3x3", 3, 3, 64, 64, 80), ("l1 64->256 1x1", 4, 1, 64, 256, 80), ("l1 256->64 1x1", 2, 1, 256, 64, 80),
          ("l2 256->128 1x1", 1, 1, 256, 128, 80), ("l2 128->128 3x3 @80", 1, 3, 128, 128, 80), ("l2 128->512 1x1", 4, 1, 128, 512, 40),
          ("l2 256->512 ds", 1, 1, 256, 512, 40), ("l2 512->128 1x1", 3, 1, 512, 128, 40), ("l2 128->128 3x3", 3, 3, 128, 128, 40),
          ("l3 512->256 1x1", 1, 1, 512, 256, 40), ("l3 256->256 3x3 @40", 1, 3, 256, 256, 40), ("l3 256->1024 1x1", 6, 1, 256, 1024, 20),
          ("l3 512->1024 ds", 1, 1, 512, 1024, 20), ("l3 1024->256 1x1", 5, 1, 1024, 256, 20), ("l3 256->256 3x3", 5, 3, 256, 256, 20),
          ("l4 1024->512 1x1", 1, 1, 1024, 512, 20), ("l4 512->512 3x3 @20", 1, 3, 512, 512, 20), ("l4 512->2048 1x1", 3, 1, 512, 2048, 10),
          ("l4 1024->2048 ds", 1, 1, 1024, 2048, 10), ("l4 2048->512 1x1", 2, 1, 2048, 512, 10), ("l4 512->512 3x3", 2, 3, 512, 512, 10)]
tot = {"fwd": 0.0, "dgrad": 0.0, "wgrad": 0.0, "ideal": 0.0}
print(f"{'layer':24s} {'n':>2s} {'fwd us':>8s} {'dgrad':>8s} {'wgrad':>8s} {'ideal':>7s}  fwd TF/s")
for name, cnt, k, ci, co, h in layers:
    M = B * h * h
    fl = 2.0 * M * co * ci * k * k
    by = 2.0 * (M * ci + M * co + co * ci * k * k)
    ideal = max(fl / PEAK_F, by / PEAK_B) * 1e6
    if k == 1:
        x, w, dy = rnd(M, ci), rnd(co, ci), rnd(M, co)
        out = torch.empty(M, co, device="cuda", dtype=bf16); st = torch.zeros(148 * 2 * co, device="cuda")
        dx = torch.empty(M, ci, device="cuda", dtype=bf16); gw = torch.zeros(co, ci, device="cuda")
        tf = run(lambda: G.linear_fwd(x, w, out=out, stats=st))
        td = run(lambda: G.linear_dgrad(dy, w, out=dx))
        tw = run(lambda: G.linear_wgrad(dy, x, out=gw, accumulate=True))
    else:
        x, wp, dy = rnd(B, h, h, ci), rnd(co, 9 * ci), rnd(B, h, h, co)
        out = torch.empty(B, h, h, co, device="cuda", dtype=bf16); st = torch.zeros(148 * 2 * co, device="cuda")
        dx = torch.empty_like(x); gw = torch.zeros(co, 9 * ci, device="cuda")
        tf = run(lambda: G.conv3x3_fwd(x, wp, stats=st, out=out))
        td = run(lambda: G.conv3x3_dgrad(dy, wp, ci, out=dx))
        tw = run(lambda: G.conv3x3_wgrad(dy, x, out=gw))
    print(f"{name:24s} {cnt:2d} {tf:8.1f} {td:8.1f} {tw:8.1f} {ideal:7.1f}  {fl / tf / 1e6:7.0f}")
    tot["fwd"] += cnt * tf; tot["dgrad"] += cnt * td; tot["wgrad"] += cnt * tw; tot["ideal"] += cnt * ideal
print({k: round(v / 1e3, 3) for k, v in tot.items()}, "ms per step (sum over layers x count)")
