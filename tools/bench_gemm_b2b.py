"""Back-to-back launches (host overhead amortised): true per-kernel time of representative shapes."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
def run(name, fn, flops, n=30):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); g.replay(); e.record(); torch.cuda.synchronize()
    t = s.elapsed_time(e) / n
    print(f"{name:44s} {t*1e3:8.1f} us  {flops/t/1e9:7.1f} TF/s")
for name, m, n, k in [("lin 307200x256x64", 307200, 256, 64), ("lin 307200x64x256", 307200, 64, 256), ("lin 4800x3072x1024", 4800, 3072, 1024),
                      ("lin 2400x3072x768", 2400, 3072, 768), ("lin 2400x768x3072", 2400, 768, 3072), ("lin 960x512x512", 960, 512, 512),
                      ("lin 960x2048x512", 960, 2048, 512), ("lin 3840x2048x512", 3840, 2048, 512), ("lin 8192^3", 8192, 8192, 8192)]:
    x, w = rnd(m, k), rnd(n, k)
    out = torch.empty(m, n, device="cuda", dtype=bf16)
    run(name, lambda: G.linear_fwd(x, w, out=out), 2.0 * m * n * k, n=10 if m * n * k > 1e11 else 30)
for n_, h, ci, co in [(48, 80, 64, 64), (48, 40, 128, 128), (48, 20, 256, 256), (48, 10, 512, 512)]:
    x = rnd(n_, h, h, ci); wp = rnd(co, 9 * ci); dy = rnd(n_, h, h, co)
    out = torch.empty(n_, h, h, co, device="cuda", dtype=bf16)
    stats = torch.zeros(148 * 2 * co, device="cuda")
    fl = 2.0 * n_ * h * h * co * 9 * ci
    run(f"conv3x3 fwd+stats {h}x{h} {ci}", lambda: G.conv3x3_fwd(x, wp, stats=stats, out=out), fl)
    dx = torch.empty_like(x)
    run(f"conv3x3 dgrad {h}x{h} {ci}", lambda: G.conv3x3_dgrad(dy, wp, ci, out=dx), fl)
    gw = torch.zeros(co, 9 * ci, device="cuda")
    run(f"conv3x3 wgrad {h}x{h} {ci}", lambda: G.conv3x3_wgrad(dy, x, out=gw), fl)
dy, x = rnd(960, 2048), rnd(960, 512)
gw = torch.zeros(2048, 512, device="cuda")
run("lin wgrad 2048x512 K=960", lambda: G.linear_wgrad(dy, x, out=gw), 2.0 * 960 * 2048 * 512)
