"""Micro-benchmark of representative tris_gemm shapes (CUDA events, L2 flushed between runs); also the ncu target."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
reps = int(os.environ.get("REPS", "5"))
def run(name, fn, flops, bytes_):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {t*1e3:8.1f} us  {flops/t/1e9:7.1f} TF/s  {bytes_/t/1e6:7.1f} GB/s")
cases = [("lin 307200x256x64 +stats", 307200, 256, 64, True), ("lin 307200x64x256 +stats", 307200, 64, 256, True),
         ("lin 76800x512x128 +stats", 76800, 512, 128, True), ("lin 19200x1024x256 +stats", 19200, 1024, 256, True),
         ("lin 4800x3072x1024", 4800, 3072, 1024, False), ("lin 2400x3072x768", 2400, 3072, 768, False),
         ("lin 960x512x512", 960, 512, 512, False), ("lin 3840x2048x512", 3840, 2048, 512, False),
         ("lin 8192x8192x8192", 8192, 8192, 8192, False)]
for name, m, n, k, st in cases:
    x, w = rnd(m, k), rnd(n, k)
    out = torch.empty(m, n, device="cuda", dtype=bf16)
    stats = torch.zeros(148 * 2 * n, device="cuda") if st else None
    run(name, lambda: G.linear_fwd(x, w, out=out, stats=stats), 2.0 * m * n * k, 2.0 * (m * k + n * k + m * n))
for n_, h, ci, co in [(48, 80, 64, 64), (48, 40, 128, 128), (48, 20, 256, 256), (48, 10, 512, 512)]:
    x = rnd(n_, h, h, ci); wp = rnd(co, 9 * ci); dy = rnd(n_, h, h, co)
    out = torch.empty(n_, h, h, co, device="cuda", dtype=bf16)
    stats = torch.zeros(148 * 2 * co, device="cuda")
    fl = 2.0 * n_ * h * h * co * 9 * ci
    by = 2.0 * (x.numel() + wp.numel() + out.numel())
    run(f"conv3x3 fwd {n_}x{h}x{h} {ci}->{co}", lambda: G.conv3x3_fwd(x, wp, stats=stats, out=out), fl, by)
    dx = torch.empty_like(x)
    run(f"conv3x3 dgrad {n_}x{h}x{h} {co}->{ci}", lambda: G.conv3x3_dgrad(dy, wp, ci, out=dx), fl, by)
    gw = torch.zeros(co, 9 * ci, device="cuda")
    run(f"conv3x3 wgrad {n_}x{h}x{h}", lambda: G.conv3x3_wgrad(dy, x, out=gw), fl, by)
dy, x = rnd(307200, 256), rnd(307200, 64)
gw = torch.zeros(256, 64, device="cuda")
run("lin wgrad 256x64 K=307200", lambda: G.linear_wgrad(dy, x, out=gw), 2.0 * 307200 * 256 * 64, 2.0 * 307200 * 320)
dy, w = rnd(307200, 256), rnd(256, 64)
dx = torch.empty(307200, 64, device="cuda", dtype=bf16)
run("lin dgrad 307200x64 K=256", lambda: G.linear_dgrad(dy, w, out=dx), 2.0 * 307200 * 256 * 64, 2.0 * 307200 * 320)
