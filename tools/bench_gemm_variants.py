import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def run(name, fn, bytes_):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[2]
    print(f"{name:50s} {t*1e3:8.1f} us {bytes_/t/1e6:7.1f} GB/s")
m, n, k = 307200, 256, 64
x, w = rnd(m, k), rnd(n, k)
out = torch.empty(m, n, device="cuda", dtype=bf16)
stats = torch.zeros(2 * n, device="cuda")
by = 2.0 * (m * k + n * k + m * n)
for bn in (256, 128, 64):
    run(f"307200x256x64 bn={bn} stats", lambda: G.linear_fwd(x, w, out=out, stats=stats, block_n=bn), by)
    run(f"307200x256x64 bn={bn} nostats", lambda: G.linear_fwd(x, w, out=out, block_n=bn), by)
