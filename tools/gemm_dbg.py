import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
def run(name, fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"dbg={os.environ.get('TRIS_GEMM_DEBUG','0')} {name:40s} {sorted(ts)[2]*1e3:8.1f} us")
dy = rnd(48, 80, 80, 64); wp = rnd(64, 576); dx = torch.empty(48, 80, 80, 64, device="cuda", dtype=bf16)
run("conv3x3 dgrad 80x80 64", lambda: G.conv3x3_dgrad(dy, wp, 64, out=dx))
x, w = rnd(307200, 256), rnd(64, 256); o = torch.empty(307200, 64, device="cuda", dtype=bf16)
run("lin 307200x64x256", lambda: G.linear_fwd(x, w, out=o))
x, w = rnd(8192, 8192), rnd(8192, 8192); o = torch.empty(8192, 8192, device="cuda", dtype=bf16)
run("lin 8192^3", lambda: G.linear_fwd(x, w, out=o))
x, w = rnd(2400, 768), rnd(3072, 768); o = torch.empty(2400, 3072, device="cuda", dtype=bf16)
run("lin 2400x3072x768", lambda: G.linear_fwd(x, w, out=o))
