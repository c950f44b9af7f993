import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
def run(name, fn):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    print(f"dbg={os.environ.get('TRIS_GEMM_DEBUG','0'):3s} stages={os.environ.get('TRIS_GEMM_STAGES','-'):2s} {name:28s} {sorted(ts)[2]*1e3:8.1f} us")
x, w = rnd(307200, 256), rnd(64, 256); o = torch.empty(307200, 64, device="cuda", dtype=bf16)
run("lin 307200x64x256 bn64", lambda: G.linear_fwd(x, w, out=o))
x2, w2 = rnd(76800, 1024), rnd(256, 1024); o2 = torch.empty(76800, 256, device="cuda", dtype=bf16)
run("lin 76800x256x1024 bn256", lambda: G.linear_fwd(x2, w2, out=o2, block_n=256))
run("lin 76800x256x1024 bn64", lambda: G.linear_fwd(x2, w2, out=o2, block_n=64))
