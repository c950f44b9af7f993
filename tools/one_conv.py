import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
x = rnd(48, 80, 80, 64); wp = rnd(64, 576); dy = rnd(48, 80, 80, 64)
out = torch.empty(48, 80, 80, 64, device="cuda", dtype=bf16)
dx = torch.empty_like(x)
for _ in range(2):
    G.conv3x3_dgrad(dy, wp, 64, out=dx)
    xx, ww = rnd(307200, 256), rnd(64, 256)
    oo = torch.empty(307200, 64, device="cuda", dtype=bf16)
    G.linear_fwd(xx, ww, out=oo)
torch.cuda.synchronize()
