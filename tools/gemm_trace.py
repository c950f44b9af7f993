import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
buf = torch.zeros(7 * 64, dtype=torch.int64, device="cuda")
os.environ["TRIS_GEMM_DBGBUF"] = str(buf.data_ptr())
os.environ["TRIS_GEMM_DEBUG"] = os.environ.get("DBG", "8")
from tris_b200 import _lib as L, gemm as G
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
which = sys.argv[1]
if which == "conv":
    dy = rnd(48, 80, 80, 64); wp = rnd(64, 576); dx = torch.empty(48, 80, 80, 64, device="cuda", dtype=bf16)
    fn = lambda: G.conv3x3_dgrad(dy, wp, 64, out=dx)
elif which == "small":
    x, w = rnd(960, 512), rnd(512, 512); o = torch.empty(960, 512, device="cuda", dtype=bf16)
    fn = lambda: G.linear_fwd(x, w, out=o)
elif which == "mid":
    x, w = rnd(2400, 768), rnd(3072, 768); o = torch.empty(2400, 3072, device="cuda", dtype=bf16)
    fn = lambda: G.linear_fwd(x, w, out=o)
else:
    x, w = rnd(307200, 256), rnd(64, 256); o = torch.empty(307200, 64, device="cuda", dtype=bf16)
    fn = lambda: G.linear_fwd(x, w, out=o)
fn(); torch.cuda.synchronize(); buf.zero_(); fn(); torch.cuda.synchronize()
b = buf.cpu().view(7, 64)
t0 = int(b[b > 0].min())
names = ["prodA tile", "prodB tile", "mma pre-wait", "mma post-wait", "epi pre-wait", "epi post-wait", "epi pre-store"]
for i in range(7):
    row = [(int(v) - t0) for v in b[i][:18]]
    print(f"{names[i]:14s}", row)
