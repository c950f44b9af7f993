"""Representative single launches for `ncu --set full` (one of each, after a warm-up call, between profiler start/stop)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G, ops
L.require_device()
bf16 = torch.bfloat16
def rnd(*s): return torch.randn(*s, device="cuda").to(bf16)
targets = []
# K7: fused Q|K|V projection of the cross-modal attention  [4800,1024] x [3072,1024]^T
x, w = rnd(4800, 1024), rnd(3072, 1024); o1 = torch.empty(4800, 3072, device="cuda", dtype=bf16)
targets.append(lambda: G.linear_fwd(x, w, out=o1))
# layer3 1x1 conv + BN statistics  [19200,256] x [1024,256]^T
x2, w2 = rnd(19200, 256), rnd(1024, 256); o2 = torch.empty(19200, 1024, device="cuda", dtype=bf16); st2 = torch.zeros(148 * 2048, device="cuda")
targets.append(lambda: G.linear_fwd(x2, w2, out=o2, stats=st2))
# layer2 3x3 conv forward / dgrad / wgrad  48x40x40x128
xc, wc, dyc = rnd(48, 40, 40, 128), rnd(128, 9 * 128), rnd(48, 40, 40, 128)
oc = torch.empty_like(xc); stc = torch.zeros(148 * 256, device="cuda"); gwc = torch.zeros(128, 9 * 128, device="cuda")
targets.append(lambda: G.conv3x3_fwd(xc, wc, stats=stc, out=oc))
targets.append(lambda: G.conv3x3_dgrad(dyc, wc, 128, out=oc))
targets.append(lambda: G.conv3x3_wgrad(dyc, xc, out=gwc))
# layer1 3x3 conv (halo re-use, weight-stationary) forward / dgrad / wgrad  48x80x80x64
xh, wh, dyh = rnd(48, 80, 80, 64), rnd(64, 9 * 64), rnd(48, 80, 80, 64)
oh = torch.empty_like(xh); sth = torch.zeros(148 * 128, device="cuda"); gwh = torch.zeros(64, 9 * 64, device="cuda")
targets.append(lambda: G.conv3x3_fwd(xh, wh, stats=sth, out=oh))
targets.append(lambda: G.conv3x3_dgrad(dyh, wh, 64, out=oh))
targets.append(lambda: G.conv3x3_wgrad(dyh, xh, out=gwh))
# layer1 1x1 conv 64 -> 256 + BN statistics (K = 64: HBM-bound, two-group epilogue)  [307200,64] x [256,64]^T
x3, w3 = rnd(307200, 64), rnd(256, 64); o3 = torch.empty(307200, 256, device="cuda", dtype=bf16); st3 = torch.zeros(148 * 512, device="cuda")
targets.append(lambda: G.linear_fwd(x3, w3, out=o3, stats=st3))
# BatchNorm backward on the largest bottleneck tensor  48x80x80x256
y = rnd(48, 80, 80, 256); dout = rnd(48, 80, 80, 256); outf = torch.relu(y)
mk = lambda: torch.rand(256, device="cuda") + 0.5
bn = ops.BNState(mk(), mk(), mk(), mk(), torch.zeros(256, device="cuda"), torch.zeros(256, device="cuda"))
stats = torch.zeros(148, 512, device="cuda")
stats[0] = torch.cat([y.float().reshape(-1, 256).sum(0), (y.float() ** 2).reshape(-1, 256).sum(0)])
stats = stats.reshape(-1)
targets.append(lambda: ops.bn_apply(y, stats, bn, True))
targets.append(lambda: ops.bn_bwd(dout, outf, y, bn))
for t in targets:
    t()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in targets:
    t()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
