"""A/B timing of the fused BatchNorm-backward reduction: producer GEMM (+/- bwd_stats epilogue) + bn_bwd (+/- ext) per layer shape."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G, ops
L.require_device()
bf16 = torch.bfloat16
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
def bn(c):
    mk = lambda: torch.rand(c, device="cuda") + 0.5
    b = ops.BNState(mk(), mk(), mk(), mk(), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"))
    b.mean.normal_(); b.invstd.fill_(1.0); b.scale.fill_(1.0); b.shift.normal_()
    return b
B = 48
print(f"{'layer':28s} {'gemm':>7s} {'gemm+st':>7s} {'bn_bwd':>7s} {'bn_ext':>7s} | {'unfused':>7s} {'fused':>7s}")
for name, h, cn, ck, conv in [("l1 bn2<-conv3 dgrad 256->64", 80, 256, 64, False), ("l2 bn2<-conv3 dgrad 512->128", 40, 512, 128, False),
                              ("l3 bn2<-conv3 dgrad 1024->256", 20, 1024, 256, False), ("l4 bn2<-conv3 dgrad 2048->512", 10, 2048, 512, False),
                              ("l1 bn3<-conv1 dgrad 64->256", 80, 64, 256, False), ("l3 bn3<-conv1 dgrad 256->1024", 20, 256, 1024, False),
                              ("stem bn<-3x3 dgrad 64->64 @160", 160, 64, 64, True), ("l1 bn1<-3x3 dgrad 64->64", 80, 64, 64, True),
                              ("l2 bn1<-3x3 dgrad 128", 40, 128, 128, True), ("l3 bn1<-3x3 dgrad 256", 20, 256, 256, True)]:
    nb = 24 if h == 160 else B
    M = nb * h * h
    y = rnd(nb, h, h, ck); b = bn(ck)
    parts = torch.empty(148 * 2 * ck + 2 * ck, device="cuda")
    if conv:
        dy, wp = rnd(nb, h, h, cn), rnd(cn, 9 * ck)
        out = torch.empty(nb, h, h, ck, device="cuda", dtype=bf16)
        t0 = run(lambda: G.conv3x3_dgrad(dy, wp, ck, out=out))
        t1 = run(lambda: G.conv3x3_dgrad(dy, wp, ck, out=out, bwd_stats=(parts, y, b.mean, b.scale, b.shift)))
    else:
        dy, w = rnd(M, cn), rnd(cn, ck)
        out = torch.empty(M, ck, device="cuda", dtype=bf16)
        t0 = run(lambda: G.linear_dgrad(dy, w, out=out))
        t1 = run(lambda: G.linear_dgrad(dy, w, out=out, bwd_stats=(parts, y.view(M, ck), b.mean, b.scale, b.shift)))
    o4 = out.view(nb, h, h, ck)
    t2 = run(lambda: ops.bn_bwd(o4, None, y, b))
    t3 = run(lambda: ops.bn_bwd(o4, None, y, b, ext=parts))
    print(f"{name:28s} {t0:7.1f} {t1:7.1f} {t2:7.1f} {t3:7.1f} | {t0 + t2:7.1f} {t1 + t3:7.1f}")
