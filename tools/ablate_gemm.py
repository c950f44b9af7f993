"""Ablation timing of the GEMM kernel on the in-step RN50 shapes: TRIS_GEMM_DEBUG bit 1 skips the TMA loads, 2 the MMA issue,
4 the epilogue work (results are garbage; only the time matters).  One process per mask (the mask is read once)."""
import os, sys, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
warnings.simplefilter("ignore")
from tris_b200 import _lib as L, gemm as G
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
mask = os.environ.get("TRIS_GEMM_DEBUG", "0")
res = []
def conv(tag, n_, h, ci, co):
    x = rnd(n_, h, h, ci); wp = rnd(co, 9 * ci); dy = rnd(n_, h, h, co)
    out = torch.empty(n_, h, h, co, device="cuda", dtype=bf16); st = torch.zeros(148 * 2 * co, device="cuda")
    dx = torch.empty_like(x); gw = torch.zeros(co, 9 * ci, device="cuda")
    res.append((tag + " fwd", run(lambda: G.conv3x3_fwd(x, wp, stats=st, out=out))))
    res.append((tag + " dgrad", run(lambda: G.conv3x3_dgrad(dy, wp, ci, out=dx))))
    res.append((tag + " wgrad", run(lambda: G.conv3x3_wgrad(dy, x, out=gw))))
def lin(tag, M, ci, co):
    x, w, dy = rnd(M, ci), rnd(co, ci), rnd(M, co)
    out = torch.empty(M, co, device="cuda", dtype=bf16); st = torch.zeros(148 * 2 * co, device="cuda")
    dx = torch.empty(M, ci, device="cuda", dtype=bf16); gw = torch.zeros(co, ci, device="cuda")
    res.append((tag + " fwd", run(lambda: G.linear_fwd(x, w, out=out, stats=st))))
    res.append((tag + " dgrad", run(lambda: G.linear_dgrad(dy, w, out=dx))))
    res.append((tag + " wgrad", run(lambda: G.linear_wgrad(dy, x, out=gw, accumulate=True))))
conv("stem2 24x160 64->64", 24, 160, 64, 64)
conv("stem3 24x160 64->128", 24, 160, 64, 128)
conv("l1 48x80 64->64", 48, 80, 64, 64)
conv("l2 48x40 128->128", 48, 40, 128, 128)
conv("l3 48x20 256->256", 48, 20, 256, 256)
conv("l4 48x10 512->512", 48, 10, 512, 512)
lin("l1 64->256", 307200, 64, 256)
lin("l1 256->64", 307200, 256, 64)
lin("l2 128->512", 76800, 128, 512)
lin("l3 256->1024", 19200, 256, 1024)
lin("l3 1024->256", 19200, 1024, 256)
lin("l4 512->2048", 4800, 512, 2048)
print("mask", mask, " ".join(f"{t:.1f}" for _, t in res))
if mask == "0":
    print("names", "|".join(n for n, _ in res))
