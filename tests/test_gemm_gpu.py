"""GPU parity of the tcgen05 GEMM / implicit-conv kernel (tris_gemm) against torch fp32 on the same
bf16-rounded operands.  bf16 outputs: <= 1e-2 rel; fp32 outputs: <= 2e-3 rel (accumulation order only)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.fixture(scope="module")
def G():
    from tris_b200 import _lib, gemm
    _lib.require_device()
    return gemm


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(torch.bfloat16)


@pytest.mark.parametrize("m,n,k", [(128, 128, 64), (960, 1536, 512), (4800, 3072, 1024), (48, 1024, 1024),
                                    (2400, 768, 3072), (307200, 64, 64), (1000, 136, 200)])
def test_linear_fwd(G, m, n, k):
    x, w = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=k ** -0.5)
    bias = torch.randn(n, device="cuda")
    ref = x.float() @ w.float().t() + bias
    out = G.linear_fwd(x, w, bias)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-2
    out32 = G.linear_fwd(x, w, bias, out_dtype=torch.float32)
    assert rel(out32, ref) < 2e-3


def test_linear_fwd_epilogue(G):
    from tris_b200 import _lib as L
    m, n, k = 960, 2048, 512
    x, w = rnd(m, k, seed=1), rnd(n, k, seed=2, scale=k ** -0.5)
    bias = torch.randn(n, device="cuda")
    res = rnd(m, n, seed=3)
    pre = x.float() @ w.float().t() + bias
    out = G.linear_fwd(x, w, bias, act=L.ACT_QUICKGELU, residual=res)
    assert rel(out, pre * torch.sigmoid(1.702 * pre) + res.float()) < 1e-2
    stats = torch.zeros(2 * n, device="cuda")
    out = G.linear_fwd(x, w, bias, act=L.ACT_RELU, stats=stats)
    assert rel(out, F.relu(pre)) < 1e-2
    assert rel(stats[:n], pre.sum(0)) < 2e-3
    assert rel(stats[n:], (pre * pre).sum(0)) < 2e-3


@pytest.mark.parametrize("m,n,k", [(960, 1536, 512), (4800, 1024, 2048), (48, 1024, 1024), (2400, 3072, 768),
                                    (307200, 256, 64)])
def test_linear_dgrad(G, m, n, k):
    dy, w = rnd(m, n, seed=1), rnd(n, k, seed=2, scale=n ** -0.5)
    out = G.linear_dgrad(dy, w)
    assert rel(out, dy.float() @ w.float()) < 1e-2


@pytest.mark.parametrize("m,n,k", [(960, 1536, 512), (4800, 1024, 2048), (48, 1024, 1024), (2400, 3072, 768),
                                    (307200, 64, 256), (19200, 1024, 256)])
def test_linear_wgrad(G, m, n, k):
    dy, x = rnd(m, n, seed=1, scale=m ** -0.5), rnd(m, k, seed=2)
    out = G.linear_wgrad(dy, x)
    assert rel(out, dy.float().t() @ x.float()) < 2e-3
    out1 = G.linear_wgrad(dy, x, split_k=1)
    assert rel(out1, dy.float().t() @ x.float()) < 2e-3


CONV_CASES = [(2, 80, 80, 64, 64), (3, 40, 40, 128, 128), (2, 20, 20, 256, 256), (3, 10, 10, 512, 512),
              (1, 20, 20, 512, 512), (2, 12, 9, 64, 128)]


@pytest.mark.parametrize("n,h,w,ci,co", CONV_CASES)
def test_conv3x3_fwd(G, n, h, w, ci, co):
    x = rnd(n, h, w, ci, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(9 * ci) ** -0.5)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    stats = torch.zeros(2 * co, device="cuda")
    out = G.conv3x3_fwd(x, G.pack_conv3x3(wt), stats=stats)
    assert rel(out, ref) < 1e-2
    assert rel(stats[:co], ref.sum((0, 1, 2))) < 3e-3
    assert rel(stats[co:], (ref * ref).sum((0, 1, 2))) < 3e-3


@pytest.mark.parametrize("n,h,w,ci,co", CONV_CASES)
def test_conv3x3_dgrad(G, n, h, w, ci, co):
    dy = rnd(n, h, w, co, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(9 * co) ** -0.5)
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    out = G.conv3x3_dgrad(dy, G.pack_conv3x3(wt), ci)
    assert rel(out, ref) < 1e-2


@pytest.mark.parametrize("n,h,w,ci,co", CONV_CASES)
def test_conv3x3_wgrad(G, n, h, w, ci, co):
    x = rnd(n, h, w, ci, seed=1)
    dy = rnd(n, h, w, co, seed=2, scale=(n * h * w) ** -0.5)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    wref = torch.nn.grad.conv2d_weight(xr, (co, ci, 3, 3), dy.float().permute(0, 3, 1, 2), padding=1)
    out = G.unpack_conv3x3_grad(G.conv3x3_wgrad(dy, x), ci)
    assert rel(out, wref) < 3e-3
