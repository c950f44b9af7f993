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
    # fp32 residual stream of the transformer towers: fp32 residual in, fp32 out, nothing rounded to bf16 in between
    res32 = torch.randn(m, n, device="cuda") * 3
    out = G.linear_fwd(x, w, bias, residual=res32)
    assert out.dtype == torch.float32 and rel(out, pre + res32) < 2e-4
    out = G.linear_fwd(x[:77], w[:520], bias[:520], residual=res32[:77, :520].contiguous())     # ragged tile edges
    assert rel(out, pre[:77, :520] + res32[:77, :520]) < 2e-4
    stats = torch.full((148 * 2 * n,), 7.0, device="cuda")     # partial rows: every row is written by the kernel
    out = G.linear_fwd(x, w, bias, act=L.ACT_RELU)
    assert rel(out, F.relu(pre)) < 1e-2
    out = G.linear_fwd(x, w, bias, stats=stats)        # column statistics of the stored (bf16) output
    stats = stats.view(148, 2 * n).sum(0)
    assert rel(stats[:n], pre.sum(0)) < 3e-3
    assert rel(stats[n:], (pre * pre).sum(0)) < 3e-3
    assert rel(stats[:n], out.float().sum(0)) < 1e-4


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


# (the first, and the last four, tile into 16 x 8 patches: they run the halo-re-use form of the kernel)
CONV_CASES = [(2, 80, 80, 64, 64), (3, 40, 40, 128, 128), (2, 20, 20, 256, 256), (3, 10, 10, 512, 512),
              (1, 20, 20, 512, 512), (2, 12, 9, 64, 128), (2, 32, 16, 64, 64), (3, 16, 8, 128, 64), (2, 48, 24, 64, 128),
              (24, 160, 160, 64, 128)]


@pytest.mark.parametrize("n,h,w,ci,co", CONV_CASES)
def test_conv3x3_fwd(G, n, h, w, ci, co):
    x = rnd(n, h, w, ci, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(9 * ci) ** -0.5)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    stats = torch.full((148 * 2 * co,), 7.0, device="cuda")
    out = G.conv3x3_fwd(x, G.pack_conv3x3(wt), stats=stats)
    stats = stats.view(148, 2 * co).sum(0)
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


def test_strided_views_and_batched(G):
    """Q/K/V thirds of a fused projection as strided operands, and per-image batched products with M, K < tile."""
    from tris_b200 import _lib as L
    Bn, P, T, C = 5, 100, 48, 1024
    qkv = rnd(Bn * P, 3 * C, seed=1)          # [B*P, 3C]
    t3 = rnd(T, 3 * C, seed=2)                # [T, 3C]
    # S[b*P+p, t] = Qv . Kt  (plain GEMM on strided views, fp32 out, N = 48)
    S = torch.empty(Bn * P, T, device="cuda", dtype=torch.float32)
    G.gemm_ex(qkv[:, :C], t3[:, C:2 * C], S, Bn * P, T, C, lda=3 * C, ldb=3 * C)
    ref = qkv[:, :C].float() @ t3[:, C:2 * C].float().t()
    assert rel(S, ref) < 2e-3
    # new_vis[b*P+p, c] = Av[b*P+p, :T] @ Vt[T, c]   (K = 48 < 64: zero-filled; B read MN-major from a strided view)
    av = rnd(Bn * P, T, seed=3).abs()
    nv = torch.empty(Bn * P, C, device="cuda", dtype=torch.bfloat16)
    G.gemm_ex(av, t3[:, 2 * C:], nv, Bn * P, C, T, b_mode=L.OP_MN2D, lda=T, ldb=3 * C)
    assert rel(nv, av.float() @ t3[:, 2 * C:].float()) < 1e-2
    # batched: new_lan[b, t, c] = At[b, t, :P] @ Vv[b, :P, c]   (M = 48, K = 100 per image)
    at = rnd(Bn, T, 104, seed=4).abs()        # row stride 104 (16-byte aligned), only [:P] valid
    nl = torch.empty(Bn, T, C, device="cuda", dtype=torch.bfloat16)
    vv = qkv[:, 2 * C:]
    G.gemm_ex(at, vv, nl, T, C, P, b_mode=L.OP_MN2D, lda=104, ldb=3 * C, batch=Bn, a_bs=T * 104, b_bs=P * 3 * C, d_bs=T * C)
    refl = torch.bmm(at[:, :, :P].float(), vv.float().reshape(Bn, P, C))
    assert rel(nl, refl) < 1e-2
    # batched score: sc[b, p, t] = v'[b, p, :] . l'[b, t, :]   (fp32 out, both K-major, per-image B)
    vp, lp = rnd(Bn, P, C, seed=5, scale=0.05), rnd(Bn, T, C, seed=6, scale=0.05)
    sc = torch.empty(Bn, P, T, device="cuda", dtype=torch.float32)
    G.gemm_ex(vp, lp, sc, P, T, C, batch=Bn, a_bs=P * C, b_bs=T * C, d_bs=P * T)
    assert rel(sc, torch.bmm(vp.float(), lp.float().transpose(1, 2))) < 2e-3
    # batched, contraction over rows (dVv-like): out[b, p, c] = At[b, :, p]^T-form: A MN-major [K=T, M=P]
    dvv = torch.empty(Bn, P, C, device="cuda", dtype=torch.bfloat16)
    G.gemm_ex(at, lp, dvv, P, C, T, a_mode=L.OP_MN2D, b_mode=L.OP_MN2D, lda=104, ldb=C, batch=Bn, a_bs=T * 104, b_bs=T * C,
              d_bs=P * C)
    assert rel(dvv, torch.bmm(at[:, :, :P].float().transpose(1, 2), lp.float())) < 1e-2


# ------------------------------------------------------------------------------------------------ round 2: deterministic reductions
def _bn_vectors(c, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mean = torch.randn(c, generator=g, device="cuda") * 0.2
    sc = torch.rand(c, generator=g, device="cuda") + 0.5
    sh = torch.randn(c, generator=g, device="cuda") * 0.3
    return mean, sc, sh


def _check_parts(parts, c, g, y, mean, tol=3e-3):
    rows = parts[: 148 * 2 * c].view(148, 2 * c).sum(0)
    gf, yf = g.float().reshape(-1, c), y.float().reshape(-1, c)
    assert rel(rows[:c], gf.sum(0)) < tol
    assert rel(rows[c:], (gf * (yf - mean)).sum(0)) < tol


@pytest.mark.parametrize("m,n,k,mask", [(4800, 512, 256, True), (19200, 256, 1024, True), (1000, 136, 64, True), (4800, 1024, 256, False)])
def test_dgrad_with_fused_bn_backward_sums(G, m, n, k, mask):
    """linear_dgrad with the BatchNorm-backward epilogue: stored g = (dy W [+ r] [* relu'(x)]) * [y*sc+sh > 0] and the
    partial rows (sum g, sum g (y - mean)) of the stored g."""
    from tris_b200 import _lib as L
    dy, w = rnd(m, n, seed=1), rnd(n, k, seed=2, scale=n ** -0.5)
    y = rnd(m, k, seed=3)
    mean, sc, sh = _bn_vectors(k, 4)
    parts = torch.full((148 * 2 * k + 2 * k,), 3.0, device="cuda")
    ref = dy.float() @ w.float()
    if mask:
        out = G.linear_dgrad(dy, w, bwd_stats=(parts, y, mean, sc, sh))
        ref = ref * ((y.float() * sc + sh) > 0)
    else:       # residual + ReLU derivative of a shared activation (bn3 of the block in front)
        r, x = rnd(m, k, seed=5), rnd(m, k, seed=6)
        out = G.linear_dgrad(dy, w, residual=r, dact_src=x, act=L.ACT_RELU, bwd_stats=(parts, y, mean, None, None))
        ref = (ref + r.float()) * (x.float() > 0)
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-2
    _check_parts(parts, k, out, y, mean)


@pytest.mark.parametrize("n,h,w,ci,co", [(3, 20, 20, 64, 64), (2, 40, 40, 128, 128), (5, 10, 10, 256, 256), (3, 12, 10, 64, 128)])
def test_conv_dgrad_with_fused_bn_backward_sums(G, n, h, w, ci, co):
    dy = rnd(n, h, w, co, seed=1)
    wt = rnd(co, ci, 3, 3, seed=2, scale=(9 * co) ** -0.5)
    y = rnd(n, h, w, ci, seed=3)
    mean, sc, sh = _bn_vectors(ci, 4)
    parts = torch.full((148 * 2 * ci + 2 * ci,), 3.0, device="cuda")
    ref = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), wt.float(), padding=1).permute(0, 2, 3, 1)
    ref = ref * ((y.float() * sc + sh) > 0)
    out = G.conv3x3_dgrad(dy, G.pack_conv3x3(wt), ci, bwd_stats=(parts, y, mean, sc, sh))
    torch.cuda.synchronize()
    assert rel(out, ref) < 1e-2
    _check_parts(parts, ci, out, y, mean)


def test_weight_gradients_are_bit_reproducible(G):
    """Split-K partials go through a workspace and are added in split order: two runs give identical bits (the round-1
    kernels used arrival-order TMA reduce-adds)."""
    dy, x = rnd(76800, 128, seed=1, scale=0.01), rnd(76800, 512, seed=2)
    a = G.linear_wgrad(dy, x).clone()
    for _ in range(3):
        assert torch.equal(a, G.linear_wgrad(dy, x))
    acc = torch.ones(128, 512, device="cuda")
    G.linear_wgrad(dy, x, out=acc, accumulate=True)
    assert rel(acc - 1.0, dy.float().t() @ x.float()) < 3e-3
    xc, dyc = rnd(48, 40, 40, 128, seed=3), rnd(48, 40, 40, 128, seed=4, scale=0.01)
    b = G.conv3x3_wgrad(dyc, xc).clone()
    for _ in range(3):
        assert torch.equal(b, G.conv3x3_wgrad(dyc, xc))
    st = torch.zeros(148 * 2 * 512, device="cuda")
    w = rnd(512, 128, seed=5)
    o1 = G.linear_fwd(dy, w, stats=st).clone()
    s1 = st.clone()
    G.linear_fwd(dy, w, stats=st)
    assert torch.equal(s1, st) and torch.equal(o1, G.linear_fwd(dy, w, stats=st))


def test_dgrad_with_bit_masked_residual(G):
    """dx = dy W + r * bit: the residual join of a bottleneck in backward without materialising the masked gradient."""
    m, n, k = 4800, 256, 1024
    dy, w, r = rnd(m, n, seed=1), rnd(n, k, seed=2, scale=n ** -0.5), rnd(m, k, seed=3)
    keep = torch.rand(m, k, device="cuda") > 0.4
    bits = (keep.reshape(m, k // 8, 8).to(torch.uint8) << torch.arange(8, device="cuda", dtype=torch.uint8)).sum(-1).to(torch.uint8)
    out = G.linear_dgrad(dy, w, residual=r, res_bits=bits.contiguous())
    assert rel(out, dy.float() @ w.float() + r.float() * keep) < 1e-2


@pytest.mark.parametrize("B,P,k,n,relu,mix", [(48, 100, 1024, 3072, True, 1.0), (48, 100, 1024, 1024, False, 0.1), (5, 49, 256, 192, True, 1.0),
                                               (3, 128, 128, 64, False, 0.5)])
def test_linear_instancenorm_fused(G, B, P, k, n, relu, mix):
    """1x1 conv + InstanceNorm2d(affine) [+ ReLU] [* mix + add] in one batched GEMM with a tile-local epilogue
    (v_proj / v_output of model/attn.py:75-105) against torch on the same bf16-rounded operands."""
    x, w = rnd(B * P, k, seed=1), rnd(n, k, seed=2, scale=k ** -0.5)
    g = torch.Generator(device="cuda").manual_seed(3)
    gamma = torch.rand(n, generator=g, device="cuda") + 0.5
    beta = torch.randn(n, generator=g, device="cuda") * 0.2
    add = rnd(B * P, n, seed=4) if not relu else None
    y, yn, mean, invstd = G.linear_in_fwd(x, w, B, gamma, beta, relu, mix_scale=mix, mix_add=add)
    torch.cuda.synchronize()
    ref = (x.float() @ w.float().t())
    assert rel(y, ref) < 1e-2
    yb = y.float().reshape(B, P, n)                     # statistics of the stored (bf16) values, like the unfused kernels
    mu = yb.mean(1)
    var = yb.var(1, unbiased=False)
    assert rel(mean, mu) < 2e-3 or (mean - mu).abs().max() < 2e-3
    assert rel(invstd, (var + 1e-5).rsqrt()) < 2e-3
    o = (yb - mu[:, None]) * (var + 1e-5).rsqrt()[:, None] * gamma + beta
    if relu:
        o = torch.relu(o)
    o = o * mix
    if add is not None:
        o = o + add.float().reshape(B, P, n)
    assert rel(yn, o.reshape(B * P, n)) < 1e-2
