"""GPU parity of every non-GEMM kernel and of the composed towers against torch fp32 / the CPU oracle on the same
seeded inputs (single op or single block, so bf16 storage noise is not amplified by depth).
Tolerances: bf16 outputs 1e-2 of the tensor's range; fp32 reductions 2e-3."""
import argparse
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def frob(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def rnd(*shape, seed=0, scale=1.0, dtype=bf16, shift=0.0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale + shift).to(dtype)


@pytest.fixture(scope="module")
def K():
    from tris_b200 import _lib, ops
    _lib.require_device()
    return ops


def nchw(x):
    return x.float().permute(0, 3, 1, 2)


def nhwc(x):
    return x.permute(0, 2, 3, 1)


def make_bn(K, c, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    u = lambda lo, hi: torch.rand(c, generator=g, device="cuda") * (hi - lo) + lo
    return K.BNState(u(0.5, 1.5), u(-0.3, 0.3), u(-0.1, 0.1), u(0.5, 1.5), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda"))


def stats_of(y, parts=148):
    """Partial-statistics rows [parts, 2C] as the GEMM epilogue writes them: the rows are split over the parts."""
    f = y.float().reshape(-1, y.shape[-1])
    rows = torch.zeros((parts, 2 * f.shape[1]), device=y.device)
    for i, chunk in enumerate(f.tensor_split(min(parts, 7))):
        rows[i] = torch.cat([chunk.sum(0), (chunk * chunk).sum(0)])
    return rows.reshape(-1).contiguous()


@pytest.mark.parametrize("c,pool,mode", [(64, 1, "plain"), (256, 2, "plain"), (512, 1, "residual"), (2048, 1, "dual"), (32, 1, "plain")])
def test_bn_apply_and_backward(K, c, pool, mode):
    n, h, w = 3, 12, 10
    y = rnd(n, h, w, c, seed=1, shift=0.3)
    bn = make_bn(K, c, 2)
    rm0, rv0 = bn.rm.clone(), bn.rv.clone()
    y1 = rnd(n, h, w, c, seed=3) if mode == "dual" else None
    bn1 = make_bn(K, c, 4) if mode == "dual" else None
    res = rnd(n, h, w, c, seed=5) if mode == "residual" else None
    out = K.bn_apply(y, stats_of(y), bn, True, relu=True, pool=pool, y1=y1, stats1=stats_of(y1) if y1 is not None else None,
                     bn1=bn1, residual=res)
    # torch reference
    yt = nchw(y).requires_grad_(True)
    gam, bet = bn.gamma.clone().requires_grad_(True), bn.beta.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    z = F.batch_norm(yt, rm, rv, gam, bet, True, 0.1, 1e-5)
    extra = []
    if mode == "dual":
        y1t = nchw(y1).requires_grad_(True)
        g1, b1 = bn1.gamma.clone().requires_grad_(True), bn1.beta.clone().requires_grad_(True)
        z = z + F.batch_norm(y1t, bn1.rm.clone(), bn1.rv.clone(), g1, b1, True, 0.1, 1e-5)
        extra = [y1t, g1, b1]
    if mode == "residual":
        z = z + nchw(res)
    ref = F.relu(z)
    if pool == 2:
        ref = F.avg_pool2d(ref, 2)
    assert rel(out, nhwc(ref)) < 1e-2
    assert rel(bn.rm, rm) < 1e-4 and rel(bn.rv, rv) < 1e-4
    # backward
    dout = rnd(*out.shape, seed=7)
    out_for_mask = out if mode != "plain" else None
    dy, dy1, gid = K.bn_bwd(dout, out_for_mask, y, bn, relu=True, pool=pool, y1=y1, bn1=bn1, want_g=(mode == "residual"))
    grads = torch.autograd.grad(ref, [yt, gam, bet] + extra, nchw(dout))
    assert rel(dy, nhwc(grads[0])) < 1.5e-2
    assert rel(bn.dgamma, grads[1]) < 1e-2 and rel(bn.dbeta, grads[2]) < 1e-2
    if mode == "dual":
        assert rel(dy1, nhwc(grads[3])) < 1.5e-2
        assert rel(bn1.dgamma, grads[4]) < 1e-2 and rel(bn1.dbeta, grads[5]) < 1e-2
    if mode == "residual":
        assert rel(gid, dout.float() * (out.float() > 0)) < 1e-2
    if mode != "plain" and pool == 1:
        # the sign-bit mask written by the forward kernel replaces re-reading `out`: identical results, bit for bit
        bits = torch.zeros((n * h * w, c // 8), device="cuda", dtype=torch.uint8)
        bn_b, bn1_b = make_bn(K, c, 2), (make_bn(K, c, 4) if mode == "dual" else None)
        out_b = K.bn_apply(y, stats_of(y), bn_b, True, relu=True, pool=pool, y1=y1, stats1=stats_of(y1) if y1 is not None else None,
                           bn1=bn1_b, residual=res, bits=bits)
        assert torch.equal(out_b, out)
        want = (out.float().reshape(n * h * w, c // 8, 8) > 0).to(torch.uint8)
        want = (want << torch.arange(8, device="cuda", dtype=torch.uint8)).sum(-1).to(torch.uint8)
        assert torch.equal(bits, want)
        dy_b, dy1_b, _ = K.bn_bwd(dout, None, y, bn_b, relu=True, pool=pool, y1=y1, bn1=bn1_b, bits=bits)
        assert torch.equal(dy_b, dy) and torch.equal(bn_b.dgamma, bn.dgamma) and torch.equal(bn_b.dbeta, bn.dbeta)
        if mode == "dual":
            assert torch.equal(dy1_b, dy1)


def test_bn_pooled_unpair_layout(K):
    """Pair-packed stem (images 2i, 2i+1 share a row as channel halves): the pooled apply kernel writes the un-paired
    [2n, h/2, w/2, c/2] layout itself and the backward kernels read the un-paired gradient -- bit-identical to the explicit
    permute copies they replace."""
    n, h, w, c = 3, 12, 8, 128
    y = rnd(n, h, w, c, seed=11, shift=0.3)
    bn_a, bn_b = make_bn(K, c, 12), make_bn(K, c, 12)
    paired = K.bn_apply(y, stats_of(y), bn_a, True, relu=True, pool=2, fold_half=64)
    unp = K.bn_apply(y, stats_of(y), bn_b, True, relu=True, pool=2, fold_half=64, unpair=True)
    assert unp.shape == (2 * n, h // 2, w // 2, c // 2)
    want = paired.view(n, h // 2, w // 2, 2, 64).permute(0, 3, 1, 2, 4).reshape(2 * n, h // 2, w // 2, 64)
    assert torch.equal(unp, want)
    dout = rnd(2 * n, h // 2, w // 2, 64, seed=13)
    dpaired = dout.reshape(n, 2, h // 2, w // 2, 64).permute(0, 2, 3, 1, 4).reshape(n, h // 2, w // 2, 128).contiguous()
    dy_a, _, _ = K.bn_bwd(dpaired, None, y, bn_a, pool=2, fold_half=64)
    dy_b, _, _ = K.bn_bwd(dout, None, y, bn_b, pool=2, fold_half=64, unpair=True)
    assert torch.equal(dy_a, dy_b) and torch.equal(bn_a.dgamma, bn_b.dgamma) and torch.equal(bn_a.dbeta, bn_b.dbeta)


def test_bn_eval_mode(K):
    y = rnd(2, 8, 8, 128, seed=1)
    bn = make_bn(K, 128, 2)
    out = K.bn_apply(y, None, bn, False, relu=True)
    ref = F.relu(F.batch_norm(nchw(y), bn.rm, bn.rv, bn.gamma, bn.beta, False, 0.1, 1e-5))
    assert rel(out, nhwc(ref)) < 1e-2


def test_avgpool(K):
    x = rnd(2, 8, 6, 64, seed=1)
    assert rel(K.avgpool2(x), nhwc(F.avg_pool2d(nchw(x), 2))) < 1e-2
    d = rnd(2, 4, 3, 64, seed=2)
    add = rnd(2, 8, 6, 64, seed=3)
    ref = F.interpolate(nchw(d), scale_factor=2, mode="nearest") * 0.25 + nchw(add)
    assert rel(K.avgpool2_bwd(d, add), nhwc(ref)) < 1e-2


@pytest.mark.parametrize("d", [512, 768])
def test_layernorm(K, d):
    rows = 333
    x = rnd(rows, d, seed=1, shift=0.2)
    g = torch.rand(d, device="cuda") + 0.5
    b = torch.randn(d, device="cuda") * 0.1
    y, mean, rstd = K.layernorm_fwd(x, g, b)
    xt = x.float().requires_grad_(True)
    gt, bt = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xt, (d,), gt, bt, 1e-5)
    assert rel(y, ref) < 1e-2
    dy = rnd(rows, d, seed=2)
    add = rnd(rows, d, seed=3)
    dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    dx = K.layernorm_bwd(dy, x, g, mean, rstd, add=add, dgamma=dg, dbeta=db)
    gr = torch.autograd.grad(ref, [xt, gt, bt], dy.float())
    assert rel(dx, gr[0] + add.float()) < 1e-2
    assert rel(dg, gr[1]) < 5e-3 and rel(db, gr[2]) < 5e-3
    # fp32 rows (the residual stream of the transformer towers): fp32 in -> bf16 out, bf16 in -> fp32 out, fp32 x in backward
    x32 = torch.randn(rows, d, device="cuda") + 0.2
    y2, mean2, rstd2 = K.layernorm_fwd(x32, g, b)
    ref2 = F.layer_norm(x32, (d,), g, b, 1e-5)
    assert y2.dtype == torch.bfloat16 and rel(y2, ref2) < 1e-2
    assert rel(mean2, x32.mean(-1)) < 1e-5
    y3, _, _ = K.layernorm_fwd(x, g, b, out_dtype=torch.float32)
    assert y3.dtype == torch.float32 and rel(y3, ref) < 1e-5
    x32t = x32.clone().requires_grad_(True)
    gr2 = torch.autograd.grad(F.layer_norm(x32t, (d,), g, b, 1e-5), x32t, dy.float())[0]
    dx2 = K.layernorm_bwd(dy, x32, g, mean2, rstd2, add=add)
    assert dx2.dtype == torch.bfloat16 and rel(dx2, gr2 + add.float()) < 1e-2


@pytest.mark.parametrize("n,l,heads,causal", [(5, 20, 8, True), (3, 50, 12, False), (2, 40, 8, True)])
def test_attention(K, n, l, heads, causal):
    d = heads * 64
    qkv = rnd(n * l, 3 * d, seed=1)
    out = K.attn_fwd(qkv, n, l, heads, causal)
    t = qkv.float().requires_grad_(True)
    q, k, v = [z.reshape(n, l, heads, 64).transpose(1, 2) for z in t.split(d, dim=-1)]
    s = q @ k.transpose(-1, -2) / 8.0
    if causal:
        s = s + torch.full((l, l), float("-inf"), device="cuda").triu_(1)
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n * l, d)
    assert rel(out, ref) < 1e-2
    do = rnd(n * l, d, seed=2)
    dqkv = K.attn_bwd(qkv, do, n, l, heads, causal)
    gr = torch.autograd.grad(ref, t, do.float())[0]
    assert rel(dqkv, gr) < 1.5e-2


def test_embedding_gather_scatter_colsum(K):
    n, l, d, vocab = 6, 20, 512, 1000
    E = torch.randn(vocab, d, device="cuda") * 0.02
    Pp = torch.randn(77, d, device="cuda") * 0.01
    ids = torch.randint(1, vocab - 1, (n, l), device="cuda", dtype=torch.int32)
    ids[:, 7] = vocab - 1
    x, eot = K.embed_fwd(ids, E, Pp)
    ref = E[ids.long()] + Pp[:l]
    assert rel(x, ref.reshape(n * l, d)) < 1e-2
    assert torch.equal(eot.long(), torch.arange(n, device="cuda") * l + ids.long().argmax(-1))
    x32, _ = K.embed_fwd(ids, E, Pp, out_dtype=torch.float32)
    assert torch.equal(x32, ref.reshape(n * l, d))
    pick = torch.tensor([3, 0, 77, 5], device="cuda", dtype=torch.int32)
    assert torch.equal(K.gather_rows(x32, pick), x32[pick.long()])
    dx = rnd(n * l, d, seed=1)
    dE, dP = torch.zeros_like(E), torch.zeros_like(Pp)
    K.embed_bwd(ids, dx, dE, dP)
    refE = torch.zeros_like(E).index_add_(0, ids.long().reshape(-1), dx.float())
    assert rel(dE, refE) < 1e-5
    assert rel(dP[:l], dx.float().reshape(n, l, d).sum(0)) < 1e-5
    g = K.gather_rows(x, eot)
    assert torch.equal(g, x[eot.long()])
    sc = K.scatter_rows(g, eot, n * l)
    assert torch.equal(sc[eot.long()], g) and sc.float().abs().sum() == g.float().abs().sum()
    cs = torch.zeros(d, device="cuda")
    K.colsum(dx, cs)
    assert rel(cs, dx.float().sum(0)) < 1e-4


def test_l2norm_instnorm_axpby(K):
    x = rnd(4800, 1024, seed=1)
    y, inv = K.l2norm_fwd(x)
    xt = x.float().requires_grad_(True)
    ref = xt / xt.norm(dim=-1, keepdim=True)
    assert rel(y, ref) < 1e-2
    dy = rnd(4800, 1024, seed=2)
    dx = K.l2norm_bwd(dy, y, inv)
    assert rel(dx, torch.autograd.grad(ref, xt, dy.float())[0]) < 1.5e-2
    # instance norm over 100 pixels per image, 3072 channels, relu
    b, p, c = 6, 100, 3072
    x = rnd(b * p, c, seed=3, shift=0.5)
    g = torch.rand(c, device="cuda") + 0.5
    be = torch.randn(c, device="cuda") * 0.3
    out, mean, invstd = K.instnorm_fwd(x, g, be, b, relu=True)
    xt = x.float().reshape(b, p, c).requires_grad_(True)
    gt, bt = g.clone().requires_grad_(True), be.clone().requires_grad_(True)
    mu = xt.mean(1, keepdim=True)
    var = xt.var(1, unbiased=False, keepdim=True)
    ref = F.relu((xt - mu) * torch.rsqrt(var + 1e-5) * gt + bt)
    assert rel(out, ref.reshape(b * p, c)) < 1e-2
    do = rnd(b * p, c, seed=4)
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    dx = K.instnorm_bwd(do, x, g, be, mean, invstd, dg, db, b, relu=True)
    gr = torch.autograd.grad(ref, [xt, gt, bt], do.float().reshape(b, p, c))
    assert rel(dx, gr[0].reshape(b * p, c)) < 1.5e-2
    assert rel(dg, gr[1]) < 1e-2 and rel(db, gr[2]) < 1e-2
    # residual mix: out = add + 0.1 * IN(x)
    add = rnd(b * p, c, seed=5)
    out2, _, _ = K.instnorm_fwd(x, g, be, b, relu=False, mix_scale=0.1, mix_add=add)
    ref2 = add.float() + 0.1 * ((xt - mu) * torch.rsqrt(var + 1e-5) * g + be).reshape(b * p, c)
    assert rel(out2, ref2) < 1e-2
    a, bb = rnd(1024, 64, seed=6), rnd(1024, 64, seed=7)
    refa = 0.5 * a.float() + 2.0 * bb.float()
    assert rel(K.axpby(a, bb.clone(), 0.5, 2.0), refa) < 1e-2


def test_stem_im2col_and_pack(K):
    img = torch.randn(2, 3, 64, 96, device="cuda")
    col = K.stem_im2col(img)
    w = torch.randn(32, 3, 3, 3, device="cuda")
    wp = torch.empty(32, 27, device="cuda", dtype=bf16)
    K.pack_conv(w, wp)
    got = col[:, :27].float() @ wp.float().t()
    ref = F.conv2d(img.to(bf16).float(), w.to(bf16).float(), stride=2, padding=1)
    assert rel(got.reshape(2, 32, 48, 32), nhwc(ref)) < 2e-3
    assert col[:, 27:].abs().max() == 0
    w3 = torch.randn(32, 32, 3, 3, device="cuda")
    p3 = torch.empty(64, 9 * 64, device="cuda", dtype=bf16)
    K.pack_conv(w3, p3, co_pad=64, ci_pad=64)
    refp = torch.zeros(64, 9, 64, device="cuda")
    refp[:32, :, :32] = w3.permute(0, 2, 3, 1).reshape(32, 9, 32)
    assert torch.equal(p3.float(), refp.reshape(64, 576).to(bf16).float())
    gp = torch.randn(64, 576, device="cuda")
    gw = torch.zeros(32, 32, 3, 3, device="cuda")
    K.unpack_conv_grad(gp, gw, ci_pad=64)
    assert torch.equal(gw, gp.reshape(64, 3, 3, 64)[:32, :, :, :32].permute(0, 3, 1, 2))


def test_fused_adamw_matches_torch():
    import ctypes as C
    from tris_b200 import _lib as L
    n, n0 = 4096 * 3, 4096
    g = torch.Generator(device="cuda").manual_seed(0)
    p0 = torch.randn(n, generator=g, device="cuda")
    pa = torch.nn.Parameter(p0[:n0].clone())
    pb = torch.nn.Parameter(p0[n0:].clone())
    opt = torch.optim.AdamW([{"params": [pa], "lr": 5e-6}, {"params": [pb], "lr": 5e-5}], lr=5e-5, weight_decay=0.01)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda x: (1 - x / 50) ** 0.9)
    p = p0.clone()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    shadow = torch.zeros(n, device="cuda", dtype=bf16)
    step = torch.zeros(1, device="cuda", dtype=torch.int32)
    for it in range(4):
        gr = torch.randn(n, generator=g, device="cuda")
        pa.grad, pb.grad = gr[:n0].clone(), gr[n0:].clone()
        opt.step(); sched.step()
        L.call("tris_adamw_step", C.c_void_p(p.data_ptr()), C.c_void_p((gr * 2).data_ptr()), C.c_void_p(m.data_ptr()),
               C.c_void_p(v.data_ptr()), C.c_void_p(shadow.data_ptr()), C.c_long(n), C.c_long(n0), C.c_void_p(step.data_ptr()),
               C.c_float(50.0), C.c_float(5e-6), C.c_float(5e-5), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8),
               C.c_float(0.01), C.c_float(0.5), C.c_float(0.9), launches=2)
    ref = torch.cat([pa.detach(), pb.detach()])
    assert rel(p, ref) < 1e-5 and step.item() == 4
    assert rel(shadow, ref) < 1e-2


# ------------------------------------------------------------------------------------------------ composed towers
def _args():
    return argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)


@pytest.fixture(scope="module")
def tris():
    import warnings
    warnings.simplefilter("ignore")
    from oracle import weights as W
    from tris_b200.model_stage1 import TRIS
    sd = W.make_tris_state_dict(0)
    m = TRIS(_args())
    m.load_state_dict(sd)
    m = m.cuda().train()
    eng = m.engine()
    eng.ensure_fresh(True)
    return m, eng, sd


def test_text_tower_fwd_bwd_vs_oracle(tris):
    from oracle import tris_oracle as O
    from oracle import weights as W
    m, eng, sd = tris
    _, ids, _ = W.synthetic_batch(6, 32, 20, 0, 5)
    hidden, rec, _ = eng.text.forward(ids.cuda(), save=True)
    keys = [k for k in sd if k.startswith("backbone.") and ("transformer" in k or k in (
        "backbone.text_projection", "backbone.positional_embedding", "backbone.token_embedding.weight",
        "backbone.ln_final.weight", "backbone.ln_final.bias"))]
    leaf = dict(sd)
    for k in keys:
        leaf[k] = sd[k].clone().requires_grad_(True)
    _, href = O.encode_text(leaf, ids, prefix="backbone.")
    assert rel(hidden, href) < 2e-2
    dh = torch.randn(6, 1024, generator=torch.Generator().manual_seed(1))
    eng.store.zero_grad()
    eng.text.backward(rec, dh.cuda().to(bf16))
    gr = torch.autograd.grad(href, [leaf[k] for k in keys], dh, allow_unused=True)
    worst = 0.0
    for k, g in zip(keys, gr):
        if g is None or g.abs().max() < 1e-7:
            continue
        e = rel(eng.store.g(k), g)
        worst = max(worst, e)
        assert e < 6e-2, (k, e)
    print("text tower worst grad rel err", worst)


def test_vit_tower_fwd_dgrad_vs_oracle():
    import warnings
    warnings.simplefilter("ignore")
    from oracle import tris_oracle as O
    from oracle import weights as W
    from tris_b200 import clip_model
    from tris_b200.engine import patchify
    aux_sd = W.make_vitb32_clip_state_dict(7, cos_bias=True)
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
    aux.load_state_dict(aux_sd)
    img = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    x = img.cuda().requires_grad_(True)
    feat = aux.encode_image(x)
    xr = img.clone().requires_grad_(True)
    ref = O.vit_tower(aux_sd, xr)
    assert rel(feat, ref) < 2e-2
    df = torch.randn(3, 512, generator=torch.Generator().manual_seed(4))
    feat.backward(df.cuda())
    ref.backward(df)
    assert rel(x.grad, xr.grad) < 6e-2
    _, ids, _ = W.synthetic_batch(5, 32, 20, 0, 6)
    seq, hid = aux.encode_text(ids.cuda())
    sref, href = O.encode_text(aux_sd, ids)
    assert rel(hid, href) < 2e-2 and rel(seq, sref) < 3e-2


@pytest.mark.parametrize("name,hw", [("layer1.0", 16), ("layer1.1", 16), ("layer2.0", 16), ("layer4.2", 6)])
def test_bottleneck_block_fwd_bwd_vs_oracle(tris, name, hw):
    from oracle import tris_oracle as O
    m, eng, sd = tris
    tower = eng.resnet
    blk = [b for b in tower.blocks if b.p.endswith(name + ".")][0]
    n = 4
    x = torch.relu(torch.randn(n, blk.cin, hw, hw, generator=torch.Generator().manual_seed(2)))
    xq = x.to(bf16).float()
    so = [0]
    tower.stats_buf.zero_()

    def stats(c):
        s = tower.stats_buf[so[0]: so[0] + 2 * c * 148]
        so[0] += 2 * c * 148
        return s

    sdb = {k: v.clone() for k, v in m.state_dict().items() if k.startswith(blk.p)}
    out, rec = tower._block_fwd(blk, nhwc(xq).contiguous().cuda().to(bf16), True, stats)
    keys = [k for k in sd if k.startswith(blk.p) and sd[k].is_floating_point() and "running" not in k]
    leaf = dict(sd)
    for k in keys:
        leaf[k] = sd[k].clone().requires_grad_(True)
    xr = xq.clone().requires_grad_(True)
    ref = O.bottleneck(xr, leaf, blk.p[:-1], blk.stride, True, {})
    assert rel(out, nhwc(ref)) < 2e-2
    dout = torch.randn(ref.shape, generator=torch.Generator().manual_seed(3))
    eng.store.zero_grad()
    dx, _ = tower._block_bwd(blk, rec, nhwc(dout).contiguous().cuda().to(bf16))
    gr = torch.autograd.grad(ref, [xr] + [leaf[k] for k in keys], dout)
    # Frobenius error: ReLU masks are recomputed from bf16-rounded pre-activations, so ~0.4 % of the mask bits differ
    # from the fp32 reference and each flip is an O(1) outlier in max-norm; a wrong formula shows up as >= 30 %.
    errs = {"dx": frob(dx, nhwc(gr[0]))}
    for k, g in zip(keys, gr[1:]):
        errs[k[len(blk.p):]] = frob(eng.store.g(k), g)
    print(name, {k: round(v, 4) for k, v in errs.items()})
    for k, v in errs.items():
        assert v < 0.12, (k, v)
    m.load_state_dict(sdb, strict=False)


def test_pair_packed_stem_matches_padded_stem(tris):
    """Even batches run the stem with two images per 64-channel row (block-diagonal weights, folded BatchNorm sums);
    it must agree with the zero-padded form used for odd batches: stem output, stem parameter gradients, running stats.
    (Stem in isolation: through the whole random-init tower at a tiny batch the comparison would be chaotic.)"""
    m, eng, sd = tris
    rn = eng.resnet
    img = torch.randn(4, 3, 128, 128, generator=torch.Generator().manual_seed(11)).cuda()
    dx = (torch.randn(4, 32, 32, 64, generator=torch.Generator().manual_seed(12)) * 0.1).cuda().to(bf16)
    p = "backbone.visual."
    keys = [p + k for k in ("conv1.weight", "bn1.weight", "bn1.bias", "conv2.weight", "bn2.weight", "bn2.bias", "conv3.weight",
                            "bn3.weight", "bn3.bias")]
    bufs = dict(m.named_buffers())
    res = {}
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    for pair in (False, True):
        m.load_state_dict(sd0)
        eng.ensure_fresh(True)
        eng.store.zero_grad()
        rn.stats_buf.zero_()
        off = [0]

        def stats(c):
            s_ = rn.stats_buf[off[0]: off[0] + 2 * c * 148]
            off[0] += 2 * c * 148
            return s_
        x, rec = (rn._stem_fwd_pair if pair else rn._stem_fwd_padded)(img, True, stats)
        rn._stem_bwd((pair,) + rec, dx.clone())
        torch.cuda.synchronize()
        res[pair] = (x.float().clone(), {k: eng.store.g(k).clone() for k in keys},
                     {k: bufs[k].clone() for k in (p + "bn1.running_mean", p + "bn2.running_var", p + "bn3.running_var")})
    m.load_state_dict(sd0)
    e_x = frob(res[True][0], res[False][0])
    errs = {k: frob(res[True][1][k], res[False][1][k]) for k in keys}
    print("stem out", e_x, errs)
    assert e_x < 1e-2
    for k, e in errs.items():
        assert e < 3e-2, (k, e)
    for k, v in res[True][2].items():
        assert rel(v, res[False][2][k]) < 2e-3, k


def test_fused_bn_backward_matches_unfused_and_is_bit_reproducible(tris):
    """RN50 tower backward with the BatchNorm-backward reductions fused into the GEMM epilogues (default) against the
    stand-alone reduction kernels on the same tape; and two runs of the default path give identical gradient bits
    (fixed-order two-stage reductions everywhere: no floating-point atomics in the image tower)."""
    import tris_b200.resnet as R
    m, eng, sd = tris
    rn = eng.resnet
    img = torch.randn(4, 3, 128, 128, generator=torch.Generator().manual_seed(21)).cuda()
    dc4 = (torch.randn(4, 4, 4, 2048, generator=torch.Generator().manual_seed(22)) * 0.05).cuda().to(bf16)
    keys = [k for k in eng.store.trainable if k.startswith("backbone.visual.") and "attnpool" not in k]
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    res = []
    saved = R.FUSE_BN_BWD
    try:
        for fuse in (True, True, False, False):
            R.FUSE_BN_BWD = fuse
            m.load_state_dict(sd0)
            eng.ensure_fresh(True)
            eng.store.zero_grad()
            c4, tape = rn.forward(img, train=True)
            rn.backward(tape, dc4.clone())
            torch.cuda.synchronize()
            res.append((c4.clone(), {k: eng.store.g(k).clone() for k in keys}))
    finally:
        R.FUSE_BN_BWD = saved
        m.load_state_dict(sd0)
    assert torch.equal(res[0][0], res[1][0])
    for k in keys:
        assert torch.equal(res[0][1][k], res[1][1][k]), f"{k}: gradient differs between two identical runs"
    for k in keys:
        assert torch.equal(res[2][1][k], res[3][1][k]), f"{k}: gradient differs between two identical (unfused) runs"
    worst = max((frob(res[0][1][k], res[2][1][k]), k) for k in keys)
    print("fused vs unfused BatchNorm backward, worst gradient difference:", worst)
    assert worst[0] < 3e-2, worst
