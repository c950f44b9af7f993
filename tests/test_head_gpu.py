"""GPU parity of the Stage-1 head kernels (head.cu) and of the composed head (K6 + K7 + K8, forward and hand-written
backward) against the oracle restatement of model/model_stage1.py:61-119, model/attn.py:111-136 and
train_stage1.py:327-364 on the same seeded inputs.  fp32 results: 2e-3; bf16 outputs: 1e-2 of range; composed
gradients: 6e-2 / 8e-2 Frobenius (bf16 operands in ~25 chained GEMMs; tools/head_error_budget.py)."""
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


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.fixture(scope="module")
def K():
    from tris_b200 import _lib, ops
    _lib.require_device()
    return ops


@pytest.mark.parametrize("B,P,T", [(3, 100, 3), (48, 100, 48), (1, 100, 1), (5, 49, 5)])
def test_xattn_softmax(K, B, P, T):
    Tp = (T + 7) // 8 * 8
    S1, S2T = rnd(B, P, Tp, seed=1, scale=20), rnd(B, P, Tp, seed=2, scale=20)
    scale = 1 / 32
    PA, PAc, PTt = K.xattn_softmax_fwd(S1, S2T, T, scale)
    ra = torch.softmax(S1[:, :, :T] * scale, dim=2)
    rt = torch.softmax(S2T[:, :, :T] * scale, dim=1)
    assert rel(PA[:, :, :T], ra) < 1e-2 and rel(PTt[:, :, :T], rt) < 1e-2
    assert rel(PAc[:, :, :T], ra - ra.mean(1, keepdim=True)) < 1e-2
    assert PA[:, :, T:].abs().max().item() == 0 if Tp > T else True
    dPA, dPTt = rnd(B, P, Tp, seed=3), rnd(B, P, Tp, seed=4)
    dS1, dS2T = K.xattn_softmax_bwd(PA, dPA, PTt, dPTt, T, scale)
    pa, pt = PA[:, :, :T].float(), PTt[:, :, :T].float()
    dpa = dPA[:, :, :T] - dPA[:, :, :T].mean(1, keepdim=True)      # the kernel takes the gradient w.r.t. the centred PAc
    r1 = scale * pa * (dpa - (dpa * pa).sum(2, keepdim=True))
    r2 = scale * pt * (dPTt[:, :, :T] - (dPTt[:, :, :T] * pt).sum(1, keepdim=True))
    assert rel(dS1[:, :, :T], r1) < 1e-2 and rel(dS2T[:, :, :T], r2) < 1e-2


@pytest.mark.parametrize("B,P", [(3, 100), (48, 100), (6, 49)])
def test_response_head_fwd_bwd(K, B, P):
    from oracle import tris_oracle as O
    T, Tp = B, (B + 7) // 8 * 8
    h = int(math.isqrt(P))
    R = rnd(B, P, Tp, seed=5, scale=0.3)
    ls = torch.tensor(math.log(1 / 0.07), device="cuda")
    cls, fg, maps, mbar, am, es = K.head_fwd(R, ls, T, 3.0, 0.01, True)
    score = (R[:, :, :T] * ls.exp()).detach().requires_grad_(True)
    o = O.tris_head(score.cpu(), (h, h), (h * 32, h * 32), True)
    assert rel(cls, o["cls_out"]) < 2e-3 and rel(fg, o["cls_fg"]) < 2e-3
    assert rel(maps.view(B, 1, h, h), o["maps10"]) < 1e-5
    assert abs(es.item() - 1 / 0.07) < 1e-3
    # backward: gradients of sum(dcls*cls_out) + sum(dfg*cls_fg) + sum(dmaps*maps)
    dcls, dfg, dmaps = rnd(B, T, seed=6), rnd(B, seed=7), rnd(B, P, seed=8)
    sc = score.detach().cpu().requires_grad_(True)
    o = O.tris_head(sc, (h, h), (h * 32, h * 32), True)
    obj = (o["cls_out"] * dcls.cpu()).sum() + (o["cls_fg"] * dfg.cpu()).sum() + (o["maps10"].reshape(B, P) * dmaps.cpu()).sum()
    (gs,) = torch.autograd.grad(obj, sc)
    dls = torch.zeros((B,), device="cuda")                    # per-image partials (added in order by the caller)
    D = K.head_bwd(R, ls, dcls, dfg, dmaps, mbar, am, dls, T, 3.0, 0.01)
    dls = dls.sum()
    assert rel(D[:, :, :T].float() / ls.exp(), gs) < 1e-2      # D = dL/dR = e^s dL/dscore (bf16)
    assert abs(dls.item() - (gs * sc.detach()).sum().item()) < 2e-2 * (gs * sc.detach()).abs().sum().item()


@pytest.mark.parametrize("B,h,H", [(3, 10, 320), (2, 7, 224), (1, 12, 384)])
def test_upsample_fwd_bwd(K, B, h, H):
    maps = rnd(B, h * h, seed=9, scale=2.0)
    relu, sig = K.upsample_fwd(maps, h, h, H, H)
    m = maps.view(B, 1, h, h).clone().requires_grad_(True)
    seg = F.interpolate(m, size=(H, H), mode="bilinear", align_corners=False)
    assert rel(relu, F.relu(seg)) < 1e-5 and rel(sig, torch.sigmoid(seg)) < 1e-5
    drelu, dsig = rnd(B, 1, H, H, seed=10), rnd(B, 1, H, H, seed=11)
    (g,) = torch.autograd.grad((F.relu(seg) * drelu).sum() + (torch.sigmoid(seg) * dsig).sum(), m)
    got = K.upsample_bwd(drelu, dsig, sig, h, h)
    assert rel(got, g.reshape(B, -1)) < 1e-4
    got = K.upsample_bwd(None, dsig, sig, h, h)
    (g,) = torch.autograd.grad((torch.sigmoid(F.interpolate(m, size=(H, H), mode="bilinear", align_corners=False)) * dsig).sum(), m)
    assert rel(got, g.reshape(B, -1)) < 1e-4


@pytest.mark.parametrize("B,S", [(3, 320), (2, 224), (2, 384)])
def test_mask_resize_fwd_bwd(K, B, S):
    from oracle import tris_oracle as O
    sig = torch.sigmoid(rnd(B, 1, S, S, seed=12)).requires_grad_(True)
    img = rnd(B, 3, S, S, seed=13)
    patches, fg = K.mask_resize_fwd(sig.detach(), img, 224, 32, want_fg=True)
    ref = O.mask_and_resize(sig, img)
    assert rel(fg, ref) < 2e-4      # fp32 interpolation-weight rounding
    pref = ref.reshape(B, 3, 7, 32, 7, 32).permute(0, 2, 4, 1, 3, 5).reshape(B * 49, 3072)
    assert rel(patches, pref) < 1e-2
    # plain patchify
    p2, _ = K.mask_resize_fwd(None, ref.detach().contiguous(), 224, 32)
    assert rel(p2, pref) < 1e-2
    dp = rnd(B * 49, 3072, seed=14, dtype=bf16)
    (g,) = torch.autograd.grad((pref * dp.float()).sum(), sig)
    got = K.mask_resize_bwd(dp, img, 224, 32)
    assert rel(got, g) < 3e-4


@pytest.mark.parametrize("B,k", [(3, 3), (48, 3), (5, 0)])
def test_stage1_loss_fwd_bwd(K, B, k):
    f = rnd(B, 512, seed=15, dtype=bf16)
    g = (rnd(B * (1 + k), 512, seed=16) * 0.5 + torch.cat([f.float(), f.float().repeat_interleave(k, 0) * 0.3])).to(bf16)
    cls = rnd(B, B, seed=17, scale=2.0)
    w = (1.0, 5.0, 2.0)
    out = K.stage1_loss_fwd(f, g, cls, k, w)
    ff = f.float().requires_grad_(True)
    cc = cls.clone().requires_grad_(True)
    fn = ff / ff.norm(dim=-1, keepdim=True)
    gn = g.float() / g.float().norm(dim=-1, keepdim=True)
    cos = (fn * gn[:B]).sum(-1)
    l1 = -torch.log(cos.clamp(0.0001, 0.9999)).mean()
    l5 = (-torch.log(1 - torch.einsum("bc,bkc->bk", fn, gn[B:].reshape(B, k, -1)))).mean(1).sum() / B if k else torch.zeros((), device="cuda")
    l4 = F.multilabel_soft_margin_loss(cc, torch.eye(B, device="cuda"))
    loss = w[0] * l1 + w[1] * l4 + w[2] * l5
    ref = torch.stack([loss, l1, l4, l5])
    assert rel(out, ref) < 2e-4, (out, ref)
    dout = torch.tensor([1.0, 0.3, -0.2, 0.5], device="cuda")
    gf, gc = torch.autograd.grad((ref * dout).sum(), [ff, cc])
    df, dcls = K.stage1_loss_bwd(f, g, cls, dout, k, w)
    assert rel(df, gf) < 1e-2 and rel(dcls, gc) < 1e-4


# ------------------------------------------------------------------------------------------------ composed head
def _args(attn_multi=0.1):
    return argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=attn_multi, FOCAL_P=3, FOCAL_LAMBDA=0.01)


@pytest.fixture(scope="module")
def tris():
    import warnings
    warnings.simplefilter("ignore")
    from oracle import weights as W
    from tris_b200.model_stage1 import TRIS
    sd = W.make_tris_state_dict(0)
    g = torch.Generator().manual_seed(3)
    for k in sd:   # give the InstanceNorm affines / biases non-trivial values so every gradient path is exercised
        if k.startswith("attn_fusion.") and (k.endswith(".1.weight") or k.endswith(".1.bias")):
            sd[k] = sd[k] + 0.2 * torch.randn(sd[k].shape, generator=g)
    m = TRIS(_args())
    m.load_state_dict(sd)
    m = m.cuda().train()
    eng = m.engine()
    eng.ensure_fresh(True)
    return m, eng, sd


HEAD_KEYS = ["vis_project.weight", "vis_project.bias", "lan_project.weight", "lan_project.bias", "logit_scale"]


@pytest.mark.parametrize("B", [4, 48])
def test_head_fwd_bwd_vs_oracle(tris, B):
    from oracle import tris_oracle as O
    m, eng, sd = tris
    h = 10
    c4 = (rnd(B, h, h, 2048, seed=20).abs() * 0.5).to(bf16)                 # post-ReLU-like features, NHWC
    hidden = rnd(B, 1024, seed=21, scale=0.3, dtype=bf16)
    keys = HEAD_KEYS + [k for k in sd if k.startswith("attn_fusion.")]
    leaf = {k: v.cuda() for k, v in sd.items() if k in keys}
    for k in keys:
        leaf[k] = leaf[k].to(bf16).float().requires_grad_(True) if leaf[k].dim() > 1 else leaf[k].requires_grad_(True)
    c4r = c4.float().permute(0, 3, 1, 2).requires_grad_(True)
    hr = hidden.float().requires_grad_(True)
    score = O.tris_score(leaf, c4r, hr)
    o = O.tris_head(score, (h, h), (320, 320), True)
    eng.fwd_id += 1
    eng.store.zero_grad()
    c4g, hg = c4.clone().requires_grad_(True), hidden.clone().requires_grad_(True)
    cls, fg, relu, sig, es = eng.head.forward(c4g, hg, (320, 320), True)
    print("fwd rel: cls", rel(cls, o["cls_out"]), "fg", rel(fg, o["cls_fg"]), "relu", rel(relu, o["relu"]), "sig", rel(sig, o["sig"]))
    assert rel(cls, o["cls_out"]) < 2e-2 and rel(fg, o["cls_fg"]) < 2e-2
    assert rel(relu, o["relu"]) < 2e-2 and rel(sig, o["sig"]) < 2e-2
    # Backward objective: cls_fg (softmax means) + sigmoid maps -- both smooth.  cls_out contains max-over-pixels, whose
    # arg-max flips between two near-equal pixels under the 1e-2 forward noise (a discontinuity of the reference function
    # itself); its backward is pinned with identical inputs in test_response_head_fwd_bwd.  At B=4 (16 maxima, none
    # near a tie for this seed) the full objective is used as well.
    dcls, dfg, dsig = rnd(B, B, seed=22), rnd(B, seed=26), rnd(B, 1, 320, 320, seed=23, scale=0.01)
    if B > 4:
        dcls = None
    obj = (o["cls_fg"] * dfg).sum() + (o["sig"] * dsig).sum() + ((o["cls_out"] * dcls).sum() if dcls is not None else 0)
    gr = torch.autograd.grad(obj, [c4r, hr] + [leaf[k] for k in keys], allow_unused=True)
    if dcls is not None:
        torch.autograd.backward([cls, fg, sig], [dcls, dfg, dsig])
    else:
        torch.autograd.backward([fg, sig], [dfg, dsig])
    e_c4 = frob(c4g.grad.float().permute(0, 3, 1, 2), gr[0])
    e_h = frob(hg.grad, gr[1])
    print("dc4", e_c4, "dhidden", e_h)
    assert e_c4 < 6e-2 and e_h < 6e-2
    worst = 0.0
    for k, g in zip(keys, gr[2:]):
        got = eng.store.g(k)
        if g is None or g.abs().max() < 1e-6:
            continue
        e = frob(got.reshape(g.shape), g)
        worst = max(worst, e)
        # conv / linear biases that feed an InstanceNorm have an analytically zero gradient (noise only)
        if k.endswith(".0.bias") and ("v_proj" in k or "v_output" in k):
            continue
        assert e < 8e-2, (k, e)
    print("head worst param-grad frobenius rel err", worst)


def test_head_eval_single_sentence(tris):
    """Eval path of validate.py (batch 1, T = 1): the v_output InstanceNorm is degenerate."""
    from oracle import tris_oracle as O
    m, eng, sd = tris
    c4 = (rnd(1, 10, 10, 2048, seed=24).abs() * 0.5).to(bf16)
    hidden = rnd(1, 1024, seed=25, scale=0.3, dtype=bf16)
    sdc = {k: (v.cuda().to(bf16).float() if v.dim() > 1 else v.cuda()) for k, v in sd.items()
           if k in HEAD_KEYS or k.startswith("attn_fusion.")}
    score = O.tris_score(sdc, c4.float().permute(0, 3, 1, 2), hidden.float())
    ref = O.tris_head(score, (10, 10), (320, 320), False)["relu"]
    (out,) = eng.head.forward(c4, hidden, (320, 320), False)
    assert out.shape == (1, 1, 320, 320)
    assert rel(out, ref) < 2e-2


def test_resize_bilinear_ac(K):
    x = rnd(2, 1, 320, 320, seed=30)
    got = K.resize_bilinear_ac(x, 480, 640)
    ref = F.interpolate(x, (480, 640), mode="bilinear", align_corners=True)
    assert rel(got, ref) < 2e-4


def test_inference_drivers_synthetic(tris):
    """validate / validate_same_sentence (PRMS) / cached-feature path agree with the plain eval forward."""
    import validate as V
    from tris_b200 import clip_model
    m, eng, sd = tris
    args = _args()
    args.size, args.val_refs, args.save_cam, args.cam_save_dir, args.name_save_dir, args.dataset = 320, 3, False, None, None, "refcoco"
    m.eval()
    idx, img, word_ids, target = next(V.synthetic_refs(args, 1))
    with torch.no_grad():
        full = m(img.cuda(), word_ids[:, :, 0].cuda().contiguous())
        c4 = m.image_features(img.cuda())
        cached = m.respond(c4, word_ids[:, :, 0].cuda().contiguous(), (320, 320))
    assert torch.equal(full, cached)
    miou, hit = V.validate(args, V.synthetic_refs(args, 3), m)
    assert 0.0 <= miou <= 1.0 and 0.0 <= hit <= 1.0
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
    miou2 = V.validate_same_sentence(args, V.synthetic_refs(args, 3), m, aux)
    assert 0.0 <= miou2 <= 1.0
    m.train()
