"""Drop-in proof: the reference's OWN, unmodified drivers -- train_one_epoch
(train_stage1.py:286-411), validate (validate.py:131-249) and validate_same_sentence (validate.py:253-387) -- are run on
the GPU box against tris_b200.TRIS / tris_b200.clip_model and against the reference modules on the same weights and
synthetic loaders, and their outputs are compared.  Needs the shipped copy of the reference (baseline/_ref, made by
baseline/install_reference.py); skipped when it is absent.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from baseline import ref_loader  # noqa: E402

if not ref_loader.available():
    pytest.skip("no copy of the reference (run baseline/install_reference.py in the build container)", allow_module_level=True)


class _Rec:
    def __init__(self):
        self.s = {}

    def add_scalar(self, k, v, it):
        self.s.setdefault(k, []).append(float(v))


@pytest.fixture(scope="module")
def ref():
    from baseline import ref_loader
    if not ref_loader.available():
        pytest.skip("no copy of the reference sources (baseline/_ref is made by __graft_entry__.build() where /root/reference exists)")
    from baseline import ref_step as RS
    ns, args = RS.load(batch=8)
    return RS, ns, args


def _ours(args, seed=0):
    from oracle import weights as W
    from tris_b200 import clip_model
    from tris_b200.model_stage1 import TRIS
    args.synthetic_weights = True
    model = TRIS(args)
    model.load_state_dict(W.make_tris_state_dict(seed), strict=True)
    model = model.cuda()
    aux = clip_model.CLIPModel("ViT-B/32", txt_length=args.max_query_len)
    aux.load_state_dict(W.make_vitb32_clip_state_dict(7, cos_bias=True), strict=True)
    return model, aux.cuda().eval()


def test_reference_train_one_epoch_drives_tris_b200(ref):
    """Same loop, same torch.optim.AdamW / LambdaLR, same loader: reference modules vs this repo's modules."""
    RS, ns, args = ref
    B, n = 8, 3
    loader = RS.make_train_loader(n, B)
    losses = {}
    for who in ("reference", "tris_b200"):
        torch.manual_seed(0)
        if who == "reference":
            model, aux = RS.build_models(ns, args, "cuda", aux_half=False)
        else:
            model, aux = _ours(args)
        opt, sched = RS.make_optimizer(model, args, 1000)
        ns.T.writer = rec = _Rec()
        w0 = model.state_dict()["vis_project.weight"].clone()
        it = RS.train_epoch_reference_loop(ns, args, model, aux, loader, opt, sched)
        assert it == n
        assert not torch.equal(w0, model.state_dict()["vis_project.weight"]), "optimizer.step() did not move the weights"
        losses[who] = rec.s
        assert all(np.isfinite(v) for v in rec.s["train/loss"])
        del model, aux, opt
        torch.cuda.empty_cache()
    r, o = losses["reference"], losses["tris_b200"]
    print("train/loss reference", r["train/loss"], "tris_b200", o["train/loss"])
    # step 1 sees identical weights: bf16 tolerance of the loss (2e-2 at this batch of 8, 1e-2 at the benchmark batch).  Steps
    # 2-3 run on weights each model updated with ITS OWN gradients (the loss drops 35 -> 20 in one AdamW step at random init):
    # trajectories drift apart, the gate only says "same optimisation, not a different one"
    for k in ("train/loss", "train/l1", "train/l4", "train/l5"):
        for i, (a, b) in enumerate(zip(r[k], o[k])):
            tol = 2e-2 if i == 0 else 1e-1
            assert abs(a - b) <= tol * max(abs(a), 1.0), (k, i, r[k], o[k])
    assert r["optim/lr"] == o["optim/lr"]


def _run_validate(ns, args, model, loader, aux=None, prms=False, out_dir=None):
    V = ns.V
    args.cam_save_dir = os.path.join(out_dir, "cam")
    args.name_save_dir = os.path.join(out_dir, "names")
    args.print_freq = 1000
    args.save_cam = True            # validate_same_sentence reads args.save_cam, not its parameter (validate.py:258)
    if prms:
        # validate_same_sentence loads its scorer itself (validate.py:282); CLIP.clip is ONE module shared with the TRIS
        # constructor, so the patch is undone afterwards
        V.clip.load = lambda *a, **k: (aux, None)
        try:
            return V.validate_same_sentence(args, loader, model, 0, save_cam=True)
        finally:
            V.clip.load = ns.fake_load
    return V.validate(args, loader, model, 0, save_cam=True)


@pytest.mark.parametrize("prms", [False, True])
def test_reference_validate_drivers_on_tris_b200(ref, tmp_path, prms):
    """validate / validate_same_sentence of the reference, unmodified, over 6 synthetic refs x 3 sentences: metrics,
    saved file names and the saved CAMs of tris_b200 (fp32 parity mode) equal the reference modules' (fp32 on the GPU)."""
    RS, ns, args = ref
    loader = RS.make_val_loader(6, sentences=3)
    res = {}
    # PRMS picks the sentence whose masked image scores highest under the frozen ViT-B/32.  With random-init weights the S
    # candidate maps of a ref are near-ties, so BOTH runs are scored by the same (reference) scorer: the test isolates the
    # parity of TRIS through the driver; the scorer's own parity is test_prms_scorer_matches_reference below.
    _, ref_aux = RS.build_models(ns, args, "cuda", aux_half=False)
    for who in ("reference", "tris_b200"):
        if who == "reference":
            model, _ = RS.build_models(ns, args, "cuda", aux_half=False)
        else:
            model, _ = _ours(args)
            model.set_precision("fp32")
        aux = ref_aux
        d = str(tmp_path / who)
        os.makedirs(d, exist_ok=True)
        res[who] = (_run_validate(ns, args, model, loader, aux, prms, d), d)
        del model, aux
        torch.cuda.empty_cache()
    (mr, dr), (mo, do) = res["reference"], res["tris_b200"]
    print("reference", mr, "tris_b200", mo)
    fr, fo = sorted(os.listdir(os.path.join(dr, "cam"))), sorted(os.listdir(os.path.join(do, "cam")))
    assert fr == fo and len(fr) == (6 if prms else 18)
    nj = f"{args.dataset}_train_names.json" if prms else f"{args.dataset}_train_cam_name.json"
    assert json.load(open(os.path.join(dr, "names", nj))) == json.load(open(os.path.join(do, "names", nj)))
    for f in fr:
        a, b = np.load(os.path.join(dr, "cam", f)), np.load(os.path.join(do, "cam", f))
        assert a.shape == b.shape == (480, 640)
        assert np.abs(a - b).max() <= 2e-3, (f, np.abs(a - b).max())      # maps are normalised to max 1 (validate.py:183)
    for a, b in zip(mr, mo):
        assert abs(float(a) - float(b)) <= 0.5, (mr, mo)                   # metrics are percentages


def test_prms_scorer_matches_reference(ref):
    """get_scores (validate.py:120-127) with this repo's frozen ViT-B/32 + text tower against the reference modules."""
    RS, ns, args = ref
    _, ref_aux = RS.build_models(ns, args, "cuda", aux_half=False)
    _, aux = _ours(args)
    g = torch.Generator().manual_seed(5)
    fg = (torch.randn(4, 3, 224, 224, generator=g) * torch.rand(4, 1, 224, 224, generator=g)).cuda()
    from tris_b200.synthetic import synthetic_batch
    ids = synthetic_batch(3, 32, args.max_query_len, 0, seed=11)[1].long().cuda()
    with torch.no_grad():
        a = ns.V.get_scores(ref_aux, fg, ids)
        b = ns.V.get_scores(aux, fg, ids)
    print("get_scores max |diff|", (a - b.float()).abs().max().item(), "range", a.min().item(), a.max().item())
    assert (a - b.float()).abs().max().item() <= 1e-2


def test_reference_validate_prms_on_tris_b200_bf16(ref, tmp_path):
    """Same PRMS driver on the default bf16 path: the selected sentence (hence the saved CAM) must agree with the
    reference wherever the reference's own score margin between the best two sentences is not a numerical tie."""
    RS, ns, args = ref
    loader = RS.make_val_loader(6, sentences=3)
    model, aux = RS.build_models(ns, args, "cuda", aux_half=False)
    d_ref = str(tmp_path / "r")
    os.makedirs(d_ref)
    _run_validate(ns, args, model, loader, aux, True, d_ref)
    del model, aux
    model, aux = _ours(args)
    d_our = str(tmp_path / "o")
    os.makedirs(d_our)
    _run_validate(ns, args, model, loader, aux, True, d_our)
    agree = 0
    for f in sorted(os.listdir(os.path.join(d_ref, "cam"))):
        a, b = np.load(os.path.join(d_ref, "cam", f)), np.load(os.path.join(d_our, "cam", f))
        agree += int(np.abs(a - b).max() <= 8e-2)
    print("PRMS bf16: CAMs agreeing with the reference", agree, "of 6")
    assert agree >= 4     # near-tied candidates at random init may flip under bf16 (see the fp32 test above)


def test_repo_prms_driver_against_reference_driver(ref, tmp_path):
    """This repo's validate.py --prms --save_cam (RN50 once per ref, S instead of S^2 scorer passes, device-side selection and
    metrics, asynchronous .npy writer) against the reference's validate_same_sentence on the same synthetic refs: same file
    names, same names JSON, maps that agree wherever the candidates are not a numerical tie, mIoU within a point."""
    import validate as MyV
    RS, ns, args = ref
    n, S = 8, 2
    loader = RS.make_val_loader(n, sentences=S)
    model, aux = RS.build_models(ns, args, "cuda", aux_half=False)
    d_ref = str(tmp_path / "r")
    os.makedirs(d_ref)
    m_ref = _run_validate(ns, args, model, loader, aux, True, d_ref)
    del model, aux
    model, aux = _ours(args)
    d_our = str(tmp_path / "o")
    args.cam_save_dir, args.name_save_dir, args.save_cam = os.path.join(d_our, "cam"), os.path.join(d_our, "names"), True
    args.val_refs, args.no_graph, args.precision = n, False, "bf16"
    args.lanes = 3                  # three refs in flight on independent streams / CUDA graphs: same results as one
    miou = MyV.validate_same_sentence(args, MyV.synthetic_refs(args, n, sentences=S), model, aux)
    fr, fo = sorted(os.listdir(os.path.join(d_ref, "cam"))), sorted(os.listdir(os.path.join(d_our, "cam")))
    assert fr == fo and len(fr) == n
    nj = f"{args.dataset}_train_names.json"
    assert sorted(json.load(open(os.path.join(d_ref, "names", nj)))) == sorted(json.load(open(os.path.join(d_our, "names", nj))))
    agree = sum(int(np.abs(np.load(os.path.join(d_ref, "cam", f)) - np.load(os.path.join(d_our, "cam", f))).max() <= 8e-2) for f in fr)
    print("repo PRMS driver vs reference driver: maps agreeing", agree, "of", n, "| mIoU", 100 * miou, "vs", float(m_ref[1]))
    assert agree >= n - 2
    assert abs(100 * miou - float(m_ref[1])) <= 2.0
    # the plain driver (every sentence of every ref) against the reference's validate() on the same refs
    args.save_cam = False
    model_r, _ = RS.build_models(ns, args, "cuda", aux_half=False)
    d2 = str(tmp_path / "r2")
    os.makedirs(d2)
    m_ref2 = _run_validate(ns, args, model_r, loader, None, False, d2)
    args.save_cam, args.cam_save_dir = False, None
    miou2, hit2 = MyV.validate(args, MyV.synthetic_refs(args, n, sentences=S), model)
    print("repo validate driver vs reference validate: mIoU", 100 * miou2, "vs", float(m_ref2[1]), "| hit", 100 * hit2, "vs", float(m_ref2[2]))
    assert abs(100 * miou2 - float(m_ref2[1])) <= 2.0 and abs(100 * hit2 - float(m_ref2[2])) <= 100.0 / n + 1e-6
