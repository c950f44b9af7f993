"""GPU: the fp32 parity mode (model.set_precision("fp32"), tris_b200/precise.py + csrc/precise.cu) against the golden
vectors produced by the UNMODIFIED reference in fp32 (tests/golden/stage1_golden.npz).  North-star tolerance for response
maps in fp32: 1e-3 relative."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / (np.abs(b).max() + 1e-12)


@pytest.fixture(scope="module")
def setup(golden):
    import warnings
    warnings.simplefilter("ignore")
    from oracle import weights as W
    from tris_b200.model_stage1 import TRIS
    b, size, l, neg, sub, s_tris, s_aux, s_data = [int(v) for v in golden["meta"]]
    args = argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=l, hidden_dim=1024, attn_multi=0.1, FOCAL_P=3,
                              FOCAL_LAMBDA=0.01)
    model = TRIS(args)
    model.load_state_dict(W.make_tris_state_dict(s_tris), strict=True)
    model = model.cuda().set_precision("fp32")
    img, ids, _ = W.synthetic_batch(b, size, l, neg, s_data)
    return dict(model=model, img=img.cuda(), ids=ids.cuda(), sub=sub)


def test_fp32_eval_response_map_vs_reference(setup, golden):
    m = setup["model"].eval()
    with torch.no_grad():
        out = m(setup["img"][:1], setup["ids"][:1])
    s = setup["sub"]
    assert out.shape == (1, 1, 320, 320) and out.dtype == torch.float32
    e = rel(out[:, :, ::s, ::s], golden["eval_relu_sub"])
    print("eval relu map rel err", e)
    assert e < 1e-3
    assert abs(out.double().sum().item() / golden["eval_relu_sum"][0] - 1) < 1e-3


def test_fp32_train_forward_vs_reference(setup, golden):
    m = setup["model"].train()
    with torch.no_grad():
        cls, cls_fg, relu_map, sig, ls = m(setup["img"], setup["ids"])
    s = setup["sub"]
    errs = dict(cls=rel(cls, golden["cls_out"]), cls_fg=rel(cls_fg, golden["cls_fg"]), relu=rel(relu_map[:, :, ::s, ::s], golden["relu_sub"]),
                sig=rel(sig[:, :, ::s, ::s], golden["sig_sub"]))
    print("train forward rel errs", errs)
    for k, e in errs.items():
        assert e < 1e-3, (k, e)
    assert abs(ls.item() - float(golden["logit_scale_exp"])) < 1e-3


def test_fp32_mode_is_forward_only(setup):
    from tris_b200 import _lib as L
    m = setup["model"].train()
    with pytest.raises(L.TrisLibError):
        m(setup["img"], setup["ids"])
    with pytest.raises(ValueError):
        m.set_precision("fp16")
    m.set_precision("fp32")


def test_fp32_step_losses_vs_reference(setup, golden):
    """Forward of the WHOLE training step in fp32 (TRIS forward, mask-and-resize, frozen ViT-B/32 + text tower on positives
    and negatives, the three losses) against the losses the unmodified reference produced (train_stage1.py:320-364)."""
    from oracle import weights as W
    from tris_b200 import clip_model
    from tris_b200.precise import PreciseStage1
    b, size, l, neg, sub, s_tris, s_aux, s_data = [int(v) for v in golden["meta"]]
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=l, allow_random_init=True)
    aux.load_state_dict(W.make_vitb32_clip_state_dict(s_aux, cos_bias=True), strict=True)
    _, _, negs = W.synthetic_batch(b, size, l, neg, s_data)
    m = setup["model"].train()
    out = PreciseStage1(m).step_losses(aux, setup["img"], setup["ids"], negs.cuda()).cpu().numpy()
    print("fp32 step losses", out, golden["losses"])
    np.testing.assert_allclose(out, golden["losses"], rtol=1e-3)
