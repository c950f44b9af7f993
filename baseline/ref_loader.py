"""Import the UNMODIFIED reference (fawnliu/TRIS) with the stub recipe of SURVEY.md Appendix D.

Source root: $TRIS_REFERENCE_ROOT, else /root/reference (build container), else the shipped copy baseline/_ref/
(git-ignored; made by baseline/install_reference.py so that the reference can run on the GPU box).  Used by
tests/golden/make_golden*.py, by tests that cross-check the oracle / the product against the live reference
(skipped when no copy is present) and by the reference arms of bench.py.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root() -> str:
    for cand in (os.environ.get("TRIS_REFERENCE_ROOT"), "/root/reference", os.path.join(_HERE, "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "model")):
            return cand
    return os.path.join(_HERE, "_ref")


REF_ROOT = _find_root()

_STUBS = [
    "turtle", "tkinter", "tkinter.messagebox", "tkinter.tix", "ftfy", "termcolor", "tensorboardX",
    "matplotlib", "matplotlib.pyplot", "matplotlib.collections", "matplotlib.patches",
    "skimage", "skimage.io", "pycocotools", "pycocotools.mask", "pycocotools.coco", "imageio", "ema_pytorch",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "model"))


def install_stubs():
    for name in _STUBS:
        if name in sys.modules and name != "turtle":
            continue
        m = types.ModuleType(name)
        m.__path__ = []  # type: ignore[attr-defined]
        sys.modules[name] = m
    sys.modules["turtle"].forward = lambda *a, **k: None
    sys.modules["ftfy"].fix_text = lambda s: s
    sys.modules["termcolor"].colored = lambda s, *a, **k: s
    sys.modules["tensorboardX"].SummaryWriter = object
    sys.modules["tkinter"].E = None
    sys.modules["tkinter"].image_names = lambda *a, **k: ()
    sys.modules["matplotlib.collections"].PatchCollection = object
    sys.modules["matplotlib.patches"].Polygon = object
    sys.modules["matplotlib.patches"].Rectangle = object
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    def _get_cmap(name=None):
        # matplotlib is absent from this image: a grey ramp stands in for the colour map that utils/box_eval_utils.py
        # (box metrics, out of scope) turns heat-maps into before thresholding
        import numpy as _np
        return lambda x: _np.stack([_np.asarray(x, dtype="float64")] * 3 + [_np.ones_like(x, dtype="float64")], axis=-1)
    sys.modules["matplotlib.pyplot"].get_cmap = _get_cmap
    sys.modules["tkinter.messagebox"].NO = None
    sys.modules["tkinter.tix"].Tree = None
    sys.modules["ema_pytorch"].EMA = object


RN50 = dict(embed_dim=1024, image_resolution=224, vision_layers=(3, 4, 6, 3), vision_width=64,
            vision_patch_size=None, context_length=77, vocab_size=49408, transformer_width=512,
            transformer_heads=8, transformer_layers=12)
VITB32 = dict(embed_dim=512, image_resolution=224, vision_layers=12, vision_width=768,
              vision_patch_size=32, context_length=77, vocab_size=49408, transformer_width=512,
              transformer_heads=8, transformer_layers=12)


def load_reference():
    """Returns a namespace with the reference's TRIS class, CLIP module, arg parser and
    train_stage1's clip_forward / MaxLoss, with clip.load patched to random-init
    constructors (no network: SURVEY F8)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import CLIP.clip as clip  # noqa
    from CLIP.clip.model import CLIP  # noqa

    def fake_load(name, device="cpu", jit=False, download_root=None, txt_length=77):
        cfg = RN50 if "RN50" in name else VITB32
        return CLIP(txt_length=txt_length, **cfg).float().eval(), None

    clip.load = fake_load
    import CLIP.clip.clip as clip_mod
    clip_mod.load = fake_load
    from model.model_stage1 import TRIS  # noqa
    from args import get_parser  # noqa
    ns = types.SimpleNamespace(TRIS=TRIS, clip=clip, CLIP=CLIP, get_parser=get_parser, fake_load=fake_load)
    try:
        import train_stage1 as T  # noqa
        import validate as V  # noqa
        ns.T, ns.V = T, V
        ns.clip_forward, ns.MaxLoss = T.clip_forward, T.MaxLoss
    except Exception as e:  # pragma: no cover - dataset deps missing
        ns.train_import_error = repr(e)
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_clip_loss", os.path.join(REF_ROOT, "loss", "clip_loss.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        ns.clip_forward = mod.clip_forward
        ns.MaxLoss = None
    from CLIP.clip.model import convert_weights  # noqa
    ns.convert_weights = convert_weights
    _isolate()
    return ns


def _isolate():
    """The reference's top-level module names (`args`, `validate`, `train_stage1`, `model`, `utils`, ...) collide with this
    repo's own entry points: once everything is imported, park them under a `tris_reference.` prefix in sys.modules and take
    the reference root off sys.path, so that a later `import validate` / `from args import ...` finds this repo's files."""
    root = os.path.realpath(REF_ROOT)
    for name, mod in list(sys.modules.items()):
        d = getattr(mod, "__dict__", {})          # (no getattr on the module: lazy modules import things on attribute access)
        f = d.get("__file__")
        try:
            paths = [f] if f else list(d.get("__path__") or [])
        except Exception:
            paths = []
        if any(os.path.realpath(q).startswith(root + os.sep) or os.path.realpath(q) == root for q in paths if isinstance(q, str)):
            sys.modules["tris_reference." + name] = sys.modules.pop(name)
    while REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)


_NS = None


def load_reference_once():
    global _NS
    if _NS is None:
        _NS = load_reference()
    return _NS


def reference_args(ns, size=320, max_len=20, negs=3, batch=48):
    return ns.get_parser().parse_args(["--size", str(size), "--max_query_len", str(max_len),
                                       "--negative_samples", str(negs), "--batch_size", str(batch)])
