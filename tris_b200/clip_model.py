"""CLIP container + ``load`` with the reference's call signature (CLIP/clip/clip.py:94-197).

``CLIPModel`` exposes the reference's attribute / state_dict layout (``visual``, ``transformer``,
``token_embedding``, ``positional_embedding``, ``ln_final``, ``text_projection``, ``logit_scale``) and the two
methods the Stage-1 drivers call on the auxiliary model: ``encode_image(img) -> [N, 512]`` and
``encode_text(ids) -> (x, [N, 512])`` (train_stage1.py:264-265,344,347; validate.py:121-122).  Used standalone
it is the frozen ViT-B/32 scorer; as ``TRIS.backbone`` its parameters are driven by the owning model's engine.
"""
from __future__ import annotations

import os
import warnings

import torch
from torch import nn

from . import _lib as L
from . import spec as S

_ALIASES = {"ViT-B/32": "ViT-B/32", "ViT-B-32": "ViT-B/32", "ViT-B-32.pt": "ViT-B/32", "RN50": "RN50", "RN101": "RN101"}
_FILES = {"ViT-B/32": "ViT-B-32.pt", "RN50": "RN50.pt", "RN101": "RN101.pt"}


def _canonical(name: str) -> str:
    base = os.path.basename(name)
    for k, v in _ALIASES.items():
        if base == k or base == k + ".pt" or name == k:
            return v
    if "RN101" in base:
        return "RN101"
    if "RN50" in base:
        return "RN50"
    if "ViT" in base:
        return "ViT-B/32"
    raise RuntimeError(f"Model {name} not found; available models = ['RN50', 'RN101', 'ViT-B/32']")


class CLIPModel(nn.Module):
    def __init__(self, kind: str = "RN50", txt_length: int = 77):
        super().__init__()
        self.kind = kind
        self.txt_length = txt_length
        self.is_vit = kind.startswith("ViT")
        self.context_length = 77
        entries = S.clip_vit_spec() if self.is_vit else S.clip_resnet_spec(kind)
        S.build_tree(self, entries)
        self._eng = None

    # ------------------------------------------------------------------ engine (standalone use = frozen aux model)
    def _engine(self):
        dev = self.positional_embedding.device
        if dev.type != "cuda":
            raise L.TrisLibError("tris_b200.CLIPModel runs on CUDA (sm_100a) only; there is no CPU path")
        if self._eng is None or not self._eng.store.still_bound() or self._eng.store.device != dev:
            from .engine import ClipEngine
            self._eng = ClipEngine(self, dev)
        return self._eng

    @property
    def dtype(self):
        return torch.float32

    def encode_image(self, image):
        return self._engine().encode_image(image)

    def encode_text(self, text):
        return self._engine().encode_text(text)

    def forward(self, image, text):
        f = self.encode_image(image)
        g = self.encode_text(text)[1]
        f = f / f.norm(dim=1, keepdim=True)
        g = g / g.norm(dim=1, keepdim=True)
        logits = self.logit_scale.exp() * f @ g.t()
        return logits, logits.t()


def _find_checkpoint(name: str, download_root=None):
    if os.path.isfile(name):
        return name
    kind = _canonical(name)
    for root in (download_root, os.path.expanduser("~/.cache/clip"), "."):
        if root and os.path.isfile(os.path.join(root, _FILES[kind])):
            return os.path.join(root, _FILES[kind])
    return None


def load(name: str, device="cuda" if torch.cuda.is_available() else "cpu", jit: bool = False, download_root: str = None,
         txt_length: int = 77, allow_random_init: bool = None):
    """Same signature / return convention as the reference's ``clip.load`` -> (model, preprocess).

    Accepts both "ViT-B/32" and the reference's "ViT-B-32" spelling (train_stage1.py:167).  Looks for an
    OpenAI checkpoint on disk (state_dict or TorchScript archive).  There is no download path in this build: like the
    reference (which downloads or raises, clip.py:122-127) a missing checkpoint is an ERROR, unless random initialisation
    is asked for explicitly -- ``allow_random_init=True`` (synthetic benchmarks / parity tests that load their own
    state_dict afterwards, drivers run with ``--synthetic-weights``) or the environment variable TRIS_ALLOW_RANDOM_INIT=1.
    """
    kind = _canonical(name)
    model = CLIPModel(kind, txt_length=txt_length)
    path = _find_checkpoint(name, download_root)
    if path is not None:
        try:
            sd = torch.jit.load(path, map_location="cpu").state_dict()
        except RuntimeError:
            sd = torch.load(path, map_location="cpu")
        sd = {k: v.float() if v.is_floating_point() else v for k, v in sd.items()
              if k not in ("input_resolution", "context_length", "vocab_size")}
        missing = model.load_state_dict(sd, strict=False)
        if missing.missing_keys:
            warnings.warn(f"clip.load({name}): missing keys {missing.missing_keys[:4]}...")
    else:
        if allow_random_init is None:
            allow_random_init = os.environ.get("TRIS_ALLOW_RANDOM_INIT", "0") == "1"
        if not allow_random_init:
            raise RuntimeError(
                f"clip.load({name!r}): no checkpoint file found (looked for {_FILES[kind]} in download_root, ~/.cache/clip and '.'); "
                "this build cannot download weights.  Pass a checkpoint path, or opt in to random initialisation with "
                "allow_random_init=True / --synthetic-weights / TRIS_ALLOW_RANDOM_INIT=1")
    model = model.to(device).eval()
    return model, None
