"""Golden vectors for the configuration edges, from the UNMODIFIED reference on CPU (build container only):

    python tests/golden/make_golden_variants.py      ->  tests/golden/variants_golden.npz

Cases: (size 224, batch 2, fusion on), (size 320, batch 2, --attn_multi 0), (size 384, batch 1, fusion on).
For each: train-mode 5-tuple and eval-mode map of reference TRIS.forward on oracle.weights parameters / inputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, HERE)

import ref_loader  # noqa: E402
from oracle import weights as W  # noqa: E402

CASES = [("s224_b2_fuse", 224, 2, 0.1), ("s320_b2_nofuse", 320, 2, 0.0), ("s384_b1_fuse", 384, 1, 0.1)]
SUB = 8


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    ns = ref_loader.load_reference()
    sd = W.make_tris_state_dict(seed=0)
    out = {"names": np.array([c[0] for c in CASES]), "sub": np.array([SUB])}
    for name, size, b, attn in CASES:
        args = ref_loader.reference_args(ns, size, 20, 0, b)
        args.attn_multi = attn
        model = ns.TRIS(args)
        print(name, model.load_state_dict(sd, strict=False))
        img, ids, _ = W.synthetic_batch(b, size, 20, 0, seed=77)
        model.train()
        with torch.no_grad():
            cls, fg, relu, sig, ls = model(img, ids)
        model.load_state_dict(sd, strict=False)      # undo the running-stat update
        model.eval()
        with torch.no_grad():
            ev = model(img, ids)
        out[name + "/meta"] = np.array([size, b, attn])
        out[name + "/cls_out"] = cls.numpy()
        out[name + "/cls_fg"] = fg.numpy()
        out[name + "/relu_sub"] = relu[:, :, ::SUB, ::SUB].numpy()
        out[name + "/sig_sub"] = sig[:, :, ::SUB, ::SUB].numpy()
        out[name + "/eval_relu_sub"] = ev[:, :, ::SUB, ::SUB].numpy()
        print(name, "cls", cls.flatten()[:3], "eval max", ev.max().item())
    np.savez_compressed(os.path.join(HERE, "variants_golden.npz"), **out)
    print("written", os.path.getsize(os.path.join(HERE, "variants_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
