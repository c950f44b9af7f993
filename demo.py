#!/usr/bin/env python
"""Single image + sentence -> response heat-map (counterpart of the reference's demo.py:50-100, with the Stage-1 network:
the reference's demo imports the Stage-2 model).

    python demo.py --synthetic                      # random 320x320 image, 'man on the right' token ids
    python demo.py --img figs/demo.png --text 'man on the right' --pretrain weights/stage1_refcocog_google.pth

Pre-processing follows demo.py:50-68 (cv2.imread BGR kept as is -> resize to size x size -> /255 -> ImageNet mean/std);
post-processing demo.py:41-48,94 (bilinear align_corners=True to the original size, min-max normalisation).  The BPE
tokenizer is out of scope: --text is looked up in a tiny table of pre-tokenised phrases, or pass
--token_ids "49406,786,525,518,1155,49407".
"""
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from args import get_parser  # noqa: E402

KNOWN = {"man on the right": [49406, 786, 525, 518, 1155, 49407]}     # SURVEY 3.5 [probe]


def prepare(args):
    L = args.max_query_len
    ids = torch.zeros((1, L), dtype=torch.int32)
    toks = [int(t) for t in args.token_ids.split(",")] if args.token_ids else KNOWN.get(args.text or "man on the right")
    if toks is None:
        raise SystemExit("demo.py: no tokenizer in this build; pass --token_ids")
    ids[0, :min(L, len(toks))] = torch.tensor(toks[:L], dtype=torch.int32)
    if args.img and not args.synthetic:
        import cv2
        im = cv2.imread(args.img)
        h, w = im.shape[:2]
        im = cv2.resize(im, (args.size, args.size), interpolation=cv2.INTER_LINEAR).astype(np.float32) / 255.0
        mean, std = np.array([0.485, 0.456, 0.406], np.float32), np.array([0.229, 0.224, 0.225], np.float32)
        img = torch.from_numpy(((im - mean) / std).transpose(2, 0, 1)).unsqueeze(0)
    else:
        h, w = 480, 640
        img = torch.randn(1, 3, args.size, args.size, generator=torch.Generator().manual_seed(0))
    return img.contiguous(), ids, (h, w)


def main(args):
    from tris_b200 import ops
    from tris_b200.model_stage1 import TRIS
    img, ids, (h, w) = prepare(args)
    torch.manual_seed(0)          # --synthetic-weights: the same random initialisation on every run
    model = TRIS(args).cuda().set_precision(args.precision).eval()
    if args.pretrain:
        ck = torch.load(args.pretrain, map_location="cpu")
        print("load:", model.load_state_dict(ck.get("model", ck), strict=False))
    with torch.no_grad():
        out = model(img.cuda(), ids.cuda())                                  # [1,1,S,S]
        cam = ops.resize_bilinear_ac(out, h, w)[0, 0]
    raw = (cam.min().item(), cam.max().item())
    cam = (cam - cam.min()) / (cam.max() - cam.min() + 1e-5)
    path = args.output or "demo_cam.npy"
    np.save(path, cam.cpu().numpy())
    print(f"response map {tuple(cam.shape)} raw range [{raw[0]:.4f}, {raw[1]:.4f}] -> min-max normalised "
          f"[{cam.min().item():.3f}, {cam.max().item():.3f}] -> {path}")


if __name__ == "__main__":
    p = get_parser()
    p.add_argument("--token_ids", default=None, type=str)
    a = p.parse_args()
    if a.size == 384:
        # the reference's demo.py ignores --size (argparse default 384) and resizes to its module constant img_size = 320
        # (demo.py:64); do the same, but say so
        print("demo.py: --size left at the argparse default 384 -> using 320 like the reference demo (pass --size explicitly to override)")
        a.size = 320
    main(a)
