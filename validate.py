#!/usr/bin/env python
"""Stage-1 evaluation entry point (counterpart of the reference's validate.py:27-103,131-249,253-387).

    python validate.py --synthetic --size 320 --max_query_len 20 --val_refs 5000 [--prms --save_cam --cam_save_dir out/]

`validate`: per ref and sentence, response map -> bilinear(align_corners=True) to the original size -> /max ->
threshold 1e-9 -> IoU / pointing-game hit (validate.py:179-191).  `validate_same_sentence` (--prms): for every sentence j
of a ref build fg_j = cam_j * img at 224 and score it against ALL sentences of the ref with the frozen ViT-B/32; keep
the map with the highest summed score and dump it as `{idx}_{img_id}.npy` (validate.py:304-332,354-359).
Differences (zero numerical change): the RN50 tower runs once per ref (not once per sentence) and
each fg / sentence is encoded once (S ViT + S text passes instead of S^2 of each).  Box metrics (OpenCV contours,
utils/box_eval_utils.py) and dataset loading are out of scope; refs are synthetic.  Refs are sharded round-robin over
ranks with no communication ("replicas only").
"""
import json
import os
import sys
import warnings

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from args import get_parser  # noqa: E402


def synthetic_refs(args, n, rank=0, world=1, sentences=2, orig=(480, 640)):
    """Generator of (idx, img[1,3,S,S], word_ids[1,L,S_], target[1,oH,oW] int64) like the eval-mode ReferDataset."""
    from tris_b200 import dp
    from tris_b200.synthetic import synthetic_batch
    for idx in dp.shard_range(n, rank, world):
        img, _, _ = synthetic_batch(1, args.size, args.max_query_len, 0, seed=77000 + idx)
        _, ids, _ = synthetic_batch(sentences, 32, args.max_query_len, 0, seed=99000 + idx)
        g = torch.Generator().manual_seed(idx)
        target = torch.zeros((1, *orig), dtype=torch.int64)
        y0, x0 = int(torch.randint(0, orig[0] // 2, (1,), generator=g)), int(torch.randint(0, orig[1] // 2, (1,), generator=g))
        target[:, y0:y0 + orig[0] // 3, x0:x0 + orig[1] // 3] = 1
        yield idx, img, ids.t().unsqueeze(0).contiguous(), target


def _prefetch(gen, depth=8):
    """Run the (CPU) ref generator in a background thread and hand out PINNED tensors, so that the .cuda(non_blocking=True) copies
    of the loop are asynchronous and the next refs are produced while the current one is being launched."""
    import queue
    import threading
    q = queue.Queue(maxsize=depth)
    end = object()
    failure = []

    def work():
        try:
            for idx, img, ids, target in gen:
                q.put((idx, img.pin_memory(), ids.pin_memory(), target.pin_memory()))
        except BaseException as e:      # re-raised in the consumer: a failing loader must not look like an empty one
            failure.append(e)
        finally:
            q.put(end)
    threading.Thread(target=work, daemon=True).start()
    while True:
        item = q.get()
        if item is end:
            if failure:
                raise failure[0]
            return
        yield item


class _Lanes:
    """K refs in flight: each lane owns a Stage1Inference (its own CUDA graphs and static buffers) and a stream.  At batch 1 a
    ref is a chain of ~700 latency-bound kernels that occupy a few SMs each, so independent refs overlap almost perfectly."""

    def __init__(self, k, make):
        self.main = torch.cuda.current_stream()
        self.lanes = [(make(), torch.cuda.Stream()) for _ in range(max(1, k))]
        for _, st in self.lanes:
            st.wait_stream(self.main)
        self.i = 0

    def next(self):
        lane = self.lanes[self.i % len(self.lanes)]
        self.i += 1
        return lane

    def join(self):
        for _, st in self.lanes:
            self.main.wait_stream(st)


def _to_original(cam, size):
    from tris_b200 import ops
    return ops.resize_bilinear_ac(cam.contiguous(), size[0], size[1])


def _finish(stats, n):
    """One device->host read for the whole run: stats f32 [n, 4] = I, U, hit, max per evaluated map."""
    h = stats[:n].cpu().double().numpy()
    iou = h[:, 0] / np.maximum(h[:, 1], 1.0)
    return (float(iou.mean()) if n else 0.0), (float(h[:, 2].mean()) if n else 0.0)


@torch.no_grad()
def validate(args, refs, model, local_rank=0):
    """validate.py:131-249 of the reference, per ref and sentence: map -> bilinear(align_corners=True) to the original
    size -> /max -> threshold 1e-9 -> IoU / pointing-game hit.  Everything stays on the device (tris_cam_metrics); the
    host reads the per-map statistics once at the end instead of three times per sentence."""
    from tris_b200 import ops
    from tris_b200.infer import AsyncCamWriter, Stage1Inference
    graphs = not getattr(args, "no_graph", False) and getattr(args, "precision", "bf16") == "bf16"
    stats = torch.zeros((max(1, args.val_refs) * 8, 4), device="cuda")
    model.eval().engine().ensure_fresh()
    torch.cuda.synchronize()                  # derived operands (side-stream weight packs) complete before any lane starts
    lanes = _Lanes(getattr(args, "lanes", 1) if graphs else 1, lambda: Stage1Inference(model, use_graphs=graphs))
    writer = AsyncCamWriter(args.cam_save_dir) if (args.save_cam and args.cam_save_dir) else None
    n = 0
    for idx, img, word_ids, target in refs:
        inf, st = lanes.next()
        with torch.cuda.stream(st):
            img, word_ids, target = img.cuda(non_blocking=True), word_ids.cuda(non_blocking=True), target.cuda(non_blocking=True)
            c4 = inf.features(img)                                  # once per ref
            for j in range(word_ids.shape[-1]):
                out = inf.respond(c4, word_ids[:, :, j].contiguous(), tuple(img.shape[2:]))
                cam = _to_original(out, target.shape[-2:])
                if n >= stats.shape[0]:
                    lanes.join()
                    stats = torch.cat([stats, torch.zeros_like(stats)])
                    torch.cuda.synchronize()
                cam_n, _ = ops.cam_metrics(cam, target[0], stats=stats[n])
                n += 1
                if writer is not None:
                    writer.submit(f"{idx}_{j}", cam_n)
    lanes.join()
    if writer is not None:
        writer.close()
    return _finish(stats, n)


@torch.no_grad()
def validate_same_sentence(args, refs, model, aux, local_rank=0):
    """PRMS map selection (validate.py:253-387) with per-ref de-duplicated encoders: the RN50 tower runs once per ref, each
    candidate map and each sentence is encoded once (S ViT + S text passes instead of S^2), and the choice of the map
    (tris_prms_select), the resize and the metrics (tris_cam_metrics) never leave the device."""
    from tris_b200 import ops
    from tris_b200.infer import AsyncCamWriter, Stage1Inference
    graphs = not getattr(args, "no_graph", False) and getattr(args, "precision", "bf16") == "bf16"
    stats = torch.zeros((max(1, args.val_refs), 4), device="cuda")
    model.eval().engine().ensure_fresh()
    aux._engine().ensure_fresh()
    torch.cuda.synchronize()                  # derived operands (side-stream weight packs) complete before any lane starts
    lanes = _Lanes(getattr(args, "lanes", 1) if graphs else 1, lambda: Stage1Inference(model, aux, use_graphs=graphs))
    writer = AsyncCamWriter(args.cam_save_dir) if (args.save_cam and args.cam_save_dir) else None
    names, n = [], 0
    for idx, img, word_ids, target in refs:
        inf, st = lanes.next()
        with torch.cuda.stream(st):
            img, word_ids, target = img.cuda(non_blocking=True), word_ids.cuda(non_blocking=True), target.cuda(non_blocking=True)
            S = word_ids.shape[-1]
            ids = word_ids[0].t().contiguous()                                     # [S, L]
            c4 = inf.features(img)
            cams = torch.cat([inf.respond(c4, ids[j:j + 1], tuple(img.shape[2:])).clone() for j in range(S)])   # [S,1,H,W]
            f, g = inf.prms_features(cams, img, ids)
            best, _ = ops.prms_select(f, g)                                         # device-side arg-max of the summed get_scores
            big = _to_original(cams, target.shape[-2:])                             # [S,1,oH,oW]
            if n >= stats.shape[0]:
                lanes.join()
                stats = torch.cat([stats, torch.zeros_like(stats)])
                torch.cuda.synchronize()
            cam_n, _ = ops.cam_metrics(big, target[0], sel=best, stats=stats[n])
            n += 1
            if writer is not None:
                writer.submit(f"{idx}_{idx}", cam_n)                                 # {idx}_{img_id}.npy, written in the background
    lanes.join()
    if writer is not None:
        names = writer.close()
    if args.save_cam and args.name_save_dir:
        os.makedirs(args.name_save_dir, exist_ok=True)
        json.dump(names, open(os.path.join(args.name_save_dir, f"{args.dataset}_train_names.json"), "w"))
    return _finish(stats, n)[0]


def main(args):
    import time
    from tris_b200 import clip_model as clip
    from tris_b200 import dp
    from tris_b200.model_stage1 import TRIS
    rank, local, world = dp.env_rank()
    torch.cuda.set_device(local)
    if not args.synthetic:
        raise SystemExit("validate.py: RefCOCO loaders are out of scope of this build (no dataset offline); use --synthetic")
    torch.manual_seed(0)          # --synthetic-weights: the same random initialisation on every run
    model = TRIS(args).cuda().set_precision(args.precision)
    if args.pretrain:
        ck = torch.load(args.pretrain, map_location="cpu")
        print("load:", model.load_state_dict(ck.get("model", ck), strict=False))
    refs = _prefetch(synthetic_refs(args, args.val_refs, rank, world))
    t0 = time.time()
    if args.prms:
        aux, _ = clip.load("ViT-B/32", device="cuda", jit=False, txt_length=args.max_query_len,
                           allow_random_init=getattr(args, "synthetic_weights", None))
        miou = validate_same_sentence(args, refs, model, aux, rank)
        torch.cuda.synchronize()
        print(f"rank {rank}: PRMS mIoU {miou:.4f}  {len(dp.shard_range(args.val_refs, rank, world)) / (time.time() - t0):.1f} refs/s")
    else:
        miou, hit = validate(args, refs, model, rank)
        torch.cuda.synchronize()
        print(f"rank {rank}: mIoU {miou:.4f} hit {hit:.4f}  {len(dp.shard_range(args.val_refs, rank, world)) / (time.time() - t0):.1f} refs/s")


if __name__ == "__main__":
    main(get_parser().parse_args())
