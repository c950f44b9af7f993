#!/usr/bin/env python
"""Stage-1 training entry point (counterpart of the reference's train_stage1.py:44-260,286-411,415-447).

    python train_stage1.py --synthetic --batch_size 48 --size 320 --max_query_len 20 --negative_samples 3 --epoch 1
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 train_stage1.py --synthetic --distributed ...

Same recipe as scripts/train_stage1.sh:3-16: AdamW (backbone lr*lr_multi, new modules lr, wd), poly-0.9 LR per step,
losses w1*fg + w4*cls + w5*neg with a frozen CLIP ViT-B/32, validation + checkpoint per epoch (state_dict layout of
utils/util.py:50-77: {"model", "epoch"} with the reference's 518 keys).  RefCOCO loading (dataset/ReferDataset.py) is
out of scope offline: without --synthetic the script stops with an explanation.
"""
import os
import sys
import time
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from args import get_parser  # noqa: E402


def main(args):
    from tris_b200 import clip_model as clip
    from tris_b200 import dp
    from tris_b200.model_stage1 import TRIS
    from tris_b200.synthetic import synthetic_batch
    from tris_b200.train_step import HostBatchPrefetcher, Stage1Trainer
    import validate as V
    rank, local, world = dp.env_rank()
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not args.synthetic:
        raise SystemExit("train_stage1.py: RefCOCO loaders are out of scope of this build (no dataset offline); use --synthetic")
    torch.manual_seed(1234)
    model = TRIS(args).cuda().train()
    if args.pretrain:
        ck = torch.load(args.pretrain, map_location="cpu")
        print("load:", model.load_state_dict(ck.get("model", ck), strict=False))
    aux, _ = clip.load("ViT-B-32", device="cuda", jit=False, txt_length=args.max_query_len,
                       allow_random_init=getattr(args, "synthetic_weights", None))
    warnings.simplefilter("ignore")      # only now: model construction / weight loading must stay loud
    max_iter = args.steps_per_epoch * args.epoch
    trainer = Stage1Trainer(model, aux, max_iter=max_iter, lr=args.lr, lr_multi=args.lr_multi, weight_decay=args.weight_decay,
                            w=(args.w1, args.w4, args.w5))
    if args.resume and args.pretrain and "optimizer" in ck:
        # utils/util.py:81-96 of the reference: resume restores optimizer + lr_scheduler + start_epoch
        trainer.load_state_dict(ck, start_step=args.start_epoch * args.steps_per_epoch if args.start_epoch else None)
        if not args.start_epoch:
            args.start_epoch = int(ck.get("epoch", -1)) + 1
        print(f"resume: optimizer state restored, step {int(trainer.step_count.item())}, start_epoch {args.start_epoch}")
    elif args.start_epoch:
        trainer.step_count.fill_(args.start_epoch * args.steps_per_epoch)      # schedule position without stored moments
    B, k = args.batch_size, args.negative_samples
    # a small pool of pinned synthetic batches, re-used round-robin (generating 48 x 3 x 320 x 320 normals on the host costs
    # more than the whole GPU step); each step still pays the H2D copy like a real DataLoader(pin_memory=True) batch
    pool = [synthetic_batch(B, args.size, args.max_query_len, k, seed=dp.shard_seed(1234, rank, i), pin=True) for i in range(4)]
    batch = lambda i: tuple(None if t is None else t.cuda(non_blocking=True) for t in pool[i % len(pool)])
    if not args.no_graph:
        trainer.capture(*batch(0), warmup=1)
    pf = HostBatchPrefetcher()
    pf.submit(pool[0])
    best = -1.0
    for epoch in range(args.start_epoch, args.epoch):
        t0, seen = time.time(), 0
        for it in range(args.steps_per_epoch):
            dev = pf.take()
            pf.submit(pool[(epoch * args.steps_per_epoch + it + 1) % len(pool)])      # H2D of the next batch under this step
            losses = trainer.step(*dev)
            seen += B * world
            if rank == 0 and (it + 1) % args.print_freq == 0 or it + 1 == args.steps_per_epoch:
                vals = {k_: float(v) for k_, v in losses.items()}        # one host sync per print, not per step
                dt = time.time() - t0
                if rank == 0:
                    print(f"epoch {epoch} it {it + 1}/{args.steps_per_epoch} loss {vals['loss']:.4f} l1 {vals['l1']:.4f} "
                          f"l4 {vals['l4']:.4f} l5 {vals['l5']:.4f}  {seen / dt:.0f} samples/s "
                          f"mem {torch.cuda.max_memory_allocated() / 2**20:.0f} MB", flush=True)
        miou, hit = V.validate(args, V.synthetic_refs(args, args.val_refs, rank, world), model, rank)
        if rank == 0:
            print(f"epoch {epoch} val mIoU {miou:.4f} hit {hit:.4f}")
            if args.output and miou > best:
                best = miou
                os.makedirs(args.output, exist_ok=True)
                # same keys as the reference's save_checkpoint (utils/util.py:50-77): model, optimizer, lr_scheduler, epoch
                torch.save({"model": model.state_dict(), "epoch": epoch, **trainer.state_dict()},
                           os.path.join(args.output, f"stage1_best_{epoch}.pth"))
        model.train()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main(get_parser().parse_args())
