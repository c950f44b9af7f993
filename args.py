"""Argument surface of the Stage-1 entry points: the flags of the reference's args.py:3-97 that the Stage-1 path reads
(same names, types and defaults), plus --synthetic / --steps_per_epoch / --val_refs for the offline synthetic data."""
import argparse


def get_parser():
    p = argparse.ArgumentParser(description="TRIS Stage-1 on tris_b200 (B200-native)")
    p.add_argument("--dataset", default="refcoco")
    p.add_argument("--max_query_len", default=20, type=int)
    p.add_argument("--negative_samples", default=0, type=int)
    p.add_argument("--positive_samples", default=1, type=int)
    p.add_argument("--bert_tokenizer", default="clip")
    p.add_argument("--refer_data_root", default="../../data/")
    p.add_argument("--splitBy", default="unc")
    p.add_argument("--lr", default=0.00005, type=float)
    p.add_argument("--weight-decay", "--weight_decay", default=0.01, type=float)
    p.add_argument("--lr_multi", default=0.1, type=float)
    p.add_argument("--batch_size", default=1, type=int, help="batch size per GPU")
    p.add_argument("--epoch", default=30, type=int)
    p.add_argument("--print-freq", default=100, type=int)
    p.add_argument("--size", default=384, type=int)
    p.add_argument("--resume", action="store_true")
    p.add_argument("--start_epoch", default=0, type=int)
    p.add_argument("--pretrain", default=None, type=str)
    p.add_argument("--eval", action="store_true")
    p.add_argument("--test_split", default="val", type=str)
    p.add_argument("--prms", action="store_true", default=False)
    p.add_argument("--output", default=None, type=str)
    p.add_argument("--distributed", action="store_true", default=False)
    p.add_argument("--attn_multi", default=0.1, type=float)
    p.add_argument("--w1", default=1, type=float)
    p.add_argument("--w4", default=5, type=float)
    p.add_argument("--w5", default=2, type=float)
    p.add_argument("--FOCAL_P", default=3, type=float)
    p.add_argument("--FOCAL_LAMBDA", default=0.01, type=float)
    p.add_argument("--backbone", default="clip-RN50", type=str)
    p.add_argument("--hidden_dim", default=1024, type=int)
    p.add_argument("--cam_save_dir", default=None, type=str)
    p.add_argument("--name_save_dir", default=None, type=str)
    p.add_argument("--save_cam", action="store_true", default=False)
    p.add_argument("--mode", default="clip", type=str)
    p.add_argument("--img", default=None, type=str)
    p.add_argument("--text", default=None, type=str)
    # ---- additions of this build
    p.add_argument("--synthetic", action="store_true", help="synthetic RefCOCOg-shaped data (the only source offline)")
    p.add_argument("--synthetic-weights", "--synthetic_weights", dest="synthetic_weights", action="store_true",
                   help="allow randomly initialised CLIP backbones when no checkpoint file is found (there is no download "
                        "path offline); without it a missing checkpoint is an error, like in the reference")
    p.add_argument("--steps_per_epoch", default=50, type=int)
    p.add_argument("--val_refs", default=64, type=int)
    p.add_argument("--lanes", default=3, type=int,
                   help="validate.py: refs in flight on independent CUDA streams (batch-1 inference is latency bound)")
    p.add_argument("--precision", default="bf16", choices=["bf16", "fp32"], help="fp32 = forward-only parity mode (validate/demo)")
    p.add_argument("--no_graph", action="store_true", help="do not capture the step in a CUDA graph")
    return p
