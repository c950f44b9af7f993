"""One Stage-1 training step: forward, the three CLIP-guided losses, backward, gradient all-reduce, fused AdamW.

Restates the loop body of train_stage1.py:320-372 (fg / cls / negative losses, weights w1,w4,w5; AdamW with two
learning-rate groups and the per-step poly-0.9 schedule).  Differences, all deliberate:
  * the frozen auxiliary ViT-B/32 image tower runs ONCE per step (the reference encodes the same fg twice) and its
    backward is data-gradient only; the auxiliary text tower runs once on [B*(1+negs), L] without a graph;
  * data parallelism = one NCCL all-reduce over the flat gradient buffer, BatchNorm statistics stay per rank.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _lib as L
from . import dp

f32 = torch.float32


def mask_and_resize(sig_out, img, size=224):
    """train_stage1.py:327-339 (fg only; the bg branch is dead code in the reference) -> fg fp32 [B,3,size,size]."""
    from . import ops
    return ops.mask_resize_fwd(sig_out.contiguous(), img.float().contiguous(), size, 32, want_fg=True)[1]


def stage1_losses(model, aux, img, word_ids, neg_word_ids, w1=1.0, w4=5.0, w5=2.0):
    """-> dict(loss, l1, l4, l5) as device scalars (no host sync).  Every arithmetic step is a libtris_sm100 kernel:
    TRIS forward -> mask-and-resize straight into ViT patches -> frozen ViT-B/32 (once) + frozen text tower on the
    positives and negatives in one batch -> fused loss kernel."""
    from .engine import OVERLAP, masked_patches, stage1_loss
    B = img.shape[0]
    eng = aux._engine()
    k = 0 if neg_word_ids is None else neg_word_ids.shape[1]
    ids = word_ids if k == 0 else torch.cat([word_ids, neg_word_ids.reshape(-1, word_ids.shape[1])], 0)
    late = os.environ.get("TRIS_AUX_TEXT_LATE", "1") != "0"
    main, side = torch.cuda.current_stream(), eng.side_stream()

    def fork_text():
        # the frozen text tower depends on the token ids only: it runs on a side stream next to other work
        eng.ensure_fresh()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            return eng.encode_text_hidden(ids)

    g = None
    if OVERLAP and not late:
        g = fork_text()                       # under the RN50 forward
    cls, _, _, sig_out, _ = model(img, word_ids)
    if OVERLAP and late:
        g = fork_text()                       # under the frozen ViT forward (both are small-kernel chains)
    patches = masked_patches(sig_out, img)
    if not OVERLAP:
        g = eng.encode_text_hidden(ids)
    f = eng.encode_patches(patches, B)
    if OVERLAP:
        main.wait_stream(side)
        g.record_stream(main)
    loss, l1, l4, l5 = stage1_loss(f, g, cls, k, (w1, w4, w5))
    return {"loss": loss, "l1": l1, "l4": l4, "l5": l5}


class Stage1Trainer:
    def __init__(self, model, aux, max_iter: int, lr=5e-5, lr_multi=0.1, weight_decay=0.01, w=(1.0, 5.0, 2.0),
                 process_group=None):
        self.model, self.aux, self.w = model, aux, w
        self.eng = model.engine()
        self._zero_stream = None
        st = self.eng.store
        self.m = torch.zeros(st.n_train, device=st.device, dtype=f32)
        self.v = torch.zeros(st.n_train, device=st.device, dtype=f32)
        self.step_count = torch.zeros(1, device=st.device, dtype=torch.int32)
        self.max_iter, self.lr, self.lr_multi, self.wd = float(max_iter), lr, lr_multi, weight_decay
        self.world = dp.world_size(process_group)
        self.pg = process_group
        dp.broadcast_parameters(st.flat, process_group)      # every rank starts from rank 0's weights (DDP semantics)
        self.graph = None
        self.static = None
        # data parallel: the flat gradient buffer is [text head | image tower | text transformer + embeddings | new modules].
        # Everything behind the image tower (76 % of the payload) is final when the text-tower backward ends, early in the
        # RN50 backward: its all-reduce is issued there (engine._TextFn.backward hook) and overlaps the RN50 backward; the
        # image-tower prefix follows after the backward.  Two NCCL calls per step, the first one hidden.
        self._early_lo = None
        self._early_work = None
        self._graph_has_optimizer = True
        self._graph_signals = False
        self._flag = torch.zeros(1, device=st.device, dtype=torch.int32)
        self._flag_target = 0
        self._comm = torch.cuda.Stream(device=st.device) if self.world > 1 else None
        # STATUS (round 2): correct and bit-exact in eager steps (tests/dp_equivalence_worker.py, 2 GPUs), but the CUDA-graph
        # step hangs with it on this stack (both with NCCL captured in the graph and with the flag-released variant below),
        # so it is opt-in: TRIS_DP_OVERLAP=1.  Default = one all-reduce of the whole buffer after the replay, as in round 1.
        if self.world > 1 and os.environ.get("TRIS_DP_OVERLAP", "0") == "1":
            vis = [k for k in st.trainable if k.startswith("backbone.visual.")]
            if vis:
                last = max(st.offsets[k] + (st.shapes[k].numel() + 7) // 8 * 8 for k in vis)
                if 0 < last < st.n_train:
                    self._early_lo = last
                    self.eng.on_text_grads_ready = self._early_all_reduce
        st.publish_grads()            # param.grad = views of the flat buffer; the trainer zeroes the buffer itself
        for k in self.eng.extra_grad_keys:
            st.params[k].grad = st.g(k)

    def _early_all_reduce(self):
        """Called on the stream of the text-tower backward once every gradient behind the image tower is final.
        Eager step: issue the all-reduce right here.  While the step is being CAPTURED into a CUDA graph no NCCL call is
        recorded: the graph only bumps a device flag, and step() parks the all-reduce on the communication stream behind
        a kernel that waits for that flag (NCCL stays outside the graph, the overlap stays)."""
        st = self.eng.store
        if torch.cuda.is_current_stream_capturing():
            L.call("tris_flag_inc", C.c_void_p(self._flag.data_ptr()))
            self._graph_signals = True
            return
        self._early_work = dist.all_reduce(st.grad[self._early_lo: st.n_train], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def _park_early_all_reduce(self):
        """Graph mode: queue [wait for this step's flag] -> [all-reduce of everything behind the image tower] on the
        communication stream; it fires in the middle of the replayed backward."""
        st = self.eng.store
        self._flag_target += 1
        with torch.cuda.stream(self._comm):
            L.call("tris_flag_wait", C.c_void_p(self._flag.data_ptr()), C.c_uint32(self._flag_target & 0xFFFFFFFF))
            self._early_work = dist.all_reduce(st.grad[self._early_lo: st.n_train], op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def optimizer_step(self):
        st = self.eng.store
        if self._early_work is not None:
            self._early_work.wait()                                   # current stream waits for the overlapped part
            self._early_work = None
            dp.all_reduce_gradients(st.grad[: self._early_lo], self.pg)   # image-tower prefix (24 % of the payload)
        else:
            dp.all_reduce_gradients(st.grad[: st.n_train], self.pg)   # ONE collective per step (NCCL over NVLink)
        L.call("tris_adamw_step", C.c_void_p(st.flat.data_ptr()), C.c_void_p(st.grad.data_ptr()), C.c_void_p(self.m.data_ptr()),
               C.c_void_p(self.v.data_ptr()), C.c_void_p(st.shadow.data_ptr()), C.c_long(st.n_train),
               C.c_long(st.group_bounds[1]), C.c_void_p(self.step_count.data_ptr()), C.c_float(self.max_iter),
               C.c_float(self.lr * self.lr_multi), C.c_float(self.lr), C.c_float(0.9), C.c_float(0.999), C.c_float(1e-8),
               C.c_float(self.wd), C.c_float(1.0 / self.world), C.c_float(0.9), launches=2)
        # masters changed behind torch's back; the kernel rewrote the bf16 shadow too, so only the derived operands (packed
        # 3x3 / stem weights, padded BN vectors) must be re-derived by the next forward -- train OR eval
        self.eng.derived_stale = True

    def _fwd_bwd(self, img, word_ids, neg_word_ids):
        self.model.train()
        # clearing the 454 MB flat gradient buffer (61 us) runs next to the forward pass; joined in front of the backward pass
        main = torch.cuda.current_stream()
        side_zero = os.environ.get("TRIS_ZERO_SIDE", "1") != "0"
        if side_zero:
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream()
            self._zero_stream.wait_stream(main)
            with torch.cuda.stream(self._zero_stream):
                self.eng.store.zero_grad()
        else:
            self.eng.store.zero_grad()
        losses = stage1_losses(self.model, self.aux, img, word_ids, neg_word_ids, *self.w)
        if side_zero:
            main.wait_stream(self._zero_stream)
        losses["loss"].backward()
        return losses

    def _step_eager(self, img, word_ids, neg_word_ids):
        losses = self._fwd_bwd(img, word_ids, neg_word_ids)
        self.optimizer_step()
        return losses

    def step(self, img, word_ids, neg_word_ids):
        """One optimisation step on device-resident inputs; returns device scalars."""
        if self.graph is None:
            return self._step_eager(img, word_ids, neg_word_ids)
        s_img, s_ids, s_neg, s_out = self.static
        s_img.copy_(img, non_blocking=True)
        s_ids.copy_(word_ids, non_blocking=True)
        if s_neg is not None:
            s_neg.copy_(neg_word_ids, non_blocking=True)
        self.graph.replay()
        if self._graph_signals:
            # multi-rank: the bulk all-reduce waits on the comm stream for the signal from inside the replay.  It is parked
            # AFTER the replay was enqueued: a spinning kernel can stall work queued behind it on a shared hardware queue,
            # and everything queued later (the rest of optimizer_step) depends on this all-reduce anyway.
            self._park_early_all_reduce()
        if not self._graph_has_optimizer:     # multi-rank: the NCCL all-reduces + AdamW follow the replayed forward/backward
            self.optimizer_step()
        # the replayed AdamW ran AFTER the replayed re-derivation of the packed operands: an eval forward that follows
        # (validate after each epoch) must re-derive them from the updated masters
        self.eng.derived_stale = True
        return s_out

    def capture(self, img, word_ids, neg_word_ids, warmup=3):
        """Capture the whole step (fwd + bwd + all-reduce + AdamW) into one CUDA graph (static shapes)."""
        s_img, s_ids = img.clone(), word_ids.clone()
        s_neg = neg_word_ids.clone() if neg_word_ids is not None else None
        # the warm-up steps below are real optimizer steps on one batch: snapshot everything they mutate (masters, Adam
        # moments, step counter, BatchNorm running statistics) and put it back, so that capture() is invisible to the
        # training trajectory and to the poly-LR schedule
        snap = self._snapshot()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._step_eager(s_img, s_ids, s_neg)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # multi-rank: forward+backward are captured, NCCL and AdamW run eagerly behind the replay; the bulk all-reduce is
        # released from INSIDE the replay by a device flag (see _early_all_reduce).  TRIS_DP_GRAPH_NCCL=1 tries to capture the
        # whole step with the NCCL calls inside the graph instead (hung on this stack when tried: off by default).
        want_nccl_in_graph = self.world > 1 and os.environ.get("TRIS_DP_GRAPH_NCCL", "0") == "1"
        g, out = None, None
        if self.world == 1 or want_nccl_in_graph:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    out = self._step_eager(s_img, s_ids, s_neg)
                self._graph_has_optimizer = True
            except Exception as e:          # pragma: no cover - depends on the NCCL / driver combination
                if self.world == 1:
                    raise
                print(f"[tris_b200] capturing NCCL inside the step graph failed ({e!r}); falling back to eager collectives", flush=True)
                g = None
                torch.cuda.synchronize()
        if g is None:
            self._graph_has_optimizer = False
            self._graph_signals = False
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._fwd_bwd(s_img, s_ids, s_neg)
        self.graph, self.static = g, (s_img, s_ids, s_neg, out)
        self._restore(snap)
        return self

    # ---- state (checkpoint / resume; utils/util.py:50-96 of the reference stores 'optimizer' and 'lr_scheduler')
    def _snapshot(self):
        st = self.eng.store
        bufs = [b for b in self.model.buffers()]
        return {"flat": st.flat.clone(), "m": self.m.clone(), "v": self.v.clone(), "step": self.step_count.clone(),
                "bufs": [b.clone() for b in bufs]}

    def _restore(self, snap):
        st = self.eng.store
        with torch.no_grad():
            st.flat.copy_(snap["flat"]); self.m.copy_(snap["m"]); self.v.copy_(snap["v"]); self.step_count.copy_(snap["step"])
            for b, s in zip(self.model.buffers(), snap["bufs"]):
                b.copy_(s)
            # masters were rewritten by a plain copy: re-cast the shadow NOW (a captured graph only re-derives the packed
            # operands, the fused AdamW keeps the shadow current from here on)
            st.refresh_shadow()
            self.eng.refresh_derived()
        self.eng._seen_versions = self.eng._versions()
        self.eng.derived_stale = True
        torch.cuda.synchronize()

    def state_dict(self):
        """{'optimizer': Adam moments as flat fp32 tensors over the trainable prefix + step, 'lr_scheduler': schedule position}."""
        return {"optimizer": {"exp_avg": self.m.detach().cpu(), "exp_avg_sq": self.v.detach().cpu(), "step": int(self.step_count.item()),
                              "lr": self.lr, "lr_multi": self.lr_multi, "weight_decay": self.wd, "n_train": int(self.m.numel())},
                "lr_scheduler": {"last_epoch": int(self.step_count.item()), "max_iter": self.max_iter}}

    def load_state_dict(self, sd, start_step=None):
        """Restore Adam moments and the step counter (bias correction + poly-0.9 schedule position).  ``start_step``
        overrides the stored step (resume at --start_epoch N: N * steps_per_epoch).  Call it BEFORE capture(): the
        schedule length is a launch argument baked into the captured graph."""
        opt = sd["optimizer"]
        if int(opt["n_train"]) != self.m.numel():
            raise ValueError(f"optimizer state holds {opt['n_train']} elements, the model has {self.m.numel()} trainable")
        self.m.copy_(opt["exp_avg"].to(self.m.device)); self.v.copy_(opt["exp_avg_sq"].to(self.v.device))
        step = int(opt["step"]) if start_step is None else int(start_step)
        self.step_count.fill_(step)
        if "lr_scheduler" in sd and sd["lr_scheduler"].get("max_iter"):
            self.max_iter = float(sd["lr_scheduler"]["max_iter"])
        return self


class HostBatchPrefetcher:
    """Pinned-host -> device pipeline for the training loop: the H2D copy of batch i+1 is issued on a copy stream while
    step i computes (what DataLoader(pin_memory=True) + .cuda(non_blocking=True) gives the reference loop,
    train_stage1.py:118-123,302-316, when the copy is issued ahead of the step)."""

    def __init__(self):
        self.stream = torch.cuda.Stream()
        self._next = None

    def submit(self, host_batch):
        """Start copying a pinned host batch (tuple of tensors / None)."""
        with torch.cuda.stream(self.stream):        # fresh device buffers each time: nothing to wait for
            self._next = tuple(None if t is None else t.cuda(non_blocking=True) for t in host_batch)

    def take(self):
        """Device batch submitted last; the current stream waits for its copy."""
        cur = torch.cuda.current_stream()
        cur.wait_stream(self.stream)
        out, self._next = self._next, None
        for t in out:
            if t is not None:
                t.record_stream(cur)
        return out
