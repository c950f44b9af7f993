"""CLIP ModifiedResNet (RN50/RN101) image tower on NHWC bf16 with hand-written kernels: forward and backward.

Arithmetic restated from CLIP/clip/model.py:10-55 (Bottleneck), :212-232 + :254-279 (stem / layers) of the reference;
the attention pool (:58-104) is NOT computed here -- Stage-1 discards it (model_stage1.py:59).

Every conv is a tcgen05 GEMM (gemm.py); the first stem conv goes through an im2col of the fp32 NCHW image.  The two
32-channel stem activations must fill 128-byte TMA rows: even batches pack the image pair (2i, 2i+1) into one 64-channel
row with block-diagonal weights (half the pixels, no padding, BatchNorm sums of the two halves folded); odd batches carry
them zero-padded to 64 channels.  BatchNorm batch statistics come out of the GEMM epilogues; BN-apply/ReLU/pool/residual
are one streaming kernel.  In backward the weight gradients (leaves of the chain) run on a side stream.
"""
from __future__ import annotations

import os
from typing import Dict, List

import torch

from . import gemm as G
from . import ops
from .ops import BNState

bf16, f32 = torch.bfloat16, torch.float32
PAIR_STEM = os.environ.get("TRIS_STEM_PAIR", "1") != "0"   # image-pair-packed stem for even batches
# BatchNorm-backward reductions (sum g, sum g (y - mean)) fused into the epilogue of the GEMM that produces the upstream
# gradient (which then stores the ReLU-masked g directly): removes one HBM pass per BatchNorm layer in backward
# (measured round 2: with y read straight from global memory in the epilogue the GEMMs slow down by more than the
# reduction kernels cost -- tools/bench_bn_fusion.py -- so the fusion is OFF by default until y is staged through TMA)
FUSE_BN_BWD = os.environ.get("TRIS_FUSE_BN_BWD", "0") != "0"


class _Block:
    __slots__ = ("p", "cin", "planes", "stride", "down")


class ResNetTower:
    def __init__(self, store, module, prefix: str, layers=(3, 4, 6, 3)):
        self.store, self.prefix, self.layers = store, prefix, layers
        dev = store.device
        self.blocks: List[_Block] = []
        inpl = 64
        for li, (planes, nb) in enumerate(zip((64, 128, 256, 512), layers), start=1):
            for b in range(nb):
                blk = _Block()
                blk.p = f"{prefix}layer{li}.{b}."
                blk.cin, blk.planes = inpl, planes
                blk.stride = 2 if (b == 0 and li > 1) else 1
                blk.down = blk.stride > 1 or inpl != planes * 4
                self.blocks.append(blk)
                inpl = planes * 4
        self.out_channels = inpl
        # ---- BN states (real ones are views of the store / module buffers)
        self.bn: Dict[str, BNState] = {}
        bufs = dict(module.named_buffers())

        def real_bn(key):
            return BNState(store.p(key + ".weight"), store.p(key + ".bias"), bufs[key + ".running_mean"],
                           bufs[key + ".running_var"], store.g(key + ".weight"), store.g(key + ".bias"))

        for blk in self.blocks:
            for nm in ("bn1", "bn2", "bn3") + (("downsample.1",) if blk.down else ()):
                self.bn[blk.p + nm] = real_bn(blk.p + nm)
        self.bn[prefix + "bn3"] = real_bn(prefix + "bn3")
        # padded (32 -> 64 channel) stem BNs: private fp32 copies
        self.pad_bn: Dict[str, BNState] = {}
        for nm in ("bn1", "bn2"):
            z = lambda: torch.zeros(64, device=dev, dtype=f32)
            st = BNState(z(), z(), z(), torch.ones(64, device=dev, dtype=f32), z(), z())
            self.pad_bn[prefix + nm] = st
        self.real_small_bn = {prefix + nm: real_bn(prefix + nm) for nm in ("bn1", "bn2")}
        # image-pair-packed stem (even batches): channels [c | c + C] of a packed row are channel c of images 2i / 2i+1
        self.pair_bn: Dict[str, BNState] = {}
        for nm, cc in (("bn1", 32), ("bn2", 32), ("bn3", 64)):
            z = lambda: torch.zeros(2 * cc, device=dev, dtype=f32)
            self.pair_bn[prefix + nm] = BNState(z(), z(), z(), torch.ones(2 * cc, device=dev, dtype=f32), z(), z())
        self.pair_real = {prefix + nm: real_bn(prefix + nm) for nm in ("bn1", "bn2", "bn3")}
        self.w_pair1 = torch.zeros((64, 64), device=dev, dtype=bf16)
        self.w_pair2 = torch.zeros((64, 9 * 64), device=dev, dtype=bf16)
        self.w_pair3 = torch.zeros((128, 9 * 64), device=dev, dtype=bf16)
        # ---- packed weights
        self.w_stem1 = torch.zeros((64, 64), device=dev, dtype=bf16)
        self.w_stem2 = torch.zeros((64, 9 * 64), device=dev, dtype=bf16)
        self.w_stem3 = torch.zeros((64, 9 * 64), device=dev, dtype=bf16)
        self.w3x3 = {blk.p: torch.empty((blk.planes, 9 * blk.planes), device=dev, dtype=bf16) for blk in self.blocks}
        n_stats = 2 * (64 * 2 + 128 + sum(b.planes * 2 + b.planes * 4 * (2 if b.down else 1) for b in self.blocks))
        # BatchNorm batch statistics: every conv GEMM writes ops.STAT_PARTS partial rows of [2C] (one per CTA, fixed order)
        self.stats_buf = torch.zeros(n_stats * ops.STAT_PARTS, device=dev, dtype=f32)
        self._pack_stream = self._pack_pending = None    # side stream of refresh() and its not-yet-joined work
        self._wgq = None             # gemm.SplitKQueue of a running backward() (None: split-K second stages run immediately)
        self._wg_post = []
        self._wg_stream = None
        self.bn_keys = [k for k, _ in module.named_buffers() if k.startswith(prefix) and k.endswith("num_batches_tracked")
                        and "attnpool" not in k]
        self.nbt = [bufs[k] for k in self.bn_keys]

    # ------------------------------------------------------------------ derived weights
    def refresh(self):
        """Re-derive packed / padded bf16 operands from the fp32 masters (after an optimizer step or a load).
        Only the stem operands are derived on the calling stream (5 small pack kernels + ONE multi-tensor copy); the sixteen
        3x3 block weights are packed on a side stream next to the stem and joined by `forward` in front of layer 1
        (profiles/r2_step_timeline_kernels.txt: the serial form was 0.17 ms at the head of every step)."""
        from .engine import OVERLAP
        st, p = self.store, self.prefix
        main = torch.cuda.current_stream()
        if OVERLAP and os.environ.get("TRIS_PACK_SIDE", "1") != "0":
            if self._pack_stream is None:
                self._pack_stream = torch.cuda.Stream()
            side = self._pack_stream
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for blk in self.blocks:
                    ops.pack_conv(st.p(blk.p + "conv2.weight"), self.w3x3[blk.p])
            self._pack_pending = side
        else:
            for blk in self.blocks:
                ops.pack_conv(st.p(blk.p + "conv2.weight"), self.w3x3[blk.p])
        t27 = self._tmp27()
        ops.pack_conv(st.p(p + "conv1.weight"), t27)                    # [32, (r,s,c)] matches stem_im2col's k order
        ops.pack_conv(st.p(p + "conv2.weight"), self.w_stem2, co_pad=64, ci_pad=64)
        ops.pack_conv(st.p(p + "conv3.weight"), self.w_stem3, co_pad=64, ci_pad=64)
        ops.pack_conv_blockdiag(st.p(p + "conv2.weight"), self.w_pair2)
        ops.pack_conv_blockdiag(st.p(p + "conv3.weight"), self.w_pair3)
        dst, src = [self.w_stem1[:32, :27], self.w_pair1[:32, :27], self.w_pair1[32:, 32:59]], [t27, t27, t27]
        for k, pad in self.pad_bn.items():
            real = self.real_small_bn[k]
            dst += [pad.gamma[:32], pad.beta[:32], pad.rm[:32], pad.rv[:32]]
            src += [real.gamma, real.beta, real.rm, real.rv]
        for k, pr in self.pair_bn.items():
            real, c = self.pair_real[k], self.pair_real[k].gamma.numel()
            for d, r in ((pr.gamma, real.gamma), (pr.beta, real.beta), (pr.rm, real.rm), (pr.rv, real.rv)):
                dst += [d[:c], d[c:]]
                src += [r, r]
        self._copy_groups(dst, src)

    @staticmethod
    def _copy_groups(dst, src):
        """Multi-tensor copy (layout only), one launch per dtype group instead of one per tensor."""
        groups = {}
        for d, r in zip(dst, src):
            groups.setdefault((d.dtype, r.dtype, d.dim()), ([], []))
            groups[(d.dtype, r.dtype, d.dim())][0].append(d)
            groups[(d.dtype, r.dtype, d.dim())][1].append(r)
        for d, r in groups.values():
            torch._foreach_copy_(d, r)

    def _join_packs(self):
        if self._pack_pending is not None:
            torch.cuda.current_stream().wait_stream(self._pack_pending)
            self._pack_pending = None

    def _tmp27(self):
        if not hasattr(self, "_t27"):
            self._t27 = torch.empty((32, 27), device=self.store.device, dtype=bf16)
        return self._t27

    def _w1x1(self, key):
        w = self.store.s(key)
        return w.view(w.shape[0], w.shape[1])

    # ------------------------------------------------------------------ forward
    def forward(self, img: torch.Tensor, train: bool):
        """img fp32 NCHW [B,3,H,W] -> (c4 bf16 NHWC [B,H/32,W/32,2048], tape)."""
        p = self.prefix
        B, _, H, W = img.shape
        tape = {} if train else None
        so = [0]
        def stats(c):
            if not train:
                return None
            n = 2 * c * ops.STAT_PARTS
            s = self.stats_buf[so[0]: so[0] + n]       # every row is (re)written by the GEMM: no clearing needed
            so[0] += n
            return s

        pair = PAIR_STEM and B % 2 == 0
        if pair:
            x, stem_rec = self._stem_fwd_pair(img, train, stats)
        else:
            x, stem_rec = self._stem_fwd_padded(img, train, stats)
        if train:
            tape["stem"] = (pair,) + stem_rec
        self._join_packs()
        for blk in self.blocks:
            x, rec = self._block_fwd(blk, x, train, stats)
            if train:
                tape[blk.p] = rec
        if train:
            torch._foreach_add_(self.nbt, 1)
            self._sync_private_running_stats()
        return x, tape

    def _sync_private_running_stats(self):
        """The stem BatchNorms run on private padded / pair-packed copies of their parameters; after a train-mode forward
        the real running statistics (already updated) are mirrored into BOTH sets so that either stem path can follow."""
        p = self.prefix
        dst, src = [], []
        for nm in ("bn1", "bn2"):
            pad, real = self.pad_bn[p + nm], self.real_small_bn[p + nm]
            dst += [pad.rm[:32], pad.rv[:32]]; src += [real.rm, real.rv]
        for nm in ("bn1", "bn2", "bn3"):
            pr, real = self.pair_bn[p + nm], self.pair_real[p + nm]
            c = real.rm.numel()
            dst += [pr.rm[:c], pr.rm[c:], pr.rv[:c], pr.rv[c:]]; src += [real.rm, real.rm, real.rv, real.rv]
        self._copy_groups(dst, src)

    def _stem_fwd_padded(self, img, train, stats):
        """Odd batches: the two 32-channel activations are carried zero-padded to 64 channels."""
        p = self.prefix
        B, _, H, W = img.shape
        col = ops.stem_im2col(img.contiguous())
        s = stats(64)
        y1 = G.linear_fwd(col, self.w_stem1, stats=s).view(B, H // 2, W // 2, 64)
        a1 = ops.bn_apply(y1, s, self.pad_bn[p + "bn1"], train)
        s2 = stats(64)
        y2 = G.conv3x3_fwd(a1, self.w_stem2, stats=s2)
        a2 = ops.bn_apply(y2, s2, self.pad_bn[p + "bn2"], train)
        s3 = stats(64)
        y3 = G.conv3x3_fwd(a2, self.w_stem3, stats=s3)
        x = ops.bn_apply(y3, s3, self.bn[p + "bn3"], train, pool=2)
        if train:   # running stats of the padded copies back into the real buffers
            pads, reals = [self.pad_bn[p + nm] for nm in ("bn1", "bn2")], [self.real_small_bn[p + nm] for nm in ("bn1", "bn2")]
            self._copy_groups([t for r in reals for t in (r.rm, r.rv)], [t for q in pads for t in (q.rm[:32], q.rv[:32])])
        return x, (col, y1, a1, y2, a2, y3)

    def _stem_fwd_pair(self, img, train, stats):
        """Even batches: images (2i, 2i+1) share one row -- [B/2, h, w, 2*C] with block-diagonal weights.  The 32-channel
        tensors are then dense 64-channel rows (half the pixels of the padded form: half the MMA work for conv1 / conv2,
        half the BatchNorm traffic) and conv3 becomes a full 128-wide tile.  Channels c and c + C are one BatchNorm channel:
        their batch sums are folded (averaged, so that sum / (pixels of B/2 images) is the full-batch mean)."""
        p = self.prefix
        B, _, H, W = img.shape
        B2 = B // 2
        col = ops.stem_im2col_pair(img.contiguous())
        s = stats(64)
        with G.algo(27 * 32 / (64 * 32)):                  # block-diagonal [2 x (27 -> 32)] inside a 64 x 64 GEMM
            y1 = G.linear_fwd(col, self.w_pair1, stats=s).view(B2, H // 2, W // 2, 64)
        a1 = ops.bn_apply(y1, s, self.pair_bn[p + "bn1"], train, fold_half=32 if train else 0)
        s2 = stats(64)
        with G.algo(0.5):                                  # block-diagonal pair packing: half of the MACs are zeros
            y2 = G.conv3x3_fwd(a1, self.w_pair2, stats=s2)
        a2 = ops.bn_apply(y2, s2, self.pair_bn[p + "bn2"], train, fold_half=32 if train else 0)
        s3 = stats(128)
        with G.algo(0.5):
            y3 = G.conv3x3_fwd(a2, self.w_pair3, stats=s3)
        # [B/2, H/4, W/4, 128] pair-packed -> written un-paired as [B, H/4, W/4, 64] by the same kernel
        x = ops.bn_apply(y3, s3, self.pair_bn[p + "bn3"], train, pool=2, fold_half=64 if train else 0, unpair=True)
        if train:
            dst, src = [], []
            for nm in ("bn1", "bn2", "bn3"):
                pr, real = self.pair_bn[p + nm], self.pair_real[p + nm]
                c = real.rm.numel()
                dst += [real.rm, real.rv]; src += [pr.rm[:c], pr.rv[:c]]
            self._copy_groups(dst, src)
        return x, (col, y1, a1, y2, a2, y3)

    def _block_fwd(self, blk: _Block, x, train, stats):
        B, H, W, Cin = x.shape
        pl, q = blk.planes, blk.p
        s1 = stats(pl)
        y1 = G.linear_fwd(x.view(-1, Cin), self._w1x1(q + "conv1.weight"), stats=s1).view(B, H, W, pl)
        a1 = ops.bn_apply(y1, s1, self.bn[q + "bn1"], train)
        s2 = stats(pl)
        y2 = G.conv3x3_fwd(a1, self.w3x3[q], stats=s2)
        a2 = ops.bn_apply(y2, s2, self.bn[q + "bn2"], train, pool=blk.stride)
        Ho, Wo = H // blk.stride, W // blk.stride
        s3 = stats(4 * pl)
        y3 = G.linear_fwd(a2.view(-1, pl), self._w1x1(q + "conv3.weight"), stats=s3).view(B, Ho, Wo, 4 * pl)
        xp = yd = None
        # sign bits of the block output: the ReLU mask of the residual join for the backward pass (1/16 of the bytes of `out`)
        bits = torch.empty((B * Ho * Wo, pl // 2), device=x.device, dtype=torch.uint8) if train else None
        if blk.down:
            xp = ops.avgpool2(x) if blk.stride > 1 else x
            sd = stats(4 * pl)
            yd = G.linear_fwd(xp.view(-1, Cin), self._w1x1(q + "downsample.0.weight"), stats=sd).view(B, Ho, Wo, 4 * pl)
            out = ops.bn_apply(y3, s3, self.bn[q + "bn3"], train, y1=yd, stats1=sd, bn1=self.bn[q + "downsample.1"], bits=bits)
        else:
            out = ops.bn_apply(y3, s3, self.bn[q + "bn3"], train, residual=x, bits=bits)
        return out, ((x, y1, a1, y2, a2, y3, xp, yd, bits) if train else None)

    # ------------------------------------------------------------------ backward
    def backward(self, tape, dout: torch.Tensor):
        """dout: gradient w.r.t. c4 (bf16 NHWC).  Writes all parameter gradients into the store.

        Weight gradients are leaves of the backward chain: they are issued on a side stream (``_wg``) so that these
        tensor-core GEMMs run next to the HBM-bound BatchNorm-backward kernels of the following layer instead of in
        front of them.  The side stream joins the main stream at the end."""
        st, p = self.store, self.prefix
        self._wg_begin()
        try:
            self._backward(tape, dout)
        finally:
            self._wg_end()

    # ---- weight-gradient side stream
    def _wg_begin(self):
        from .engine import OVERLAP
        self._wg_stream = None
        self._wgq = G.SplitKQueue()      # deferred split-K second stages of this backward pass (one launch at the end)
        self._wg_post = []               # work that reads the reduced gradients (conv gradient un-packing)
        if OVERLAP:
            if not hasattr(self, "_wg_side"):
                self._wg_side = torch.cuda.Stream()
            self._wg_stream = self._wg_side
            self._wg_main = torch.cuda.current_stream()
            self._wg_stream.wait_stream(self._wg_main)
            self._wg_keep = []

    def _wg_end(self):
        def finish():
            self._wgq.flush()
            for fn in self._wg_post:
                fn()
            self._wg_post = []
        if self._wg_stream is not None:
            with torch.cuda.stream(self._wg_stream):
                finish()
            self._wg_main.wait_stream(self._wg_stream)
            self._wg_keep = []
            self._wg_stream = None
        else:
            finish()
        self._wgq = None

    def _post(self, fn):
        """Run fn after the deferred split-K reductions of this backward pass (immediately when nothing is deferred)."""
        if self._wgq is None:
            fn()
        else:
            self._wg_post.append(fn)

    def _wg(self, fn, *tensors):
        """Run fn() (a weight-gradient launch sequence reading `tensors`) on the side stream after everything issued so
        far on the main stream; the operands are kept alive until the join."""
        if getattr(self, "_wg_stream", None) is None:
            return fn()
        self._wg_stream.wait_stream(self._wg_main)
        for t in tensors:
            t.record_stream(self._wg_stream)
        self._wg_keep.extend(tensors)
        with torch.cuda.stream(self._wg_stream):
            fn()

    def _backward(self, tape, dout):
        ext = None      # partial sums of this block's bn3 backward, when the GEMM that produced `dout` already fused them
        for i in reversed(range(len(self.blocks))):
            blk = self.blocks[i]
            prev = self.blocks[i - 1] if i > 0 else None
            # this block's input-gradient GEMM can carry the bn3-backward reduction of the block in front of it when both
            # are plain residual blocks (the join is "+ identity": mask = sign of the shared activation)
            fuse_prev = FUSE_BN_BWD and prev is not None and not blk.down and not prev.down
            dout, ext = self._block_bwd(blk, tape[blk.p], dout, ext, prev if fuse_prev else None,
                                        tape[prev.p] if fuse_prev else None)
        self._stem_bwd(tape["stem"], dout)

    def _parts(self, c):
        """Partial-sum rows [STAT_PARTS, 2c] of a fused BatchNorm-backward reduction (+ 2c floats for the finalized sums)."""
        return torch.empty((ops.STAT_PARTS * 2 * c + 2 * c,), device=self.store.device, dtype=f32)

    def _bwd_stats(self, y, bn, mask=True):
        parts = self._parts(y.shape[-1])
        return parts, (parts, y, bn.mean, bn.scale if mask else None, bn.shift if mask else None)

    def _stem_bwd(self, stem_rec, dout):
        st, p = self.store, self.prefix
        pair, col, y1, a1, y2, a2, y3 = stem_rec
        if pair:
            return self._stem_bwd_pair(dout, col, y1, a1, y2, a2, y3)
        dy3, _, _ = ops.bn_bwd(dout, None, y3, self.bn[p + "bn3"], pool=2)
        self._wgrad3x3(dy3, a2, p + "conv3.weight", ci_pad=64)
        da2 = G.conv3x3_dgrad(dy3, self.w_stem3, 64)
        pb2 = self.pad_bn[p + "bn2"]
        pb2.dgamma.zero_(); pb2.dbeta.zero_()
        dy2, _, _ = ops.bn_bwd(da2, None, y2, pb2)
        self._wgrad3x3(dy2, a1, p + "conv2.weight", ci_pad=64)
        da1 = G.conv3x3_dgrad(dy2, self.w_stem2, 64)
        pb1 = self.pad_bn[p + "bn1"]
        pb1.dgamma.zero_(); pb1.dbeta.zero_()
        dy1, _, _ = ops.bn_bwd(da1, None, y1, pb1)

        def stem1():
            gw = G.linear_wgrad(dy1.view(-1, 64), col, queue=self._wgq)   # [64, 64] fp32
            g1 = st.g(p + "conv1.weight")                                # [32,3,3,3]
            self._post(lambda: g1.add_(gw[:32, :27].reshape(32, 3, 3, 3).permute(0, 3, 1, 2)))
        self._wg(stem1, dy1, col)
        for nm, pad in (("bn1", pb1), ("bn2", pb2)):
            st.g(p + nm + ".weight").add_(pad.dgamma[:32])
            st.g(p + nm + ".bias").add_(pad.dbeta[:32])

    def _stem_bwd_pair(self, dout, col, y1, a1, y2, a2, y3):
        st, p = self.store, self.prefix
        B, h4, w4, _ = dout.shape
        B2 = B // 2
        torch._foreach_zero_([t for nm in ("bn1", "bn2", "bn3") for t in (self.pair_bn[p + nm].dgamma, self.pair_bn[p + nm].dbeta)])
        dy3, _, _ = ops.bn_bwd(dout.contiguous(), None, y3, self.pair_bn[p + "bn3"], pool=2, fold_half=64, unpair=True)   # re-pairs on read
        self._wgrad3x3_pair(dy3, a2, p + "conv3.weight")
        ext2, bs2 = self._bwd_stats(y2, self.pair_bn[p + "bn2"]) if FUSE_BN_BWD else (None, None)
        with G.algo(0.5):
            da2 = G.conv3x3_dgrad(dy3, self.w_pair3, 64, bwd_stats=bs2)
        dy2, _, _ = ops.bn_bwd(da2, None, y2, self.pair_bn[p + "bn2"], fold_half=32, ext=ext2)
        self._wgrad3x3_pair(dy2, a1, p + "conv2.weight")
        ext1, bs1 = self._bwd_stats(y1, self.pair_bn[p + "bn1"]) if FUSE_BN_BWD else (None, None)
        with G.algo(0.5):
            da1 = G.conv3x3_dgrad(dy2, self.w_pair2, 64, bwd_stats=bs1)
        dy1, _, _ = ops.bn_bwd(da1, None, y1, self.pair_bn[p + "bn1"], fold_half=32, ext=ext1)

        def stem1():
            with G.algo(27 * 32 / (64 * 32)):
                gw = G.linear_wgrad(dy1.view(-1, 64), col, queue=self._wgq)   # [64, 64] fp32, two diagonal 32 x 27 blocks
            g1 = st.g(p + "conv1.weight")
            self._post(lambda: g1.add_((gw[:32, :27] + gw[32:, 32:59]).reshape(32, 3, 3, 3).permute(0, 3, 1, 2)))
        self._wg(stem1, dy1, col)
        for nm in ("bn1", "bn2", "bn3"):     # folded sums hold the pair AVERAGE: the full-batch gradient is twice that
            pr = self.pair_bn[p + nm]
            c = pr.dgamma.numel() // 2
            st.g(p + nm + ".weight").add_(pr.dgamma[:c], alpha=2.0)
            st.g(p + nm + ".bias").add_(pr.dbeta[:c], alpha=2.0)

    def _wgrad3x3_pair(self, dy, x, key):
        gw = self.store.g(key)

        def run():
            with G.algo(0.5):
                gp = G.conv3x3_wgrad(dy, x, queue=self._wgq)
            self._post(lambda: ops.unpack_conv_grad_blockdiag(gp, gw, 2))
        self._wg(run, dy, x)

    def _wgrad3x3(self, dy, x, key, ci_pad=None):
        gw = self.store.g(key)

        def run():
            gp = G.conv3x3_wgrad(dy, x, queue=self._wgq)
            self._post(lambda: ops.unpack_conv_grad(gp, gw, ci_pad=ci_pad))
        self._wg(run, dy, x)

    def _wgrad1x1(self, dy2d, x2d, key):
        gw = self.store.g(key)
        self._wg(lambda: G.linear_wgrad(dy2d, x2d, out=gw.view(gw.shape[0], gw.shape[1]), accumulate=True, queue=self._wgq),
                 dy2d, x2d)

    def _block_bwd(self, blk: _Block, rec, dout, ext=None, prev=None, prev_rec=None):
        """-> (dx, ext_prev).  ext: fused partial sums of this block's bn3 (then `dout` is the masked gradient g).
        prev / prev_rec: the block in front, when this block's input-gradient GEMM is to fuse ITS bn3 reduction."""
        x, y1, a1, y2, a2, y3, xp, yd, bits = rec
        q, pl = blk.p, blk.planes
        Cin = x.shape[3]
        g_bits = None          # non-None: the identity-branch gradient is `dout` masked by these bits (never materialised)
        if blk.down:
            dy3, dyd, g = ops.bn_bwd(dout, None, y3, self.bn[q + "bn3"], y1=yd, bn1=self.bn[q + "downsample.1"], bits=bits)
        elif ext is not None:
            dy3, dyd, g = ops.bn_bwd(dout, None, y3, self.bn[q + "bn3"], ext=ext)[0], None, dout
        else:
            dy3, dyd, g = ops.bn_bwd(dout, None, y3, self.bn[q + "bn3"], bits=bits)[0], None, dout
            g_bits = bits
        self._wgrad1x1(dy3.view(-1, 4 * pl), a2.view(-1, pl), q + "conv3.weight")
        ext2, bs2 = self._bwd_stats(y2, self.bn[q + "bn2"]) if (FUSE_BN_BWD and blk.stride == 1) else (None, None)
        da2 = G.linear_dgrad(dy3.view(-1, 4 * pl), self._w1x1(q + "conv3.weight"), bwd_stats=bs2).view(a2.shape)
        dy2, _, _ = ops.bn_bwd(da2, None, y2, self.bn[q + "bn2"], pool=blk.stride, ext=ext2)
        self._wgrad3x3(dy2, a1, q + "conv2.weight")
        ext1, bs1 = self._bwd_stats(y1, self.bn[q + "bn1"]) if FUSE_BN_BWD else (None, None)
        da1 = G.conv3x3_dgrad(dy2, self.w3x3[q], pl, bwd_stats=bs1)
        dy1, _, _ = ops.bn_bwd(da1, None, y1, self.bn[q + "bn1"], ext=ext1)
        self._wgrad1x1(dy1.view(-1, pl), x.view(-1, Cin), q + "conv1.weight")
        w1 = self._w1x1(q + "conv1.weight")
        if blk.down:
            self._wgrad1x1(dyd.view(-1, 4 * pl), xp.view(-1, Cin), q + "downsample.0.weight")
            dxp = G.linear_dgrad(dyd.view(-1, 4 * pl), self._w1x1(q + "downsample.0.weight"))
            if blk.stride > 1:
                dx = G.linear_dgrad(dy1.view(-1, pl), w1).view(x.shape)
                dx = ops.avgpool2_bwd(dxp.view(xp.shape), add=dx)
            else:
                dx = G.linear_dgrad(dy1.view(-1, pl), w1, residual=dxp).view(x.shape)
            return dx, None
        if prev is not None:
            # x is the previous block's output relu(bn3(y3') + identity'): this GEMM stores g' = (dy1 W1 + g) * [x > 0] and the
            # partial sums (sum g', sum g' (y3' - mean')) of the previous block's bn3 backward
            extp, bsp = self._bwd_stats(prev_rec[5], self.bn[prev.p + "bn3"], mask=False)
            dx = G.linear_dgrad(dy1.view(-1, pl), w1, residual=g.view(-1, Cin), dact_src=x.view(-1, Cin), act=ops.L.ACT_RELU,
                                bwd_stats=bsp, res_bits=g_bits).view(x.shape)
            return dx, extp
        dx = G.linear_dgrad(dy1.view(-1, pl), w1, residual=g.view(-1, Cin), res_bits=g_bits).view(x.shape)
        return dx, None
