"""Python front-end of the tcgen05 GEMM / implicit-conv kernel (tris_gemm in include/tris_sm100.h).

All tensors are CUDA, contiguous; activations / weights bf16, bias / stats / weight gradients fp32.
Layouts: linear x [M,K], w [N,K] (nn.Linear layout); conv activations NHWC, 3x3 weights packed [Cout, 9*Cin]
with k = (r*3+s)*Cin + ci  (``pack_conv3x3``).
"""
from __future__ import annotations

import os

import torch

from . import _lib as L

SMS = 148


# measured cost of one 64-deep k-block of a 128 x BN tile on one SM (us): the L2 -> smem feed ((128 + BN) * 128 B at ~120 GB/s
# per SM) for narrow tiles, the MMA itself (128 * BN * 64 MACs at 9.4 TFLOP/s per SM) for BN = 256 (tools/gemm_trace.py)
_KBLOCK_US = {32: 0.17, 64: 0.20, 128: 0.27, 256: 0.45}


# measured: isolated short-K GEMMs gain 8-20 % with 128-wide tiles, the whole step LOSES 0.2 ms (more A re-reads) -> off
SHORTK_BN128 = os.environ.get("TRIS_BN_SHORTK", "0") == "1"


def _bn_for(n: int, tiles_m: int, mn_major_b: bool = False, k: int = 0) -> int:
    """UMMA N per tile: the width that minimises  waves(148 SMs) x (time of one k-block at that width).
    Short contractions (k <= 128: one or two k-blocks per tile) are bound by the epilogue, not the mainloop: 128-wide
    tiles leave room for two staging buffers, so that the two epilogue groups can overlap consecutive tiles."""
    if n <= 32 and not mn_major_b:
        return 32
    if SHORTK_BN128 and 0 < k <= 128 and n >= 128 and tiles_m * ((n + 127) // 128) > 2 * SMS:
        return 128
    n_pad = ((n + 63) // 64) * 64
    best, best_cost = 64, None
    for bn in (64, 128, 256):
        if bn > n_pad and bn != 64:
            continue
        tiles = tiles_m * ((n + bn - 1) // bn)
        waves = (tiles + SMS - 1) // SMS
        cost = waves * _KBLOCK_US[bn]
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = bn, cost
    return best


WGRAD_MAX_CTAS = int(os.environ.get("TRIS_WGRAD_CTAS", "0"))   # experiment: cap the SMs a side-stream weight gradient takes


# Share of a GEMM call's 2*M*N*K that is algorithmic work (bench.py's roofline counts un-padded FLOPs): the stem convs
# run with zero-padded / block-diagonal operands (27 -> 64 im2col columns, image-pair packing), see resnet.py.
algo_share = 1.0


class algo(object):
    """with algo(0.5): ... -- tags the GEMM calls inside with their algorithmic share (read by bench.py only)."""

    def __init__(self, share):
        self.share = share

    def __enter__(self):
        global algo_share
        self.prev, algo_share = algo_share, self.share

    def __exit__(self, *a):
        global algo_share
        algo_share = self.prev


STAT_PARTS = 148   # rows of a partial-statistics buffer (ops.STAT_PARTS): one per CTA of the persistent kernel


# bench.py's kernel timing: when set to [uint64 tensor [n, 2], next slot, records], every GEMM launch gets a time-stamp slot
# (the kernel writes its device-side start / end there) and a record (algorithmic flops, issued flops, bytes)
timing = None


def _desc(**kw) -> L.GemmDesc:
    d = L.GemmDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    if d.stats:
        d.stats_parts = STAT_PARTS
    if timing is not None and timing[1] < timing[0].shape[0]:
        d.tstamp = timing[0].data_ptr() + 16 * timing[1]
        timing[1] += 1
        taps = d.taps if (d.wgrad and d.taps > 1) else 1
        nb = max(1, d.batch)
        full = 2.0 * d.M * d.N * d.K * taps * nb
        esz = 4 if d.out_dtype == L.DT_F32 else 2
        if d.a_mode == L.OP_CONV and not d.wgrad:       # activation read once (not once per tap) + weights + output
            by = 2.0 * d.M * (d.K // max(1, d.taps)) + 2.0 * d.N * d.K + esz * d.M * d.N
        else:
            by = 2.0 * d.M * d.K * (nb if d.a_batch_stride else 1) + 2.0 * d.N * d.K * (nb if d.b_batch_stride else 1) \
                + esz * d.M * d.N * taps * nb
        timing[2].append((full * algo_share, full, by))
    return d


def _bwd_stats(kw, bwd_stats):
    """bwd_stats = (parts fp32 [STAT_PARTS*2N (+spare)], y bf16 (shape of the output), mean fp32 [N], mask_sc | None, mask_sh | None):
    the epilogue stores g = out * [y*mask_sc + mask_sh > 0] and accumulates the BatchNorm-backward sums (sum g, sum g (y - mean))
    of the stored g as partial rows -- the dgamma / dbeta reduction of the layer that produced y, fused into this GEMM."""
    if bwd_stats is None:
        return kw
    parts, y, mean, msc, msh = bwd_stats
    assert y.dtype == torch.bfloat16 and y.is_contiguous()
    kw.update(stats=L.ptr(parts), stats_mode=1, stats_y=L.ptr(y), stats_mu=L.ptr(mean), mask_sc=L.ptr(msc), mask_sh=L.ptr(msh))
    return kw


class SplitKQueue:
    """Pending second stages of split-K weight gradients (partials already stored to their workspaces by the GEMMs).
    ``flush()`` reduces all of them -- in split order, bit-reproducible -- with ONE launch per 64 tensors, on the current
    stream: call it on the stream the GEMMs ran on, before anything reads the gradients."""

    def __init__(self):
        self.items = []          # (ws tensor, out tensor, rows, w, ldd, split, accumulate)

    def push(self, ws, out, rows, w, ldd, split, accumulate):
        self.items.append((ws, out, rows, w, ldd, split, accumulate))

    def flush(self):
        if not self.items:
            return
        n = len(self.items)
        arr = (L.ReduceItem * n)()
        for i, (ws, out, rows, w, ldd, split, acc) in enumerate(self.items):
            arr[i] = L.ReduceItem(ws.data_ptr(), out.data_ptr(), rows, w, ldd, split, int(acc), 0)
        L.call("tris_splitk_reduce_multi", arr, n, launches=(n + 63) // 64)
        self.items = []          # the caching allocator keeps the workspaces valid for the kernels already queued on this stream


def _launch_wgrad(d, ws, out, rows, w, ldd, accumulate, queue):
    """Run a (possibly split-K) weight-gradient GEMM; with a queue, its second stage is deferred to queue.flush()."""
    split = d.split_k
    if queue is not None and split > 1:
        d.defer_reduce = 1
        L.gemm_raw(d)
        if d.split_used > 1:
            queue.push(ws, out, rows, w, ldd, d.split_used, accumulate)
    else:
        L.gemm_raw(d, launches=2 if split > 1 else 1)


def _ws(split, m, w, dev):
    """fp32 workspace of a split-K weight gradient: [split][m][w]."""
    return torch.empty((split * m * w,), device=dev, dtype=torch.float32) if split > 1 else None


def _chk(t, dtype, name):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), f"{name}: need contiguous CUDA {dtype}, got {t.dtype} {t.shape}"


def linear_fwd(x, w, bias=None, act=L.ACT_NONE, residual=None, out=None, out_dtype=torch.bfloat16, stats=None,
               block_n=None, d_pre=None):
    """out[M,N] = act(x[M,K] @ w[N,K]^T + bias) + residual ; optional column stats (sum, sumsq) into stats[2N]."""
    _chk(x, torch.bfloat16, "x"); _chk(w, torch.bfloat16, "w")
    m, k = x.shape
    n = w.shape[0]
    assert w.shape[1] == k
    res_f32 = residual is not None and residual.dtype == torch.float32     # fp32 residual stream: fp32 in, fp32 out
    if out is None:
        out = torch.empty((m, n), device=x.device, dtype=torch.float32 if res_f32 else out_dtype)
    assert not res_f32 or out.dtype == torch.float32
    tiles_m = (m + 127) // 128
    bn = block_n or _bn_for(n, tiles_m, k=k)
    if out.dtype == torch.float32:
        bn = min(bn, 128)
    d = _desc(a=L.ptr(x), b=L.ptr(w), d=L.ptr(out), bias=L.ptr(bias), residual=L.ptr(residual), stats=L.ptr(stats),
              a_mode=L.OP_K2D, b_mode=L.OP_K2D, M=m, N=n, K=k, lda=k, ldb=k, ldd=n, taps=1, block_n=bn, split_k=1,
              act=act, out_dtype=L.DT_F32 if out.dtype == torch.float32 else L.DT_BF16, d_pre=L.ptr(d_pre),
              residual_f32=int(res_f32))
    L.gemm_raw(d)
    return out


def linear_in_fwd(x, w, batch, gamma, beta, relu, mix_scale=1.0, mix_add=None, eps=1e-5, block_n=128):
    """Fused 1x1 conv + InstanceNorm2d(affine) [+ ReLU] [* mix_scale + mix_add] over the P = rows / batch pixels of each image
    (v_proj / v_output of model/attn.py:75-105): one batched GEMM whose tiles are image-aligned (P <= 128 rows x block_n
    channels), so the per-(image, channel) statistics are complete inside a tile and the normalisation happens in the epilogue.
    -> (y raw bf16 [rows, N] (saved for backward), y_norm bf16 [rows, N], mean f32 [batch, N], invstd f32 [batch, N])."""
    _chk(x, torch.bfloat16, "x"); _chk(w, torch.bfloat16, "w")
    rows, k = x.shape
    n = w.shape[0]
    P = rows // batch
    assert rows == batch * P and P <= 128 and w.shape[1] == k and batch >= 2 and n % 8 == 0
    y = torch.empty((rows, n), device=x.device, dtype=torch.bfloat16)
    yn = torch.empty_like(y)
    mean = torch.empty((batch, n), device=x.device, dtype=torch.float32)
    invstd = torch.empty_like(mean)
    d = _desc(a=L.ptr(x), b=L.ptr(w), d=L.ptr(y), d_norm=L.ptr(yn), a_mode=L.OP_K2D, b_mode=L.OP_K2D, M=P, N=n, K=k, lda=k, ldb=k, ldd=n,
              taps=1, block_n=block_n, split_k=1, out_dtype=L.DT_BF16, batch=batch, a_batch_stride=P * k, b_batch_stride=0,
              d_batch_stride=P * n, in_gamma=L.ptr(gamma), in_beta=L.ptr(beta), in_mean=L.ptr(mean), in_invstd=L.ptr(invstd),
              in_add=L.ptr(mix_add), in_eps=eps, in_mix=mix_scale, in_relu=int(relu))
    L.gemm_raw(d)
    return y, yn, mean, invstd


def linear_dgrad(dy, w, out=None, out_dtype=torch.bfloat16, residual=None, block_n=None, dact_src=None,
                 act=L.ACT_NONE, bias=None, bwd_stats=None, res_bits=None):
    """dx[M,K] = dy[M,N] @ w[N,K]   (w read MN-major: no transposed weight copy)."""
    _chk(dy, torch.bfloat16, "dy"); _chk(w, torch.bfloat16, "w")
    m, n = dy.shape
    k = w.shape[1]
    assert w.shape[0] == n
    if out is None:
        out = torch.empty((m, k), device=dy.device, dtype=out_dtype)
    tiles_m = (m + 127) // 128
    bn = block_n or _bn_for(k, tiles_m, True, k=n)
    if out.dtype == torch.float32:
        bn = min(bn, 128)
    d = _desc(**_bwd_stats(dict(a=L.ptr(dy), b=L.ptr(w), d=L.ptr(out), residual=L.ptr(residual), bias=L.ptr(bias), a_mode=L.OP_K2D,
              b_mode=L.OP_MN2D, M=m, N=k, K=n, lda=n, ldb=k, ldd=k, taps=1, block_n=bn, split_k=1, act=act,
              dact_src=L.ptr(dact_src), out_dtype=L.DT_F32 if out.dtype == torch.float32 else L.DT_BF16,
              res_bits=L.ptr(res_bits)), bwd_stats))
    L.gemm_raw(d)
    return out


def _split_for(tiles: int, kblocks: int) -> int:
    """Split-K factor for a weight-gradient GEMM with `tiles` output tiles and `kblocks` k-blocks: minimise
    waves(148 SMs) x k-blocks per work unit (+ a small per-unit cost for prologue / fp32 reduce-add epilogue)."""
    if tiles >= 2 * SMS:
        return 1
    best, best_cost = 1, None
    for sk in range(1, max(1, min(kblocks // 2, 4 * SMS // max(1, tiles) + 1)) + 1):
        per = (kblocks + sk - 1) // sk
        units = tiles * ((kblocks + per - 1) // per)
        waves = (units + SMS - 1) // SMS
        cost = waves * (per + 6)
        if best_cost is None or cost < best_cost:
            best, best_cost = sk, cost
    return best


def linear_wgrad(dy, x, out=None, accumulate=False, block_n=None, split_k=None, queue=None):
    """dw[N,K] (f32) (+)= dy[M,N]^T @ x[M,K]  (both operands MN-major, contraction over rows, split-K)."""
    _chk(dy, torch.bfloat16, "dy"); _chk(x, torch.bfloat16, "x")
    m, n = dy.shape
    k = x.shape[1]
    assert x.shape[0] == m
    tiles_m = (n + 127) // 128
    bn = min(block_n or _bn_for(k, tiles_m, True), 128)
    tiles = tiles_m * ((k + bn - 1) // bn)
    kblocks = (m + 63) // 64
    sk = split_k or _split_for(tiles, kblocks)
    sk = max(1, min(sk, kblocks))
    if out is None:
        out = (torch.zeros if accumulate else torch.empty)((n, k), device=dy.device, dtype=torch.float32)
    _chk(out, torch.float32, "out")
    ws = _ws(sk, n, (k + 3) // 4 * 4, dy.device)
    d = _desc(a=L.ptr(dy), b=L.ptr(x), d=L.ptr(out), a_mode=L.OP_MN2D, b_mode=L.OP_MN2D, M=n, N=k, K=m, lda=n, ldb=k,
              ldd=k, taps=1, block_n=bn, split_k=sk, out_dtype=L.DT_F32, atomic=1 if accumulate else 0, max_ctas=WGRAD_MAX_CTAS,
              splitk_ws=L.ptr(ws))
    _launch_wgrad(d, ws, out, n, (k + 3) // 4 * 4, k, accumulate, queue)
    return out


# --------------------------------------------------------------------------------------- 3x3 convolution
def pack_conv3x3(w_oihw: torch.Tensor) -> torch.Tensor:
    """OIHW fp32/bf16 -> bf16 [Cout, 9*Cin], k = (r*3+s)*Cin + ci."""
    co, ci, kh, kw = w_oihw.shape
    return w_oihw.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(torch.bfloat16).contiguous()


def unpack_conv3x3_grad(dwp: torch.Tensor, cin: int) -> torch.Tensor:
    co = dwp.shape[0]
    return dwp.reshape(co, 3, 3, cin).permute(0, 3, 1, 2).contiguous()


def conv_tile(h: int, w: int, max_rows: int = 128, mult: int = 1):
    """Spatial patch (th, tw) for one tile: maximise useful rows / covered rows."""
    best, best_eff = None, -1.0
    for tw in range(1, min(w, max_rows) + 1):
        for th in range(1, min(h, max_rows // tw) + 1):
            rows = th * tw
            if rows % mult:
                continue
            cover = ((h + th - 1) // th) * th * ((w + tw - 1) // tw) * tw
            eff = (h * w / cover) * (rows / max_rows)
            if eff > best_eff + 1e-9 or (abs(eff - best_eff) < 1e-9 and tw > best[1]):
                best, best_eff = (th, tw), eff
    return best


CONV_HALO = os.environ.get("TRIS_CONV_HALO", "1") != "0"    # halo re-use for 3x3 convs whose images tile into 16 x 8 patches


def _halo_ok(h, w, taps):
    return CONV_HALO and taps == 9 and h % 16 == 0 and w % 8 == 0


def conv3x3_fwd(x, wp, stats=None, out=None, block_n=None, taps=9):
    """y[n,h,w,co] = conv3x3(x[n,h,w,ci], pad 1, stride 1); wp packed [co, taps*ci]. taps=1 -> 1x1 conv via TMA-4D."""
    _chk(x, torch.bfloat16, "x"); _chk(wp, torch.bfloat16, "wp")
    n, h, w, ci = x.shape
    co = wp.shape[0]
    assert wp.shape[1] == taps * ci and ci % 64 == 0
    if out is None:
        out = torch.empty((n, h, w, co), device=x.device, dtype=torch.bfloat16)
    halo = _halo_ok(h, w, taps)
    th, tw = (16, 8) if halo else conv_tile(h, w)
    tiles_m = n * ((h + th - 1) // th) * ((w + tw - 1) // tw)
    bn = block_n or _bn_for(co, tiles_m)
    d = _desc(a=L.ptr(x), b=L.ptr(wp), d=L.ptr(out), stats=L.ptr(stats), a_mode=L.OP_CONV, b_mode=L.OP_K2D,
              M=n * h * w, N=co, K=taps * ci, ldb=taps * ci, ldd=co, img_n=n, img_h=h, img_w=w, tile_h=th, tile_w=tw,
              taps=taps, block_n=bn, split_k=1, out_dtype=L.DT_BF16, conv_halo=int(halo))
    L.gemm_raw(d)
    return out


def conv3x3_dgrad(dy, wp, cin, out=None, block_n=None, bwd_stats=None):
    """dx[n,h,w,ci] = conv_transpose3x3(dy[n,h,w,co]); wp is the forward packing [co, 9*ci] read MN-major."""
    _chk(dy, torch.bfloat16, "dy"); _chk(wp, torch.bfloat16, "wp")
    n, h, w, co = dy.shape
    assert wp.shape == (co, 9 * cin) and co % 64 == 0
    if out is None:
        out = torch.empty((n, h, w, cin), device=dy.device, dtype=torch.bfloat16)
    halo = _halo_ok(h, w, 9) and bwd_stats is None
    th, tw = (16, 8) if halo else conv_tile(h, w)
    tiles_m = n * ((h + th - 1) // th) * ((w + tw - 1) // tw)
    bn = block_n or _bn_for(cin, tiles_m, True)
    d = _desc(**_bwd_stats(dict(a=L.ptr(dy), b=L.ptr(wp), d=L.ptr(out), a_mode=L.OP_CONV, b_mode=L.OP_MN2D, M=n * h * w, N=cin,
              K=9 * co, ldb=9 * cin, ldd=cin, img_n=n, img_h=h, img_w=w, tile_h=th, tile_w=tw, taps=9, flip=1,
              b_tap_stride=cin, block_n=bn, split_k=1, out_dtype=L.DT_BF16, conv_halo=int(halo)), bwd_stats))
    L.gemm_raw(d)
    return out


def conv3x3_wgrad(dy, x, out=None, block_n=None, split_k=None, queue=None):
    """dwp[co, 9*ci] (f32) = sum_pixels dy[.,co] * x[. + tap, ci]."""
    _chk(dy, torch.bfloat16, "dy"); _chk(x, torch.bfloat16, "x")
    n, h, w, co = dy.shape
    ci = x.shape[3]
    assert x.shape[:3] == dy.shape[:3]
    th, tw = conv_tile(h, w, max_rows=96, mult=16)
    kblocks = n * ((h + th - 1) // th) * ((w + tw - 1) // tw)
    tiles_m = (co + 127) // 128
    bn = min(block_n or _bn_for(ci, 9 * tiles_m, True), 128)
    tiles = 9 * tiles_m * ((ci + bn - 1) // bn)
    sk = max(1, min(split_k or _split_for(tiles, kblocks), kblocks))
    if out is None:
        out = torch.empty((co, 9 * ci), device=dy.device, dtype=torch.float32)
    if sk == 1:
        out.zero_()    # a single partial per tile leaves through TMA reduce-add (one add per element: order-free)
    ws = _ws(sk, co, 9 * ci, dy.device)
    d = _desc(a=L.ptr(dy), b=L.ptr(x), d=L.ptr(out), a_mode=L.OP_CONV, b_mode=L.OP_CONV, M=co, N=ci, K=n * h * w,
              ldd=9 * ci, img_n=n, img_h=h, img_w=w, tile_h=th, tile_w=tw, taps=9, wgrad=1, block_n=bn, split_k=sk,
              out_dtype=L.DT_F32, atomic=0, max_ctas=WGRAD_MAX_CTAS, splitk_ws=L.ptr(ws))
    _launch_wgrad(d, ws, out, co, 9 * ci, 9 * ci, False, queue)
    return out


# --------------------------------------------------------------------------------------- generic / batched entry
def gemm_ex(a, b, out, M, N, K, a_mode=L.OP_K2D, b_mode=L.OP_K2D, lda=None, ldb=None, ldd=None, batch=1, a_bs=0, b_bs=0,
            d_bs=0, bias=None, act=L.ACT_NONE, atomic=0, split_k=1, block_n=None, scale=0.0, residual=None):
    """out[b] (M x N, row stride ldd) = act(A[b] . B[b]^T-ish + bias) on explicit extents / strides, so strided views
    (e.g. the Q/K/V thirds of a fused projection) and per-image batches can be used without copies.
    a_mode K2D: A[b] is [M, K] rows (stride lda);  MN2D: A[b] is [K, M] rows.   b likewise with N."""
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_cuda and b.is_cuda
    lda = lda if lda is not None else (K if a_mode == L.OP_K2D else M)
    ldb = ldb if ldb is not None else (K if b_mode == L.OP_K2D else N)
    ldd = ldd if ldd is not None else N
    f32 = out.dtype == torch.float32
    tiles_m = ((M + 127) // 128) * batch
    bn = block_n or _bn_for(N, tiles_m, b_mode != L.OP_K2D)
    if f32:
        bn = min(bn, 128)
    kblocks = (K + 63) // 64
    split_k = max(1, min(split_k, kblocks))
    ws = _ws(split_k, M, (N + 3) // 4 * 4, a.device)
    d = _desc(a=L.ptr(a), b=L.ptr(b), d=L.ptr(out), bias=L.ptr(bias), a_mode=a_mode, b_mode=b_mode, M=M, N=N, K=K, lda=lda,
              ldb=ldb, ldd=ldd, taps=1, block_n=bn, split_k=split_k, act=act, out_dtype=L.DT_F32 if f32 else L.DT_BF16,
              atomic=atomic, batch=batch, a_batch_stride=a_bs, b_batch_stride=b_bs, d_batch_stride=d_bs, scale=scale,
              residual=L.ptr(residual), splitk_ws=L.ptr(ws))
    L.gemm_raw(d, launches=2 if split_k > 1 else 1)
    return out
