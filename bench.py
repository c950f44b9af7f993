#!/usr/bin/env python
"""Stage-1 training throughput benchmark (BASELINE.json metric: samples/sec at 320x320, len 20, bs 48 per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 3 --warmup 1        # CPU arm: the unmodified reference on the host cores
    python bench.py --impl reference-gpu --ref-mode bf16          # the unmodified reference in PyTorch eager on cuda:0

One "step" = forward + three losses + backward + gradient all-reduce + AdamW on one synthetic batch of 48 samples per
GPU.  `value` times steps whose inputs are already in HBM; `e2e` times the same step through the public trainer API
with pinned HOST batches (H2D copy of img/ids/negatives and a D2H read of the loss inside the timed region).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import warnings

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_SAMPLE = 98.0        # algorithmic fwd+bwd work of the necessary step (SURVEY 8d / BASELINE.md)
METRIC = "Stage-1 training samples/sec (320x320, len20, bs48)"
# cross-modal attention (K7) + vis/lan projection (K6), forward, per step of 48: 47.75 + 20.2 GFLOP;
# compulsory bf16 HBM bytes 41.3 MB (K7) + c4 19.7 MB + W_v 4.2 MB
WORKLOAD = ("Stage-1 train step (configs[1]): CLIP-RN50 + text tower + cross-modal fusion + aux ViT-B/32 losses, 320x320, "
            "len 20, 3 negatives")
K7_GFLOP_PER_STEP_B48 = 47.75 + 20.2
K7_MBYTES_PER_STEP_B48 = 41.3 + 19.7 + 4.2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return d.get("bf16_tflops_sustained", 1400.0), d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        clocks, reasons = [], set()
        for r in rows:
            try:
                clocks.append(float(r[0]))
                out["sm_max_mhz"] = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if clocks:
            out["sm_mhz"] = statistics.median(clocks)
        out["reasons"] = sorted(reasons)
        return out


def make_args():
    return argparse.Namespace(synthetic_weights=True, bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)


# ------------------------------------------------------------------------------------------------ CPU arm
def _port_step_factory(batch):
    """Oracle PORT of the reference step (used only when no copy of the reference sources is available)."""
    from oracle import tris_oracle as O
    from oracle import weights as W
    sd = W.make_tris_state_dict(0)
    aux = W.make_vitb32_clip_state_dict(7, cos_bias=True)
    keys = O.trainable_keys(sd)
    m = {k: torch.zeros_like(sd[k]) for k in keys}
    v = {k: torch.zeros_like(sd[k]) for k in keys}
    it = [0]

    def step(b):
        img, ids, negs = W.synthetic_batch(b, 320, 20, 3, 1234 + it[0])
        losses, grads, new_stats, _ = O.train_step(sd, aux, img, ids, negs)
        for k in keys:
            grp = O.param_group_of(k)
            if grp < 0:
                continue
            lr = 5e-5 * (0.1 if grp == 0 else 1.0) * O.poly_lr(it[0], 1000)
            sd[k], m[k], v[k] = O.adamw_step(sd[k], grads[k], m[k], v[k], it[0] + 1, lr)
        sd.update(new_stats)
        it[0] += 1
        return float(losses["loss"])
    return step, "port"


def _reference_step_factory():
    """The UNMODIFIED reference modules (baseline/_ref or /root/reference) on CPU fp32: TRIS + aux ViT-B/32 +
    torch.optim.AdamW + LambdaLR, loop body of train_stage1.py:320-372."""
    from baseline import ref_step as RS
    from tris_b200.synthetic import synthetic_batch
    ns, args = RS.load(batch=48)
    model, aux = RS.build_models(ns, args, "cpu")
    model.train()
    opt, sched = RS.make_optimizer(model, args, 100000)
    it = [0]

    def step(b):
        img, ids, neg = synthetic_batch(b, 320, 20, 3, seed=1234 + it[0])
        it[0] += 1
        return float(RS.cpu_train_step(ns, args, model, aux, opt, sched, img, ids.long(), neg.long())["loss"])
    return step, "reference"


def cpu_arm(steps, warmup, budget_s, threads):
    """Times the reference's CPU implementation of the step on the host cores.  Each step is a bounded sample of the
    bs48 workload: the largest batch in {48, 24, 16, 8, 4} for which (steps + warmup) steps fit in `budget_s`.
    -> dict(value samples/s, sec per step, batch, kind, loss)."""
    torch.set_num_threads(threads)
    try:
        from baseline import ref_loader
        have_ref = ref_loader.available()
    except Exception:
        have_ref = False
    step, kind = _reference_step_factory() if have_ref else _port_step_factory(4)
    t0 = time.perf_counter()
    step(4)                                  # probe (also the first warm-up: allocator, oneDNN primitive caches)
    per_sample = (time.perf_counter() - t0) / 4
    batch = 4
    cap = int(os.environ.get("TRIS_CPU_ARM_MAX_BATCH", "48"))     # the CPU test-suite caps the sample to keep it short
    for b in (48, 24, 16, 8):
        if b <= cap and (steps + warmup) * b * per_sample * 0.6 <= budget_s:   # larger batches run ~1.5x more efficiently than the probe
            batch = b
            break
    times, loss = [], float("nan")
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss = step(batch)
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    sec = statistics.median(times)
    return {"value": batch / sec, "sec": sec, "batch": batch, "kind": kind, "loss": loss, "steps": steps, "warmup": warmup}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    warnings.simplefilter("ignore")
    threads = os.cpu_count() or 1
    steps, warm = max(1, a.steps), max(0, a.warmup)
    r = cpu_arm(steps, warm, budget_s=150.0, threads=threads)
    what = ("UNMODIFIED reference modules (model_stage1.TRIS + CLIP ViT-B/32 + torch.optim.AdamW), loop body of "
            "train_stage1.py:320-372" if r["kind"] == "reference" else "oracle port of train_stage1.py:320-372")
    sample = (f"each step = batch {r['batch']} of the bs48 workload on the host cores (CPU, fp32, {what}: fwd + 3 losses + "
              f"bwd + AdamW); median of {steps} after {warm} warm-up + 1 probe step")
    line = {"metric": METRIC, "value": r["value"], "unit": "samples/s", "n_gpus": a.gpus, "steps": steps,
            "warmup": warm, "ms_per_step": r["sec"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "per_gpu_batch": 48, "global_batch": 48 * a.gpus, "parallelism": f"dp{a.gpus}",
                       "sample": sample, "sample_batch": r["batch"]},
            "cpu_baseline": {"value": r["value"], "unit": "samples/s", "cores": threads, "kind": r["kind"],
                             "sample": sample + f"; last loss {r['loss']:.4f}"},
            "e2e": {"value": r["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_reference_gpu_arm(a):
    """The UNMODIFIED reference on the same B200 through its own train_one_epoch (PyTorch eager -> cuDNN / cuBLAS):
    the kernel-for-kernel bar of SURVEY 8(d).  --ref-mode fp32 (as shipped: TF32 off for matmul, cudnn TF32 default),
    tf32 (allow_tf32 everywhere) or bf16 (torch.autocast around the loop).  Not the contract's reference arm (that one
    is the CPU path); printed for BASELINE.md section 5."""
    warnings.simplefilter("ignore")
    from baseline import ref_step as RS
    ns, args = RS.load(batch=a.batch)
    torch.backends.cudnn.deterministic = False       # the reference's setup_seed sets it; eager speed is what is measured
    torch.backends.cudnn.benchmark = True
    if a.ref_mode == "tf32":
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.backends.cudnn.allow_tf32 = True
    model, aux = RS.build_models(ns, args, "cuda")
    opt, sched = RS.make_optimizer(model, args, 100000)
    loader = RS.make_train_loader(3, a.batch)
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if a.ref_mode == "bf16" else torch.autocast("cuda", enabled=False)

    def run(n):
        it = 0
        for _ in range(n):
            with ctx:
                it = RS.train_epoch_reference_loop(ns, args, model, aux, loader[it % 3: it % 3 + 1], opt, sched, iteration=it)
    run(max(1, a.warmup))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run(a.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    line = {"metric": METRIC, "value": a.batch / (ms * 1e-3), "unit": "samples/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "impl": "reference-gpu", "dtype": a.ref_mode,
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "per_gpu_batch": a.batch,
                       "how": "unmodified reference train_one_epoch (train_stage1.py:286-411) in PyTorch eager on cuda:0, "
                              "host batches (its own .cuda() copies + per-step synchronize inside the timed region)"}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu_arm(a):
    import torch.distributed as dist
    warnings.simplefilter("ignore")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from tris_b200 import _lib as L
    from tris_b200 import clip_model, gemm
    from tris_b200.model_stage1 import TRIS
    from tris_b200.synthetic import synthetic_batch
    from tris_b200.train_step import Stage1Trainer
    L.require_device()
    torch.manual_seed(1234)
    B = a.batch
    model = TRIS(make_args()).cuda().train()
    # reference init zero-inits bn3.weight (CLIP/clip/model.py:520-523); give the residual branches a small gain so
    # every conv carries signal and gradient during the benchmark (work is shape-, not value-dependent anyway)
    with torch.no_grad():
        for k, p in model.named_parameters():
            if k.endswith("bn3.weight") and "layer" in k:
                p.uniform_(0.1, 0.3)
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20, allow_random_init=True)
    trainer = Stage1Trainer(model, aux, max_iter=100000)
    n_pool = 3
    host = [synthetic_batch(B, 320, 20, 3, seed=1234 + rank * 1000 + i, pin=True) for i in range(n_pool)]
    dev = [tuple(t.cuda() for t in hb) for hb in host]

    # eager steps: warm-up + launch counting
    torch.cuda.synchronize()
    trainer.step(*dev[0])
    torch.cuda.synchronize()
    l0 = L.launch_count
    trainer.step(*dev[1])
    torch.cuda.synchronize()
    launches_per_step = L.launch_count - l0

    use_graph = not a.no_graph
    if use_graph:
        try:
            trainer.capture(*dev[0], warmup=1)
        except Exception as e:  # pragma: no cover
            if rank == 0:
                print(f"[bench] CUDA graph capture failed, running eager: {e!r}", file=sys.stderr)
            use_graph = False
            trainer.graph = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    dev_step = lambda i: trainer.step(*dev[i % n_pool])

    from tris_b200.train_step import HostBatchPrefetcher
    pf = HostBatchPrefetcher()
    pf.submit(host[0])

    def e2e_step(i):
        # public trainer API with pinned HOST batches: every step issues the H2D copy of one full batch (the next one, on
        # the copy stream, while this step computes) and reads the loss back to the host
        out = trainer.step(*pf.take())
        pf.submit(host[(i + 1) % n_pool])
        return out["loss"].item()

    for i in range(a.warmup):
        dev_step(i)
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(dev_step, a.steps)
    clocks = sampler.stop() if sampler else None
    for i in range(2):
        e2e_step(i)
    ms_e2e = timed(e2e_step, a.steps)
    pf.take()

    # ---- roofline of the dominant kernel (the tcgen05 GEMM / implicit conv), measured INSIDE the step: a second copy of
    # the whole step (same streams, same overlap) is captured into a CUDA graph in which every GEMM launch carries a
    # time-stamp slot; the kernel itself leaves min(%globaltimer at CTA start) / max(%globaltimer at CTA end) there.  The
    # per-launch durations therefore come from graph replays of the real step -- no profiler, no Python enqueue gaps
    # (round 1's eager event pairs summed to more than the step itself).  FLOPs are algorithmic: 2*M*N*K*taps times the share
    # of the call that is not zero padding (stem).
    prof = None
    if rank == 0:
        buf = torch.empty((4096, 2), dtype=torch.int64, device="cuda")
        saved = (trainer.graph, trainer.static, trainer._graph_has_optimizer, trainer._graph_signals)
        gemm.timing = [buf, 0, []]
        try:
            trainer.graph = None
            # world 1: one eager pass + the captured pass (both hand out slots in the same order).  Multi-rank: only rank 0 is
            # here, so no warm-up step (it would all-reduce); the captured graph is forward+backward without collectives.
            wu = 1 if world == 1 else 0
            trainer.capture(*dev[0], warmup=wu)
        finally:
            slots, recs = gemm.timing[1], gemm.timing[2]
            gemm.timing = None
        per_pass = slots // (wu + 1)
        recs = recs[wu * per_pass: (wu + 1) * per_pass]
        sums, spans = [], []
        for _ in range(3):
            buf[:, 0] = torch.iinfo(torch.int64).max
            buf[:, 1] = 0
            trainer.graph.replay()                 # the stamped graph (static inputs = batch 0)
            torch.cuda.synchronize()
            sl = buf[wu * per_pass: (wu + 1) * per_pass].cpu()
            sums.append((sl[:, 1] - sl[:, 0]).clamp(min=0).double().sum().item() * 1e-6)
            spans.append((sl[:, 1].max() - sl[:, 0].min()).item() * 1e-6)
        t_ms = statistics.median(sums)
        fl = sum(r[0] for r in recs)
        prof = {"launches": per_pass, "ms": t_ms, "tflops": fl / (t_ms * 1e-3) / 1e12, "gflop": fl / 1e9,
                "how": "device time stamps (%globaltimer: min over CTA starts, max over CTA ends) left by every GEMM launch inside "
                       "CUDA-graph replays of the whole overlapped step; sum over the launches of one step, median of 3 replays",
                "bytes": sum(r[2] for r in recs), "gflop_issued": sum(r[1] for r in recs) / 1e9,
                "span_ms": statistics.median(spans)}
        trainer.graph, trainer.static, trainer._graph_has_optimizer, trainer._graph_signals = saved

    # ---- cross-modal attention group (K6 tail + K7 + K8: north-star "attn HBM GB/s"): forward of the head on resident
    # c4 / hidden, one CUDA-event pair per repetition, L2 flushed (256 MB write) between repetitions
    k7 = None
    if rank == 0:
        eng = model.engine()
        g = torch.Generator(device="cuda").manual_seed(7)
        c4 = torch.randn(B, 10, 10, 2048, device="cuda", generator=g).abs().to(torch.bfloat16)
        hid = (torch.randn(B, 1024, device="cuda", generator=g) * 0.3).to(torch.bfloat16)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        with torch.no_grad():
            for i in range(8):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                eng.head._fwd(c4, hid, (320, 320), True, save=False)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        t_ms = statistics.median(ts[2:])
        gflop = K7_GFLOP_PER_STEP_B48 * B / 48.0
        mbytes = K7_MBYTES_PER_STEP_B48 * B / 48.0
        k7 = {"what": "vis/lan projection + bilateral cross-modal attention + score + response head, forward, eager launches",
              "ms": t_ms, "gflop": gflop, "tflops": gflop / t_ms, "compulsory_mb": mbytes, "hbm_gbs": mbytes / t_ms}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sust, burst, hbm, src = peaks()
    traffic = share_ncu = None
    tpath = os.path.join(ROOT, "profiles", "r2_gemm_traffic.json")
    if os.path.exists(tpath) and B == 48:
        tj = json.load(open(tpath))
        traffic = tj["dram_read_bytes_per_step"] + tj["dram_write_bytes_per_step"]
        share_ncu = tj.get("share_of_step")
    sps = B * world * a.steps / (ms * 1e-3)
    sps_e2e = B * world * a.steps / (ms_e2e * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        threads = os.cpu_count() or 1
        r = cpu_arm(1, 0, budget_s=25.0, threads=threads)
        cpu = {"value": r["value"], "unit": "samples/s", "cores": threads, "kind": r["kind"],
               "sample": f"1 step of batch {r['batch']} of the same workload (fwd + 3 losses + bwd + AdamW, "
                         f"{'unmodified reference modules' if r['kind'] == 'reference' else 'oracle port'}, CPU fp32) in {r['sec']:.1f} s "
                         "on the host cores, after a batch-4 probe step"}
    line = {
        "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": B * world,
                   "parallelism": f"dp{world}", "cuda_graph": use_graph, "l2": "activations per step (>5 GB) exceed the 126 MB L2; "
                   f"{n_pool} rotating input batches", "optimizer": "fused AdamW, 2 lr groups, poly 0.9"},
        "e2e": {"value": sps_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches_per_step * a.steps,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "achieved": prof["tflops"], "peak": sust, "unit": "TFLOP/s", "frac": prof["tflops"] / sust,
                     "traffic": traffic, "traffic_unit": "bytes per step (all GEMM launches, ncu dram__bytes_read+write, profiles/r2_gemm_traffic.json)",
                     "algorithmic_bytes": prof["bytes"],
                     "kernel": "tris_umma_gemm_kernel", "peak_source": f"{src} sustained bf16", "timing": prof["how"],
                     "launches_per_step": prof["launches"], "kernel_ms_per_step": prof["ms"], "kernel_gflop_per_step": prof["gflop"],
                     "kernel_gflop_issued_per_step": prof["gflop_issued"], "first_to_last_gemm_ms": prof["span_ms"],
                     "kernel_share_of_step_ncu": share_ncu,
                     "step_frac_of_peak": (GFLOP_PER_SAMPLE * 1e9 * sps / world) / (sust * 1e12)},
        "cross_modal_attention": k7,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=48, help="per-GPU batch (BASELINE config: 48)")
    ap.add_argument("--impl", default="tris_b200", choices=["tris_b200", "reference", "reference-gpu"])
    ap.add_argument("--ref-mode", default="fp32", choices=["fp32", "tf32", "bf16"], help="--impl reference-gpu precision")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    if a.impl == "reference-gpu":
        return run_reference_gpu_arm(a)
    a.warmup = max(a.warmup, 3)
    run_gpu_arm(a)


if __name__ == "__main__":
    main()
