#!/usr/bin/env python
"""Stage-1 training throughput benchmark (BASELINE.json metric: samples/sec at 320x320, len 20, bs 48 per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps 2 --warmup 0        # CPU arm (oracle port of the reference step)

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
    return argparse.Namespace(bert_tokenizer="clip", backbone="clip-RN50", max_query_len=20, hidden_dim=1024,
                              attn_multi=0.1, FOCAL_P=3, FOCAL_LAMBDA=0.01)


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_step(batch, threads, steps=1, warmup=0):
    """Times the oracle port of the reference training step (train_stage1.py:320-372) on the host cores."""
    from oracle import tris_oracle as O
    from oracle import weights as W
    torch.set_num_threads(threads)
    sd = W.make_tris_state_dict(0)
    aux = W.make_vitb32_clip_state_dict(7, cos_bias=True)
    img, ids, negs = W.synthetic_batch(batch, 320, 20, 3, 1234)
    keys = O.trainable_keys(sd)
    m = {k: torch.zeros_like(sd[k]) for k in keys}
    v = {k: torch.zeros_like(sd[k]) for k in keys}
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        losses, grads, new_stats, _ = O.train_step(sd, aux, img, ids, negs)
        for k in keys:
            grp = O.param_group_of(k)
            if grp < 0:
                continue
            lr = 5e-5 * (0.1 if grp == 0 else 1.0) * O.poly_lr(it, 1000)
            sd[k], m[k], v[k] = O.adamw_step(sd[k], grads[k], m[k], v[k], it + 1, lr)
        sd.update(new_stats)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    return batch / statistics.mean(times), statistics.mean(times), float(losses["loss"])


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = 4
    sps, sec, loss = cpu_reference_step(batch, threads, steps=max(1, a.steps), warmup=min(a.warmup, 1))
    line = {"metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": a.gpus, "steps": max(1, a.steps),
            "warmup": min(a.warmup, 1), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": WORKLOAD, "per_gpu_batch": 48, "global_batch": 48 * a.gpus, "parallelism": f"dp{a.gpus}",
                       "sample": f"each step = batch {batch} of the bs48 workload on the host cores (CPU, fp32, oracle port of "
                                 "train_stage1.py:320-372: fwd + 3 losses + bwd + AdamW)"},
            "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{max(1, a.steps)} step(s) of batch {batch} (fwd + 3 losses + bwd + AdamW), loss {loss:.4f}"},
            "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
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
    aux, _ = clip_model.load("ViT-B/32", device="cuda", txt_length=20)
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

    # ---- roofline of the dominant kernel (the tcgen05 GEMM / implicit conv): one CUDA-event pair per GEMM launch of one
    # forward+backward; side-stream overlap is off for this pass so no other kernel shares the SMs with a timed GEMM.
    prof = None
    if rank == 0:
        import tris_b200.engine as E
        saved_graph, trainer.graph = trainer.graph, None
        saved_overlap, E.OVERLAP = E.OVERLAP, False
        gemm_calls = []
        orig = L.gemm_raw

        def timed_gemm(desc):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            orig(desc)
            e.record()
            taps = desc.taps if (desc.wgrad and desc.taps > 1) else 1
            nb = max(1, desc.batch)
            esz = 4 if desc.out_dtype == L.DT_F32 else 2
            if desc.a_mode == L.OP_CONV and not desc.wgrad:       # activation read once (not once per tap) + weights + output
                by = 2.0 * desc.M * (desc.K // max(1, desc.taps)) + 2.0 * desc.N * desc.K + esz * desc.M * desc.N
            else:
                by = 2.0 * desc.M * desc.K * (nb if desc.a_batch_stride else 1) + 2.0 * desc.N * desc.K * (nb if desc.b_batch_stride else 1) \
                    + esz * desc.M * desc.N * taps * nb
            gemm_calls.append((s, e, 2.0 * desc.M * desc.N * desc.K * taps * nb, by))

        L.gemm_raw = timed_gemm
        gemm.L.gemm_raw = timed_gemm
        # queue the whole eager pass behind a ~100 ms spin kernel so that the launches are already enqueued when the GPU
        # reaches them: the per-launch event pairs then measure kernel durations rather than Python launch gaps
        how = "CUDA-event pair per launch, eager fwd+bwd queued behind a 100 ms spin kernel, side-stream overlap off"
        torch.cuda.synchronize()
        try:
            torch.cuda._sleep(int(2.0e8))
        except Exception:  # pragma: no cover  (private helper; without it the figure only gets more conservative)
            how = "CUDA-event pair per launch, eager fwd+bwd, side-stream overlap off"
        trainer._fwd_bwd(*dev[0])
        torch.cuda.synchronize()
        t_ms = sum(c[0].elapsed_time(c[1]) for c in gemm_calls)
        L.gemm_raw = orig
        gemm.L.gemm_raw = orig
        trainer.graph = saved_graph
        E.OVERLAP = saved_overlap
        fl = sum(c[2] for c in gemm_calls)
        prof = {"launches": len(gemm_calls), "ms": t_ms, "tflops": fl / (t_ms * 1e-3) / 1e12, "gflop": fl / 1e9, "how": how,
                "bytes": sum(c[3] for c in gemm_calls)}

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
    tpath = os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")
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
        v, sec, loss = cpu_reference_step(4, threads, steps=1, warmup=0)
        cpu = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
               "sample": f"1 step of batch 4 of the same workload (fwd + 3 losses + bwd + AdamW) in {sec:.1f} s on the host cores"}
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
                     "traffic": traffic, "traffic_unit": "bytes per step (all GEMM launches, ncu dram__bytes_read+write, profiles/r1_gemm_traffic.json)",
                     "algorithmic_bytes": prof["bytes"],
                     "kernel": "tris_umma_gemm_kernel", "peak_source": f"{src} sustained bf16", "timing": prof["how"],
                     "launches_per_step": prof["launches"], "kernel_ms_per_step": prof["ms"], "kernel_gflop_per_step": prof["gflop"],
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
    ap.add_argument("--impl", default="tris_b200", choices=["tris_b200", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    a.warmup = max(a.warmup, 3)
    run_gpu_arm(a)


if __name__ == "__main__":
    main()
