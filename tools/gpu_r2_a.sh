#!/bin/bash
# round-2 visit A: drop-in tests against the shipped reference, reference-on-B200 eager lines, autocast-bias table, bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.txt 2>&1; nproc >> gpurun_out/probe.txt
timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -q -s --timeout 600 2>&1 | tail -40 > gpurun_out/r2_dropin.txt; tail -15 gpurun_out/r2_dropin.txt
for m in fp32 tf32 bf16; do timeout 300 python bench.py --impl reference-gpu --ref-mode $m --steps 5 --warmup 2 2>gpurun_out/refgpu_$m.err | tail -1 >> gpurun_out/r2_reference_gpu.jsonl; done
cat gpurun_out/r2_reference_gpu.jsonl
timeout 600 python tools/ref_autocast_bias.py 48 > gpurun_out/r2_autocast_bias.txt 2>&1; cat gpurun_out/r2_autocast_bias.txt | tail -12
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>gpurun_out/bench_ref.err; cat gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
