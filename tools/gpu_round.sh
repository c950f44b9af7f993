#!/bin/bash
# One GPU visit: parity tests, bench line, ncu launch list of one eager step.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.txt 2>&1; nproc >> gpurun_out/probe.txt
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
head -45 gpurun_out/launch_summary.txt
