#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py tests/test_stage1_gpu.py -m gpu -q --timeout 300 2>&1 | tail -8 | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | cut -c1-200; tail -2 gpurun_out/bench.err
