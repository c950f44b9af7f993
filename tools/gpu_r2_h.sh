#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -q --timeout 300 -x 2>&1 | grep -v "Missing key\|Unexpected key" | cut -c1-300 > gpurun_out/r2_pytest_gpu.txt; grep -n "^E \|Error\|FAILED\|passed\|failed\|worst" gpurun_out/r2_pytest_gpu.txt | head -30
for v in 1 0; do TRIS_PDL=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench.err | cut -c1-190; tail -2 gpurun_out/bench.err; done
