#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/micro/mma_rate > gpurun_out/r2_mma_rate.txt 2>&1; cat gpurun_out/r2_mma_rate.txt
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q --timeout 600 -x 2>&1 | grep -v "Missing key\|Unexpected key" | cut -c1-400 | tail -60 > gpurun_out/r2_dropin.txt; tail -60 gpurun_out/r2_dropin.txt
