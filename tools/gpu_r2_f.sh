#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_head_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5 | cut -c1-300
for m in 0 5; do TRIS_GEMM_DEBUG=$m timeout 200 python tools/ablate_gemm.py 2>&1 | tail -2 | head -1; done > gpurun_out/r2_ablate_elect.txt
cat gpurun_out/r2_ablate_elect.txt
python tools/bench_rn50_layers.py > gpurun_out/r2_rn50_layers_elect.txt 2>&1; tail -3 gpurun_out/r2_rn50_layers_elect.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-1200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
