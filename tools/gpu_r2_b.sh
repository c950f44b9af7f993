#!/bin/bash
mkdir -p gpurun_out
for m in 0 1 2 4 3 5 6; do TRIS_GEMM_DEBUG=$m timeout 200 python tools/ablate_gemm.py 2>&1 | tail -2; done > gpurun_out/r2_ablate.txt
cat gpurun_out/r2_ablate.txt
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q --timeout 600 2>&1 | tail -5 | cut -c1-300
