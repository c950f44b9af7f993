#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_ops_gpu.py tests/test_head_gpu.py -m gpu -q --timeout 300 -x 2>&1 | tail -5 | cut -c1-300
for v in 2 1; do TRIS_GEMM_NG=$v timeout 200 python tools/ablate_gemm.py 2>&1 | tail -2 | head -1; done
for v in 2 1 2; do TRIS_GEMM_NG=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ng', $v, d['ms_per_step'])"; done
