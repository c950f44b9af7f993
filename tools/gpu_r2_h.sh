#!/bin/bash
mkdir -p gpurun_out
for v in 1 0; do TRIS_BN_SHORTK=$v timeout 200 python tools/ablate_gemm.py 2>&1 | tail -2 | head -1 | cut -d' ' -f20-38; done
for v in 1 0 1; do TRIS_BN_SHORTK=$v timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('shortk', $v, d['ms_per_step'])"; done
