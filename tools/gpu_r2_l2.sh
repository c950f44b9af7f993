#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py tests/test_stage1_gpu.py -m gpu -q --timeout 300 2>&1 | tail -8 | cut -c1-300
P="python tools/loss_bias_probe.py 4:99 3:1234 48:4321 48:1234 48:7 48:99 48:2024"
( $P; TRIS_RESIDUAL_F32=0 $P ) 2>&1 | grep -v "^$" | tee gpurun_out/loss_probe2.txt
python tools/debug_l4_split.py 48:4321 4:99 2>&1 | tee gpurun_out/l4_split2.txt
for i in 1 2; do python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['value'])"; done
TRIS_RESIDUAL_F32=0 python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench bf16 stream', d['ms_per_step'], d['value'])"
