#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_stage1_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -5 | cut -c1-300
python tools/debug_tower_attrib.py 8 4321 2>&1 | tail -12 | tee gpurun_out/tower_attrib.txt
python tools/debug_tower_attrib.py 48 4321 2>&1 | tail -12 | tee -a gpurun_out/tower_attrib.txt
for i in 1 2; do python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['value'])"; done
