#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_head_gpu.py tests/test_gemm_gpu.py tests/test_stage1_gpu.py -m gpu -q --timeout 300 2>&1 | tail -4 | cut -c1-300
b() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3))"; }
python bench.py --steps 20 --warmup 5 2>/dev/null | b new
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
python bench.py --steps 20 --warmup 5 2>/dev/null | b new
python tools/profile_sections.py 48 2>&1 | tail -14
