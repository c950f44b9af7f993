#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "Missing key\|Unexpected key" | tail -30 | cut -c1-400 > gpurun_out/r2_pytest_gpu.txt; tail -12 gpurun_out/r2_pytest_gpu.txt
python tools/loss_bias_probe.py 4:99 3:1234 8:4321 48:4321 48:1234 48:7 48:99 48:2024 2>&1 | grep -v "^$" | tee gpurun_out/loss_probe3.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
b() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3))"; }
python bench.py --steps 20 --warmup 5 2>/dev/null | b new_default
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
python bench.py --steps 20 --warmup 5 2>/dev/null | b new_default
