#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_stage1_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -v "Missing key\|Unexpected key" | cut -c1-400 > gpurun_out/r2_pytest_stage1.txt; grep -n "^E \|Error\|FAILED\|passed\|failed\|rel cls\|losses\|^48\|^8 \|differ" gpurun_out/r2_pytest_stage1.txt | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
