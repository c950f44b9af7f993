#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dropin_gpu.py tests/test_head_gpu.py -m gpu -q -s --timeout 600 2>&1 | grep -v "Missing key\|Unexpected key" | cut -c1-300 > gpurun_out/r2_pytest_gpu.txt; grep -n "^E \|Error\|FAILED\|passed\|failed\|agreeing\|get_scores" gpurun_out/r2_pytest_gpu.txt | head -20
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 1000 2>&1 | tail -1
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 1000 --prms 2>&1 | tail -1
timeout 600 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 5000 --prms --save_cam --cam_save_dir /tmp/cams --name_save_dir /tmp/names 2>&1 | tail -1; ls /tmp/cams | wc -l
