#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -3 | cut -c1-300
for L in 1 2 3 4; do
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 3000 --prms --save_cam --cam_save_dir /tmp/cams$L --name_save_dir /tmp/names --lanes $L 2>&1 | tail -1 | sed "s/^/lanes $L: /"
done
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 3000 --prms --lanes 3 2>&1 | tail -1 | sed "s/^/lanes 3 no save: /"
for L in 1 3; do
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 2000 --lanes $L 2>&1 | tail -1 | sed "s/^/plain lanes $L: /"
done
cmp /tmp/cams1/17_17.npy /tmp/cams3/17_17.npy && echo "lane outputs identical"
