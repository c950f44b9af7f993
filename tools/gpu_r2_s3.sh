#!/bin/bash
for W in 4 8 12; do
TRIS_CAM_WRITERS=$W timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 3000 --prms --save_cam --cam_save_dir /tmp/cams$W --name_save_dir /tmp/names --lanes 3 2>&1 | tail -1 | sed "s/^/writers $W: /"
done
df -h /tmp | tail -1
