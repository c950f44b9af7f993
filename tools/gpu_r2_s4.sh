#!/bin/bash
for W in 2 4; do
TRIS_CAM_WRITERS=$W timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 3000 --prms --save_cam --cam_save_dir /dev/shm/cams$W --name_save_dir /dev/shm/names --lanes 3 2>&1 | tail -1 | sed "s/^/shm writers $W: /"
rm -rf /dev/shm/cams$W
done
df -h /dev/shm | tail -1; free -g | head -2
