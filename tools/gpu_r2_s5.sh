#!/bin/bash
for W in 2 4; do
TRIS_CAM_WRITERS=$W timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 3000 --prms --save_cam --cam_save_dir /tmp/cams$W --name_save_dir /tmp/names --lanes 3 2>&1 | tail -1 | sed "s/^/raw-write writers $W: /"
done
ls /tmp/cams2 | wc -l
python -c "
import numpy as np; a=np.load('/tmp/cams2/17_17.npy'); b=np.load('/tmp/cams4/17_17.npy'); print(a.shape,a.dtype,float(a.max()),np.array_equal(a,b))"
timeout 600 python -m pytest tests/test_dropin_gpu.py -m gpu -q -x --timeout 300 2>&1 | tail -2 | cut -c1-300
