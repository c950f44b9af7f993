#!/bin/bash
mkdir -p gpurun_out
b() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3))"; }
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
python bench.py --steps 20 --warmup 5 2>/dev/null | b new
TRIS_RESIDUAL_F32=0 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream
TRIS_RESIDUAL_F32=0 TRIS_PACK_SIDE=0 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream_packserial
TRIS_RESIDUAL_F32=0 TRIS_ZERO_SIDE=0 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream_zeroserial
TRIS_RESIDUAL_F32=0 TRIS_ZERO_SIDE=0 TRIS_PACK_SIDE=0 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream_bothserial
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
