#!/bin/bash
mkdir -p gpurun_out
b() { python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', round(d['ms_per_step'],3))"; }
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
python bench.py --steps 20 --warmup 5 2>/dev/null | b new_default
TRIS_RESIDUAL_F32=0 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream
TRIS_RESIDUAL_F32=0 TRIS_LIB_PATH=$PWD/tris_b200/lib/libtris_variant.so python bench.py --steps 20 --warmup 5 2>/dev/null | b new_bf16stream_gemm_without_resf32
TRIS_VIT_RESIDUAL_F32=1 python bench.py --steps 20 --warmup 5 2>/dev/null | b new_vit_f32_too
(cd _old && python bench.py --steps 20 --warmup 5 2>/dev/null | b old)
python bench.py --steps 20 --warmup 5 2>/dev/null | b new_default
