#!/bin/bash
for m in 0 32 64 96 4; do TRIS_GEMM_DEBUG=$m timeout 200 python tools/ablate_gemm.py 2>&1 | tail -2 | head -1; done | tee gpurun_out/r2_ablate_epilogue.txt
