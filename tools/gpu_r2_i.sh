#!/bin/bash
mkdir -p gpurun_out
P="python tools/loss_bias_probe.py 4:99 48:4321 48:1234 48:7"
( $P; TRIS_FUSE_IN=0 $P; TRIS_CONV_HALO=0 $P; TRIS_GEMM_NG=1 $P; TRIS_FUSE_IN=0 TRIS_CONV_HALO=0 TRIS_GEMM_NG=1 $P ) 2>&1 | grep -v "^$" | tee gpurun_out/loss_probe.txt
