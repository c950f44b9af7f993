#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for cfg in 1; do TRIS_DP_OVERLAP=$cfg timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_dp_$cfg.err | cut -c1-200; echo "overlap=$cfg rc=$?"; grep -i "falling back\|error\|trap" gpurun_out/bench_dp_$cfg.err | head -3; done | tee gpurun_out/r2_dp_bench_n$N.txt
