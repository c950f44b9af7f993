#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json > gpurun_out/r2_launch_summary2.txt 2>&1
head -45 gpurun_out/r2_launch_summary2.txt
