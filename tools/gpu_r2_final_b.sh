#!/bin/bash
# round-2 evidence visit B: ncu launch list of one eager step, ncu --set full of every fusion-head launch and of the top kernels
# (exported to CSV on the box: the .ncu-rep files exceed the 64 MiB return limit), section / layer / timeline tools
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json > gpurun_out/launch_summary.txt 2>&1
head -5 gpurun_out/launch_summary.txt
gzip -9 gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_head python tools/ncu_head.py > gpurun_out/ncu_head.log 2>&1; tail -1 gpurun_out/ncu_head.log
python tools/ncu_table.py /tmp/prof_head.ncu-rep > gpurun_out/ncu_head_table.txt 2>&1; tail -3 gpurun_out/ncu_head_table.txt
timeout 900 ncu --set full --clock-control none --profile-from-start off -o /tmp/prof_targets python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log
python tools/ncu_table.py /tmp/prof_targets.ncu-rep > gpurun_out/ncu_targets_table.txt 2>&1
python tools/ncu_report.py /tmp/prof_targets.ncu-rep > gpurun_out/ncu_targets_report.txt 2>&1
cat gpurun_out/ncu_targets_table.txt
python tools/profile_sections.py 48 > gpurun_out/sections.txt 2>&1
python tools/bench_rn50_layers.py > gpurun_out/rn50_layers.txt 2>&1
python tools/step_timeline.py > gpurun_out/step_timeline.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2_smoke.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r2_smoke.txt
timeout 120 python demo.py --synthetic --synthetic-weights --output gpurun_out/demo_cam.npy 2>&1 | tail -1 | tee gpurun_out/demo_entry.txt
rm -f gpurun_out/demo_cam.npy
du -sh gpurun_out
