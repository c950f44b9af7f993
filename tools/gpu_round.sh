#!/bin/bash
# One GPU visit: parity tests, bench line, entry points, ncu launch list of one eager step, ncu --set full of the top kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.txt 2>&1; nproc >> gpurun_out/probe.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -30 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" != "quick" ]; then
timeout 300 python train_stage1.py --synthetic --synthetic-weights --batch_size 48 --size 320 --max_query_len 20 --negative_samples 3 --epoch 1 --steps_per_epoch 20 --print-freq 10 --val_refs 8 > gpurun_out/train_entry.txt 2>&1; tail -4 gpurun_out/train_entry.txt
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 200 > gpurun_out/validate_entry.txt 2>&1; tail -2 gpurun_out/validate_entry.txt
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 200 --prms >> gpurun_out/validate_entry.txt 2>&1; tail -1 gpurun_out/validate_entry.txt
timeout 120 python demo.py --synthetic --synthetic-weights --output gpurun_out/demo_cam.npy > gpurun_out/demo_entry.txt 2>&1; tail -1 gpurun_out/demo_entry.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off -c 3000 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu.log 2>&1
python tools/summarize_launches.py gpurun_out/launches.csv gpurun_out/gemm_traffic.json > gpurun_out/launch_summary.txt 2>&1
head -30 gpurun_out/launch_summary.txt
python tools/profile_sections.py 48 > gpurun_out/sections.txt 2>&1; cat gpurun_out/sections.txt
rm -f gpurun_out/prof_targets.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_targets python tools/ncu_targets.py > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
REPS=3 timeout 300 python tools/bench_gemm_shapes.py > gpurun_out/gemm_shapes.txt 2>&1; cat gpurun_out/gemm_shapes.txt
fi
