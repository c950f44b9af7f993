#!/bin/bash
# round-2 evidence visit A: full GPU test-suite, smoke (twice: bit-identical), bench (both arms), entry points
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe.txt 2>&1; nproc >> gpurun_out/probe.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "Missing key\|Unexpected key" | tail -30 | cut -c1-300 > gpurun_out/r2_pytest_gpu.txt; tail -6 gpurun_out/r2_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/r2_smoke.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee -a gpurun_out/r2_smoke.txt
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/r2_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/bench.err; cat gpurun_out/r2_bench.json; tail -3 gpurun_out/bench.err
timeout 300 python train_stage1.py --synthetic --synthetic-weights --batch_size 48 --size 320 --max_query_len 20 --negative_samples 3 --epoch 1 --steps_per_epoch 20 --print-freq 10 --val_refs 8 > gpurun_out/train_entry.txt 2>&1; tail -4 gpurun_out/train_entry.txt
timeout 300 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 1000 > gpurun_out/validate_entry.txt 2>&1; tail -1 gpurun_out/validate_entry.txt
timeout 600 python validate.py --synthetic --synthetic-weights --size 320 --max_query_len 20 --val_refs 5000 --prms --save_cam --cam_save_dir /tmp/cams --name_save_dir /tmp/names >> gpurun_out/validate_entry.txt 2>&1; tail -1 gpurun_out/validate_entry.txt; ls /tmp/cams | wc -l >> gpurun_out/validate_entry.txt
timeout 120 python demo.py --synthetic --synthetic-weights --output gpurun_out/demo_cam.npy > gpurun_out/demo_entry.txt 2>&1; tail -1 gpurun_out/demo_entry.txt
