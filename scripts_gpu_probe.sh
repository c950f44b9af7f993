#!/bin/bash
# first GPU contact: environment probe + GEMM parity, each test file under its own timeout
mkdir -p gpurun_out
{ nvidia-smi; nproc; free -g | head -2; ls /root/reference 2>&1 | head -3; python -c "import torch;print(torch.__version__, torch.cuda.get_device_name(0))"; } > gpurun_out/probe.txt 2>&1
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 120 2>&1 | tail -40 > gpurun_out/gemm_test.txt
cat gpurun_out/gemm_test.txt
