#!/bin/bash
# round 2, call 1 (2 GPUs): TMA streaming experiment, the new large-template scoring test, the NCCL parity test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 tools/exp/scan_exp.bin > gpurun_out/scan_exp.txt 2>&1; echo "exp rc=$?"
cat gpurun_out/scan_exp.txt
timeout 900 python -m pytest tests/test_dist_nccl_gpu.py tests/test_score_gpu.py -m gpu -x -q -rs > gpurun_out/r02_pytest_nccl_score.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02_pytest_nccl_score.log
