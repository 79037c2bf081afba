#!/bin/bash
# 2 GPUs: all parity tests (incl. NCCL lean path), N=1 bench with trace + ncu, N=2 bench
mkdir -p gpurun_out
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-file-leg --trace > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); r=d['roofline']
print('N=1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scan', round(r['scan_kernel_ms'],3), 'resolve', round(r['resolve_kernel_ms'],3), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3), 'e2e', round(d['e2e']['value'],2))"
tail -3 gpurun_out/bench_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --trace > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); r=d['roofline']
print('N=2 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
grep trace gpurun_out/bench_n2.err | tail -3
ncu --set full --clock-control none --import-source on -k "regex:kj_warp_filter|kj_resolve" -s 9 -c 3 -f -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
