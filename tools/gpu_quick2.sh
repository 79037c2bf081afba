#!/bin/bash
# quick iteration on the whole step (1 GPU): count + score parity tests, a short bench with trace, launch list of the first steps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_count_gpu.py tests/test_score_gpu.py -m gpu -x -q > gpurun_out/pytest_count.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_count.log
timeout 600 python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg --trace > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scan', round(r['scan_kernel_ms'],3), 'resolve', round(r['resolve_kernel_ms'],3), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3))"
tail -3 gpurun_out/bench_quick.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
