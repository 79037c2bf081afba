#!/bin/bash
# quick iteration on the extraction kernels (1 GPU): count parity tests, a short bench with trace, one ncu capture
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_count_gpu.py -m gpu -x -q > gpurun_out/pytest_count.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_count.log
timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg --trace > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); r=d['roofline']
print('value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scan', round(r['scan_kernel_ms'],3), 'resolve', round(r['resolve_kernel_ms'],3), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3))"
tail -3 gpurun_out/bench_quick.err
if [ "${NCU:-1}" == "1" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:kj_warp_filter|kj_resolve" -s 6 -c 2 -f -o gpurun_out/prof_scan python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
