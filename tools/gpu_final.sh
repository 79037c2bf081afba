#!/bin/bash
# 1 GPU: all parity tests, smoke, bench with trace, the reference arm, launch list, full ncu of the extraction kernels
mkdir -p gpurun_out
set -o pipefail
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 --trace > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); r=d['roofline']
print('N=1 value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'scan', round(r['scan_kernel_ms'],3), 'resolve', round(r['resolve_kernel_ms'],3), 'frac', round(r['frac'],3), 'share', round(r['kernel_share_of_step'],3), 'e2e', round(d['e2e']['value'],2), 'file', d['e2e'].get('file',{}).get('value'), 'parity', d.get('parity_checked'))"
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_ref.json

SHORT="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg"
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:kj_warp_filter|kj_resolve" -s 6 -c 2 -f -o gpurun_out/prof_scan $SHORT > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
