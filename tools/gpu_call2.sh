#!/bin/bash
# round 2 (1 GPU): smoke, all GPU parity tests, bench with trace, reference arm, ncu launch list + full capture of the
# extraction kernels.  Logs -> gpurun_out/.
mkdir -p gpurun_out
set -o pipefail
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q -rs > gpurun_out/pytest_gpu.log 2>&1; T=$?; echo "pytest rc=$T" | tee -a gpurun_out/summary.txt
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --trace > gpurun_out/bench.json 2> gpurun_out/bench.err; B=$?; echo "bench rc=$B" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "${REF:-1}" == "1" ]; then
  timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt
  cat gpurun_out/bench_ref.json
fi
if [ "$B" == "0" ] && [ "${NCU:-1}" == "1" ]; then
  SHORT="python bench.py --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
  ncu --set full --clock-control none --import-source on -k "regex:kj_warp_filter|kj_resolve" -s 6 -c 4 -f -o gpurun_out/prof_scan $SHORT > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
fi
