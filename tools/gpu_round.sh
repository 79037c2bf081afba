#!/bin/bash
# Runs on the B200 box under gpurun: smoke, GPU parity tests, bench (+ reference arm), then the ncu
# launch list of a short bench run (only if everything before it exited 0).  Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
node --version > gpurun_out/node_probe.txt 2>&1 || echo "node: not found" >> gpurun_out/node_probe.txt
set -o pipefail
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/summary.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; T=$?; echo "pytest rc=$T" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; B=$?; echo "bench rc=$B" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_ref.json
if [ "$T" == "0" ] && [ "$B" == "0" ] && [ "${NCU:-1}" == "1" ]; then
  SHORT="python bench.py --reads ${READS:-10000000} --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
  $SHORT > gpurun_out/plain_short.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $SHORT > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?" | tee -a gpurun_out/summary.txt
  if [ "${NCU_FULL:-0}" == "1" ]; then
    ncu --set full --clock-control none --import-source on -k regex:kj_scan -s 3 -c 2 -o gpurun_out/prof_scan $SHORT > gpurun_out/ncu_full.log 2>&1
    echo "ncu full rc=$?" | tee -a gpurun_out/summary.txt
  fi
fi
