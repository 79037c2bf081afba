#!/bin/bash
# ncu full capture of the scan kernel on a short bench run (+ host phase trace).  gpurun_out/.
mkdir -p gpurun_out
SHORT="python bench.py --reads ${READS:-2000000} --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$SHORT --trace > gpurun_out/plain_short.json 2> gpurun_out/plain_short.err; R=$?
cat gpurun_out/plain_short.json; tail -3 gpurun_out/plain_short.err
if [ "$R" == "0" ]; then
  ncu --set full --clock-control none --import-source on -k "regex:kj_scan|kj_verify" -s ${SKIP:-6} -c ${COUNT:-4} -f -o gpurun_out/prof_scan $SHORT > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
