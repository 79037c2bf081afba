#!/bin/bash
# ncu full capture of the verify kernel at full size
mkdir -p gpurun_out
CMD="python bench.py --reads ${READS:-10000000} --steps 1 --warmup 3 --e2e-steps 1 --no-cpu-baseline"
$CMD > gpurun_out/plain_v.json 2> gpurun_out/plain_v.err && ncu --set full --clock-control none --import-source on -k "regex:kj_verify" -s 5 -c 1 -f -o gpurun_out/prof_verify $CMD > gpurun_out/ncu_verify.log 2>&1
echo "rc=$?"
