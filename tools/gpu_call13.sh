#!/bin/bash
bash tools/gpu_quick2.sh
for c in c4 c5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --e2e-steps 1 --no-file-leg --trace > gpurun_out/bench_${c}_n1.json 2> gpurun_out/bench_${c}_n1.err; echo "$c rc=$?"
  tail -2 gpurun_out/bench_${c}_n1.err; head -c 1500 gpurun_out/bench_${c}_n1.json; echo
done
