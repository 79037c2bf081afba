#!/bin/bash
bash tools/gpu_final.sh
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --e2e-steps 1 --no-file-leg --trace > gpurun_out/bench_c5_n1.json 2> gpurun_out/bench_c5_n1.err; echo "c5 rc=$?"
grep trace gpurun_out/bench_c5_n1.err | tail -2; head -c 400 gpurun_out/bench_c5_n1.json; echo
