#!/bin/bash
# A/B of kernel variants built side by side (kmerjs_b200/variants/*.so): short bench per variant
mkdir -p gpurun_out
for lib in kmerjs_b200/variants/*.so; do
  KMERJS_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-file-leg > gpurun_out/v.json 2> gpurun_out/v.err || { echo "$lib FAILED"; tail -3 gpurun_out/v.err; continue; }
  python -c "
import json,sys; d=json.load(open('gpurun_out/v.json')); r=d['roofline']
print('$lib', 'ms/step', round(d['ms_per_step'],3), 'scan', round(r['scan_kernel_ms'],3), 'resolve', round(r['resolve_kernel_ms'],3), 'frac', round(r['frac'],3), 'rows', d['result']['rows'], 'uniq', d['result']['unique_kmers'])" | tee -a gpurun_out/variants.txt
done
