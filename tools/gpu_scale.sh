#!/bin/bash
# tools/gpu_scale.sh N [configs...]  -- N GPUs of one box: NCCL parity tests (N = 2 only), then bench lines per config
N=${1:-2}; shift
CONFIGS=${@:-c3}
mkdir -p gpurun_out
if [ "$N" == "2" ] || [ "$N" == "4" ]; then
  timeout 900 python -m pytest tests/test_dist_nccl_gpu.py -m gpu -x -q -rs > gpurun_out/pytest_nccl.log 2>&1; echo "pytest nccl rc=$?"; tail -3 gpurun_out/pytest_nccl.log
fi
for c in $CONFIGS; do
  extra=""
  if [ "$c" != "c3" ]; then extra="--no-file-leg --e2e-steps 1"; fi
  steps=10; if [ "$c" == "c5" ]; then steps=3; fi
  timeout ${BENCH_TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config $c --steps $steps --warmup 3 --trace $extra > gpurun_out/bench_${c}_n$N.json 2> gpurun_out/bench_${c}_n$N.err; echo "$c N=$N rc=$?"
  python -c "
import json
for l in open('gpurun_out/bench_${c}_n$N.json'):
    if l.startswith('{'):
        d=json.loads(l); r=d.get('roofline',{})
        print('$c N=$N value', round(d['value'],1), d['unit'], 'ms/step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'frac', round(r.get('frac',0),3), 'clocks', d.get('clocks',{}).get('sm_mhz'))"
  grep trace gpurun_out/bench_${c}_n$N.err | tail -3 | cut -c1-700
  tail -2 gpurun_out/bench_${c}_n$N.err | cut -c1-300
done
