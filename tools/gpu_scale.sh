#!/bin/bash
# weak-scaling bench at the given rank counts on one box: bash tools/gpu_scale.sh "1 2 4 8"
set -u
mkdir -p gpurun_out
for N in $1; do
  for mode in ${SB_BENCH_MODES:-peer}; do
    if [ "$N" = "1" ]; then LAUNCH="python"; else LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
    SB_COMM=$mode timeout 600 $LAUNCH bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n${N}_$mode.json 2> gpurun_out/scale_n${N}_$mode.err
    echo "bench N=$N mode=$mode rc=$?"
    python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/scale_n${N}_$mode.json').read().strip().splitlines()[-1])
    print('  value', round(d['value'],1), d['unit'], ' ms/it', round(d['ms_per_step'],4), ' it/s', round(d['cg']['iterations_per_sec'],1), ' e2e', round(d['e2e']['value'],1))
    print('  regions', {k: round(v,4) for k,v in d['cg']['kernel_ms_per_iteration'].items()})
except Exception as e:
    print('  no result', e); print(open('gpurun_out/scale_n${N}_$mode.err').read()[-1500:])
PY
  done
done
