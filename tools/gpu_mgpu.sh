#!/bin/bash
# multi-GPU parity + bench under torchrun: bash tools/gpu_mgpu.sh N [extra bench args]
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
true
for mode in ${SB_BENCH_MODES:-peer}; do
SB_COMM=$mode SB_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
echo "bench N=$N mode=$mode rc=$?"; grep -v "bench r" gpurun_out/bench_n${N}_$mode.err | tail -5
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n${N}_$mode.json').read().strip().splitlines()[-1])
print('  value', d['value'], d['unit'], ' ms/it', d['ms_per_step'], ' it/s', d['cg']['iterations_per_sec'])
print('  regions', d['cg']['kernel_ms_per_iteration'])
print('  e2e', d['e2e']['value'], ' launches', d['gpu_launches'])
PY
done
