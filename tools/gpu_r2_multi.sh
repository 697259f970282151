#!/bin/bash
# round-2 multi-GPU check on N ranks: full parity script (both transports), the GPU test-suite's multi-GPU tests,
# the driver-style bench line (20 steps) and a 100-step line:   bash tools/gpu_r2_multi.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in ${SB_MODES:-peer nccl}; do
  SB_COMM=$mode timeout 900 $LAUNCH --master-port 29541 tests/mgpu_check.py > gpurun_out/mcheck_${mode}_n$N.log 2>&1
  echo "mcheck mode=$mode N=$N rc=$?"; grep -E "FAIL|PASS|rror" gpurun_out/mcheck_${mode}_n$N.log | head -20
done
if [ "${SB_PYTEST:-1}" = "1" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q --timeout 800 > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_n$N.log
fi
for steps in ${SB_STEPS:-20 100}; do
  SB_BENCH_VERBOSE=1 SB_CG_TRACE=1 timeout 900 $LAUNCH --master-port 29533 bench.py --gpus $N --steps $steps --warmup 5 > gpurun_out/r2_scale_n${N}_s$steps.json 2> gpurun_out/r2_scale_n${N}_s$steps.err
  echo "bench N=$N steps=$steps rc=$?"
  grep "sbSolveCG\] create" gpurun_out/r2_scale_n${N}_s$steps.err | tail -4
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r2_scale_n${N}_s$steps.json').read().strip().splitlines()[-1])
    print('  value', round(d['value'],1), d['unit'], ' ms/it', round(d['ms_per_step'],4), ' it/s', round(d['cg']['iterations_per_sec'],1), ' e2e', round(d['e2e']['value'],1), 'e2e ms/step', round(d['e2e']['ms_per_step'],3))
    print('  e2e breakdown', d['e2e']['breakdown_ms'])
    print('  regions', {k: round(v,4) for k,v in d['cg']['kernel_ms_per_iteration'].items()})
    print('  parity', d.get('parity'))
    if d.get('configs4'):
        for k,v in d['configs4']['formats'].items(): print('  configs4', k, {kk: v.get(kk) for kk in ('iterations_per_sec','frac_of_peak','spmv_frac_of_peak','failed')})
except Exception as e:
    print('  no result', e); print(open('gpurun_out/r2_scale_n${N}_s$steps.err').read()[-2500:])
PY
done
