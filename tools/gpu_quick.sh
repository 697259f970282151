#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
for w in crs128 sell256; do
SB_CG_TRACE=1 timeout 600 python bench.py --workload $w --no-cpu-baseline > gpurun_out/q_$w.json 2> gpurun_out/q_$w.err; echo "rc=$?"; grep sbSolveCG gpurun_out/q_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/q_$w.json').read().strip().splitlines()[-1])
print('$w value', round(d['value'],1), 'it/s', round(d['cg']['iterations_per_sec'],1), 'e2e', round(d['e2e']['value'],1), 'e2e it/s', round(d['e2e']['iterations_per_sec'],1))
PY
done
