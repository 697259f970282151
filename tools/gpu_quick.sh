#!/bin/bash
set -u
mkdir -p gpurun_out
SB_CG_TRACE=1 timeout 900 python bench.py --no-cpu-baseline --steps 100 > gpurun_out/bench_sell256.json 2> gpurun_out/bench_sell256.err; echo "rc=$?"
grep sbSolveCG gpurun_out/bench_sell256.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_sell256.json').read().strip().splitlines()[-1])
for k in ('metric','value','unit','ms_per_step','e2e','roofline','cg'): print(k, d[k])
"
