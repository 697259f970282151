#!/bin/bash
# round-2 single-GPU check: parity tests, smoke, default bench line (with configs1 / configs4 / cpu_baseline), reference arm
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SB_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_sell256.json 2> gpurun_out/bench_sell256.err; echo "bench sell256 rc=$?"
tail -c 600 gpurun_out/bench_sell256.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_sell256.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_sell256.json','gpurun_out/bench_ref_sell256.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],4), 'e2e', d['e2e'] and round(d['e2e']['value'],1))
        if 'roofline' in d: print('  roofline', round(d['roofline']['frac'],3), 'clocks', d['clocks'])
        if d.get('e2e') and 'breakdown_ms' in d['e2e']: print('  e2e breakdown', d['e2e']['breakdown_ms'], d['e2e']['answer_check'])
        if d.get('configs1'): print('  configs1', {k: d['configs1'].get(k) for k in ('ms_per_step','iterations_per_sec','frac_of_peak','failed')})
        if d.get('configs4'):
            for k,v in d['configs4']['formats'].items(): print('  configs4', k, {kk: v.get(kk) for kk in ('iterations_per_sec','frac_of_peak','spmv_frac_of_peak','failed','setup_s')})
        if d.get('cpu_baseline'): print('  cpu', d['cpu_baseline'])
    except Exception as e:
        print(f, 'no result', e)
PY
