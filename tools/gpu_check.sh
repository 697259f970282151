#!/bin/bash
# One GPU-box session: parity tests, bench, launch list, one full ncu capture. Outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
SB_BENCH_VERBOSE=1 timeout 600 python bench.py --workload sell256 --steps 50 --no-cpu-baseline > gpurun_out/b_sell256.log 2> gpurun_out/b_sell256.err; echo "sell256 rc=$?"
timeout 600 python bench.py --workload crs128 --steps 100 --no-cpu-baseline > gpurun_out/b_crs128.log 2> gpurun_out/b_crs128.err; echo "crs128 rc=$?"
for f in SCS CRS CCRS; do timeout 300 python tools/spmv_probe.py --n 256 --fmt $f --reps 4 > gpurun_out/probe_$f.log 2>&1; done
cat gpurun_out/probe_*.log
timeout 300 python tools/spmv_probe.py --n 128 --fmt SCS --reps 3 --cg 4 > gpurun_out/plain_ncu.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmvSell32 -s 1 -c 2 -f -o gpurun_out/prof_sell128 python tools/spmv_probe.py --n 128 --fmt SCS --reps 3 --cg 4 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
