#!/bin/bash
# bench both arms + ncu launch list + one full capture of the dominant kernel
set -u
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | cut -c1-400
timeout 900 python bench.py > gpurun_out/bench_sell256.json 2> gpurun_out/bench_sell256.err; echo "sell256 rc=$?"
timeout 900 python bench.py --workload crs128 --no-cpu-baseline > gpurun_out/bench_crs128.json 2> gpurun_out/bench_crs128.err; echo "crs128 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sell256.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/spmv_probe.py --n 256 --fmt SCS --reps 3 --cg 2 > gpurun_out/plain_probe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmvSell32Tma -s 2 -c 2 -f -o gpurun_out/prof_sell256_r1 python tools/spmv_probe.py --n 256 --fmt SCS --reps 3 --cg 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
