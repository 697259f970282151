#!/bin/bash
# round-end single-GPU evidence: parity tests, smoke, bench lines, ncu launch list + full captures
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_sell256.json 2> gpurun_out/bench_sell256.err; echo "bench sell256 rc=$?"
timeout 900 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref_sell256.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 600 python bench.py --workload crs128 --no-cpu-baseline > gpurun_out/bench_crs128.json 2> gpurun_out/bench_crs128.err; echo "bench crs128 rc=$?"
timeout 600 python bench.py --workload crs256 --no-cpu-baseline > gpurun_out/bench_crs256.json 2> gpurun_out/bench_crs256.err; echo "bench crs256 rc=$?"
timeout 600 python bench.py --workload ccrs256 --no-cpu-baseline > gpurun_out/bench_ccrs256.json 2> gpurun_out/bench_ccrs256.err; echo "bench ccrs256 rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/plain_short.json 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_sell256.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 300 python tools/spmv_probe.py --n 256 --fmt SCS --reps 2 --cg 2 > gpurun_out/plain_probe.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmvSell32Tma -s 4 -c 3 -f -o gpurun_out/prof_sell256_final python tools/spmv_probe.py --n 256 --fmt SCS --reps 2 --cg 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu sell rc=$?"
