#!/bin/bash
# round-2 ncu evidence (one GPU; every ncu run directly follows a plain run of the same command that exited 0):
#   launch list of the default bench loop, --set full of the SELL / CRS SpMV kernels (plain, fused dot, gated order)
#   and of the two CG vector kernels
set -u
PART=${1:-all}    # launches | crs | sell | all  (a --set full capture is ~4 MB per kernel; gpurun brings back <= 64 MiB per call)
mkdir -p gpurun_out
if [ "$PART" = "launches" ] || [ "$PART" = "all" ]; then
B="python bench.py --steps 5 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
$B > gpurun_out/r2_plain_short.json 2> gpurun_out/r2_plain_short.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_sell256_launches.csv $B > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
B1="python bench.py --workload crs128 --steps 5 --warmup 3 --no-extra --no-e2e --no-cpu-baseline"
$B1 > gpurun_out/r2_plain_crs128.json 2> gpurun_out/r2_plain_crs128.err &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_crs128_launches.csv $B1 > gpurun_out/r2_ncu_launches128.log 2>&1
echo "ncu launches crs128 rc=$?"
fi
if [ "$PART" = "crs" ] || [ "$PART" = "all" ]; then
P="python tools/spmv_probe.py --n 256 --fmt CRS --reps 2 --dot --ordered --cg 2"
timeout 300 $P > gpurun_out/r2_probe_crs.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:spmvRowsPipe -s 8 -c 7 -f -o gpurun_out/r2_prof_crs256 $P > gpurun_out/r2_ncu_crs.log 2>&1
echo "ncu crs rc=$?"
fi
if [ "$PART" = "sell" ] || [ "$PART" = "all" ]; then
P="python tools/spmv_probe.py --n 256 --fmt SCS --reps 2 --dot --ordered --cg 2"
timeout 300 $P > gpurun_out/r2_probe_sell.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"spmvSell32Tma|cgUpdate" -s 8 -c 9 -f -o gpurun_out/r2_prof_sell256 $P > gpurun_out/r2_ncu_sell.log 2>&1
echo "ncu sell rc=$?"
fi
ls -la gpurun_out/*.ncu-rep
