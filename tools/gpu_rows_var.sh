#!/bin/bash
# needs a sweep build: SB_BUILD_SWEEPS=1 python -m sparsebench_b200.build --force  (the default build carries only the chosen configuration)
# SB_ROWS_VAR sweep of the CRS pipeline kernel (4 lanes per row): plain vs fused-dot SpMV at 256^3, CG at 128^3
set -u
mkdir -p gpurun_out
for v in ${SB_VARS:-0 1 2 3 4 5}; do
  echo "== SB_ROWS_VAR=$v"
  SB_ROWS_VAR=$v timeout 300 python tools/spmv_probe.py --n 256 --fmt CRS --reps 30 --dot --cg 30 2>&1 | grep -E "^spmv|^cg"
  SB_ROWS_VAR=$v timeout 300 python tools/spmv_probe.py --n 128 --fmt CRS --reps 50 --dot --cg 100 2>&1 | grep -E "^spmv|^cg"
done 2>&1 | tee gpurun_out/rows_var_sweep.log
