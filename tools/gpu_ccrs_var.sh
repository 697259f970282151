#!/bin/bash
# needs a sweep build: SB_BUILD_SWEEPS=1 python -m sparsebench_b200.build --force  (the default build carries only the chosen configuration)
for v in 0 4 8 12; do echo "== SB_ROWS_VAR=$v"; SB_ROWS_VAR=$v python tools/spmv_probe.py --n 256 --fmt CCRS --reps 20 --dot --cg 20 2>&1 | grep -E "^spmv|^cg k"; done
