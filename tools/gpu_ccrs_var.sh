#!/bin/bash
for v in 0 4 8 12; do echo "== SB_ROWS_VAR=$v"; SB_ROWS_VAR=$v python tools/spmv_probe.py --n 256 --fmt CCRS --reps 20 --dot --cg 20 2>&1 | grep -E "^spmv|^cg k"; done
