#!/bin/bash
set -u
run() { echo "== $N $FMT $*: $(env "$@" timeout 200 python tools/spmv_probe.py --n $N --fmt $FMT --reps 10 --cg 100 2>&1 | tail -2 | tr '\n' ' ' | cut -c1-330)"; }
for N in 128 256; do for FMT in CRS SCS; do
run A=1
done; done
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -3
