#!/bin/bash
set -u
run() { echo "== $N $FMT $*"; env "$@" timeout 200 python tools/spmv_probe.py --n $N --fmt $FMT --reps 5 --cg 60 2>&1 | tail -2; }
N=256
FMT=CRS; run SB_DOT_MODE=0; run SB_DOT_MODE=1; run SB_DOT_MODE=2; run SB_DOT_MODE=3; run SB_CG_SPLIT_DOT=1
FMT=SCS; run A=1; run SB_CG_SPLIT_DOT=1
