#!/bin/bash
# A/B of the CG loop mechanics: programmatic dependent launch (SB_NO_PDL) x evict_last vector lines (SB_VEC_KEEP)
set -u
mkdir -p gpurun_out
run() { # label, env...
  label=$1; shift
  for fmt_n in "CRS 128" "SCS 128" "SCS 256"; do
    set -- "$@"
    f=${fmt_n% *}; n=${fmt_n#* }
    out=$(env "$@" timeout 300 python tools/spmv_probe.py --n $n --fmt $f --reps 10 --cg ${SB_CG_ITERS:-200} 2>&1 | grep -E "^cg k|rror" | head -2 | tr '\n' ' ')
    echo "$label $f $n^3: $out"
  done
}
for rep in 1 2; do
  run "pdl+keep " SB_X=1
  run "nopdl+keep" SB_NO_PDL=1
  run "pdl+nokeep" SB_VEC_KEEP=0
  run "nopdl+nokeep" SB_NO_PDL=1 SB_VEC_KEEP=0
done 2>&1 | tee gpurun_out/loop_ab.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
