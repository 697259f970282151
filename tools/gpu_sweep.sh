#!/bin/bash
# sweep SELL TMA kernel configurations (SB_SELL_CFG), then profile the fastest with ncu
set -u
mkdir -p gpurun_out
best=0; bestms=999
for cfg in 0 1 2 3 4 5; do
  SB_SELL_CFG=$cfg timeout 120 python tools/spmv_probe.py --n 256 --fmt SCS --reps 4 > gpurun_out/sweep_$cfg.log 2>&1
  ms=$(tail -1 gpurun_out/sweep_$cfg.log | sed 's/.*: \([0-9.]*\) ms.*/\1/')
  echo "cfg $cfg: $(tail -1 gpurun_out/sweep_$cfg.log)"
  if python -c "import sys; sys.exit(0 if float('$ms') < float('$bestms') else 1)" 2>/dev/null; then best=$cfg; bestms=$ms; fi
done
echo "best cfg $best ($bestms ms)"
SB_SELL_CFG=$best timeout 120 python tools/spmv_probe.py --n 128 --fmt SCS --reps 3 > gpurun_out/plain_ncu.log 2>&1 &&
SB_SELL_CFG=$best timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmvSell32 -s 1 -c 1 -f -o gpurun_out/prof_sell128_tma python tools/spmv_probe.py --n 128 --fmt SCS --reps 3 > gpurun_out/ncu_full.log 2>&1
echo "ncu rc=$?"
