#!/bin/bash
# multi-GPU parity check: bash tools/gpu_mcheck.sh N
set -u
N=${1:-2}
mkdir -p gpurun_out
for mode in ${SB_MODES:-default}; do
  SB_COMM=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py > gpurun_out/mcheck_${mode}_n$N.log 2>&1
  echo "mcheck mode=$mode N=$N rc=$?"; grep -E "FAIL|PASS|rror" gpurun_out/mcheck_${mode}_n$N.log | head -40
done
