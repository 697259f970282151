#!/bin/bash
set -u
N=${1:-2}
for mode in peer nccl; do
SB_COMM=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/comm_probe.py 2>&1 | grep -E "rank|rror"
done
