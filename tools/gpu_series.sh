#!/bin/bash
# full scaling series on one 8-GPU box: weak 256^3/GPU at N=1,2,4,8 (+ NCCL transport at 8), strong 512^3 at 2,4,8, parity at 4
set -u
mkdir -p gpurun_out
bash tools/gpu_scale.sh "1 2 4 8"
SB_BENCH_MODES=nccl bash tools/gpu_scale.sh "8"
SB_MODES=peer bash tools/gpu_mcheck.sh 8
for w in strong512sell strong512crs; do for N in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --workload $w > gpurun_out/${w}_n$N.json 2> gpurun_out/${w}_n$N.err
  echo "$w N=$N rc=$? $(python -c "
import json
d=json.loads(open('gpurun_out/${w}_n$N.json').read().strip().splitlines()[-1]); print(round(d['value'],1), d['unit'], round(d['ms_per_step'],4), 'ms/it', round(d['cg']['iterations_per_sec'],1), 'it/s')" 2>&1 | tail -1)"
done; done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 5 --warmup 3 > gpurun_out/ref_n8.json 2> gpurun_out/ref_n8.err; echo "ref N=8 rc=$?"; cut -c1-250 gpurun_out/ref_n8.json
