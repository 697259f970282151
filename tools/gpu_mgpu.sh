#!/bin/bash
# multi-GPU bench under torchrun: bash tools/gpu_mgpu.sh N [extra bench args]
set -u
N=${1:-2}; shift || true
mkdir -p gpurun_out
nvidia-smi -L | head -8
nvidia-smi topo -m 2>/dev/null | head -12
SB_BENCH_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | cut -c1-3000
