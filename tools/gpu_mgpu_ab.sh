#!/bin/bash
# where does the multi-GPU iteration lose time? A/B of the transport pieces at N ranks (kernel regions from the profiled pass)
set -u
N=${1:-2}
mkdir -p gpurun_out
LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() {
  label=$1; shift
  env "$@" timeout 600 $LAUNCH bench.py --gpus $N --steps 100 --warmup 5 --no-extra --no-e2e > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/ab_$label.json').read().strip().splitlines()[-1])
    print('%-22s ms/it %.4f  regions %s' % ('$label', d['ms_per_step'], {k: round(v,4) for k,v in d['cg']['kernel_ms_per_iteration'].items() if v}))
except Exception as e:
    print('$label failed', e)
PY
  grep "sync trace r0" gpurun_out/ab_$label.err | tail -1 | cut -c1-330
}
timeout 600 python bench.py --steps 100 --warmup 5 --no-extra --no-e2e --no-cpu-baseline > gpurun_out/ab_single.json 2> gpurun_out/ab_single.err
python -c "
import json; d=json.loads(open('gpurun_out/ab_single.json').read().strip().splitlines()[-1]); print('%-22s ms/it %.4f  regions %s' % ('single GPU', d['ms_per_step'], {k: round(v,4) for k,v in d['cg']['kernel_ms_per_iteration'].items() if v}))"
run default SB_SYNC_TRACE=1
for v in ${SB_AB_VARIANTS:-no_fused_put no_fused_reduce no_overlap nccl no_pdl}; do
  case $v in
    no_fused_put) run $v SB_NO_FUSED_PUT=1;;
    no_fused_reduce) run $v SB_NO_FUSED_REDUCE=1;;
    no_overlap) run $v SB_CG_NO_OVERLAP=1;;
    nccl) run $v SB_COMM=nccl;;
    no_pdl) run $v SB_NO_PDL=1;;
    expnc) run $v SB_LIB=$PWD/sparsebench_b200/libsparsebench_b200_expnc.so;;   # timing experiment: read-only gathers behind the gate
  esac
done
