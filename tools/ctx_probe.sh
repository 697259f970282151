#!/bin/bash
# does a second CUDA context on the GPU slow the linked reference driver down? (managed x/y in -t spmv vs CG)
E=./integration/_build/sparseBench-SCS-B200
echo "== alone"; $E -t spmv -x 256 -y 256 -z 256 -i 131 | grep -E "spMVM"
python - <<'PY' &
import time, ctypes as C, sys
import numpy as np
sys.path.insert(0, '.')
from sparsebench_b200 import api
g = api.matrixGenerate(128, 128, 128, device=True)
A = api.convertMatrix(api.FMT_CRS, g)
k, hist, _, _ = api.solveCG(A, 30, 0.0)
bufs = [api.DeviceBuffer(1 << 30) for _ in range(8)]
u = api.allocate(64, 1 << 26); u.host()[:] = 1.0; api.lib().sbPrefetchManaged(u.ptr)
api.lib().sbDeviceSynchronize()
print("parent context up (matrix, 8 GiB, a managed block)", flush=True)
time.sleep(60)
PY
sleep 15
for mode in 0 1 2 3; do
  echo "== with another process's context, SB_MANAGED_MODE=$mode"; SB_MANAGED_MODE=$mode $E -t spmv -x 256 -y 256 -z 256 -i 131 | grep -E "spMVM"
done
wait
