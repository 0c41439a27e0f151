#!/bin/bash
# development loop on the GPU box: microbenchmarks, smoke, the GPU test-suite; logs go to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python - > gpurun_out/microbench.txt 2>&1 <<'PY'
import ctypes as C, sys
sys.path.insert(0, ".")
from gbrl_b200 import _capi
L = _capi.lib()
names = {0: "ATOMS conflict-free (G lane-atomics/s)", 1: "ATOMS random addr", 2: "ATOMS 3 planes conflict-free", 3: "REDG.64 spread", 4: "stream read GB/s"}
for w, it in ((0, 20000), (1, 20000), (2, 8000), (3, 4000), (4, 5)):
    r = C.c_double()
    _capi.check(L.gbrl_b200_microbench(w, it, C.byref(r)))
    print(w, names[w], "%.1f" % r.value, flush=True)
PY
cat gpurun_out/microbench.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.txt
timeout ${PYTEST_TIMEOUT:-1200} python -m pytest tests -m gpu -q --maxfail=${MAXFAIL:-40} -p no:cacheprovider ${PYTEST_ARGS} > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -60 gpurun_out/pytest_gpu.txt
