#!/bin/bash
# round-2 validation on the GPU box: smoke, the GPU test-suite, then the default bench line; logs under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.txt
if [ -z "$SKIP_TESTS" ]; then
timeout ${PYTEST_TIMEOUT:-1800} python -m pytest tests -m gpu -q --maxfail=${MAXFAIL:-40} -p no:cacheprovider --durations=15 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?"
tail -45 gpurun_out/pytest_gpu.txt
fi
if [ -z "$SKIP_BENCH" ]; then
timeout 900 python bench.py ${BENCH_ARGS} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
fi
