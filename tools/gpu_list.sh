#!/bin/bash
# ncu launch list only (device time of every engine kernel) for workload $WORKLOAD; output in gpurun_out/
mkdir -p gpurun_out
W=${WORKLOAD:-c2}
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gb:: -c 2400 --csv \
    --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps ${LIST_STEPS:-2} --warmup 1 --no-e2e --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; tail -1 gpurun_out/ncu_list.log | cut -c1-1500
