#!/bin/bash
# ncu launch list (device time of every engine kernel) for one workload; outputs in gpurun_out/
W=${WORKLOAD:-c2}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gb:: -c ${NCU_COUNT:-3000} --csv \
    --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps ${STEPS:-2} --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ncu_list_$W.log 2>&1
echo "ncu list rc=$?"; tail -2 gpurun_out/ncu_list_$W.log | cut -c1-300
