#!/bin/bash
# bench + ncu evidence on the GPU box; outputs in gpurun_out/
mkdir -p gpurun_out
W=${WORKLOAD:-c2}
python bench.py --workload $W --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_$W.json 2> gpurun_out/bench_$W.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/bench_$W.json; tail -3 gpurun_out/bench_$W.err
if [ -n "$REFARM" ]; then
python bench.py --impl reference --workload $W --steps 2 --warmup 1 --ref-budget ${REF_BUDGET:-60} > gpurun_out/bench_${W}_reference.json 2>> gpurun_out/bench_$W.err; cut -c1-600 gpurun_out/bench_${W}_reference.json
fi
if [ -n "$NCU" ]; then
# launch list (device time of every engine kernel; compare SHARES, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:gb:: -c 2400 --csv \
    --log-file gpurun_out/launches_$W.csv python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
# full capture of the dominant kernel (3 launches: root level + two deeper levels)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:hist_stream -s ${NCU_SKIP:-6} -c 3 \
    -o gpurun_out/prof_hist_$W python bench.py --workload $W --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
fi
