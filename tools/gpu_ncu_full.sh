#!/bin/bash
# `ncu --set full` captures of the kernels DESIGN.md budgets (one short bench run per capture); reports in gpurun_out/
mkdir -p gpurun_out
cap() {   # name, workload, kernel regex, skip, count
  timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$3 -s $4 -c $5 -f \
      -o gpurun_out/full_$1 python bench.py --workload $2 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/full_$1.log 2>&1
  echo "capture $1 rc=$?"; ls -la gpurun_out/full_$1.ncu-rep 2>/dev/null | awk '{print $5}'
}
cap hist_c2 c2 hist_stream 6 3
cap walk_c2 c2 "wide_walk_kernel" 2 3
cap dwalk_c2 c2 wide_dense_walk 1 2
cap part_c2 c2 "part_" 12 2
cap scan_c2 c2 "scan_kernel" 8 2
cap tabs_c3 c3 "wide_tabs_kernel" 4 2
cap hist_c3 c3 hist_stream 8 2
cap bits_c3 c3 wide_bits 3 2
cap hist_c5 c5 hist_stream 6 2
cap predict_c4 c4 predict_tiles 1 1
cap sel_c2 c2 "sel_pass_kernel" 0 3
