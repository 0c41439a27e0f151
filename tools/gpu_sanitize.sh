#!/bin/bash
# compute-sanitizer over the smoke invocation (2 boosting iterations x {oblivious/cosine, greedy/L2} against the oracle)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_$tool.txt 2>&1
  echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok" gpurun_out/sanitizer_$tool.txt | tail -3
done
