#!/bin/bash
# compute-sanitizer over one call of every kernel family: memcheck on everything, racecheck and
# synccheck on the fit / forward kernels that use the TMA rings, mbarriers and tcgen05.  Summaries -> gpurun_out/sanitize_*.txt
mkdir -p gpurun_out
S=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck ${SANITIZE_TOOLS:-racecheck synccheck}; do
  what=all; [ "$tool" != memcheck ] && what=${SANITIZE_WHAT:-fit_ragged}
  timeout ${SANITIZE_LIMIT:-900} $S --tool $tool --print-limit 20 python scripts/sanitize_driver.py $what > gpurun_out/sanitize_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/sanitize_$tool.log
  { echo "== compute-sanitizer --tool $tool: python scripts/sanitize_driver.py $what"; grep -E "^ok |driver done|ERROR SUMMARY|RACECHECK SUMMARY|rc=|Error:|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -30; } > gpurun_out/sanitize_$tool.txt
  cat gpurun_out/sanitize_$tool.txt
done
