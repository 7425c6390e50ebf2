#!/bin/bash
# compute-sanitizer initcheck + synccheck over the end-to-end cases (uninitialised global-memory reads, barrier misuse)
mkdir -p gpurun_out; : > gpurun_out/r3b_initcheck.txt
for c in toy mnist ffma4 chain gru sde ffjord; do
  echo "== initcheck $c" >> gpurun_out/r3b_initcheck.txt
  timeout 500 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | grep -v "^=========     at\|^=========     by\|Host Frame\|^=========         in " | tail -25 >> gpurun_out/r3b_initcheck.txt
done
for c in mnist sde ffjord; do
  echo "== synccheck $c" >> gpurun_out/r3b_initcheck.txt
  timeout 500 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | tail -4 >> gpurun_out/r3b_initcheck.txt
done
cat gpurun_out/r3b_initcheck.txt | cut -c1-220
