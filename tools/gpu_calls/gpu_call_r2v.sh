#!/bin/bash
# compute-sanitizer over the small end-to-end cases (first-dt term included), then fresh ncu captures of the two big kernels
mkdir -p gpurun_out
for tool in memcheck; do
  for c in toy stream mnist ffma4 cluster8 chain gru sde; do
    echo "== $tool $c" >> gpurun_out/r2v_sanitizer.txt
    timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | grep -v "^$" | tail -4 >> gpurun_out/r2v_sanitizer.txt
  done
done
for c in toy mnist chain; do
  echo "== racecheck $c" >> gpurun_out/r2v_sanitizer.txt
  timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | grep -v "^$" | tail -4 >> gpurun_out/r2v_sanitizer.txt
done
cat gpurun_out/r2v_sanitizer.txt | tail -60
