#!/bin/bash
# initcheck of the cluster-4 case again, this time keeping the FIRST reports and their count
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool initcheck python tools/sanitize_cases.py ffma4 > /tmp/ic.txt 2>&1
grep -c "Uninitialized access" /tmp/ic.txt > gpurun_out/r3d_initcheck_ffma4.txt
grep -B1 -A1 "Uninitialized access\|Host API memory access" /tmp/ic.txt | grep -v "Saved host\|^--" | awk '!seen[$0]++' | head -150 >> gpurun_out/r3d_initcheck_ffma4.txt
grep -E "cudaMemcpy|cudaMemset|rnde_|in .*libregnde" /tmp/ic.txt | awk '!seen[$0]++' | head -20 >> gpurun_out/r3d_initcheck_ffma4.txt
tail -3 /tmp/ic.txt >> gpurun_out/r3d_initcheck_ffma4.txt
head -120 gpurun_out/r3d_initcheck_ffma4.txt | cut -c1-200
