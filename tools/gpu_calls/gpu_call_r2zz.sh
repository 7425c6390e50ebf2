#!/bin/bash
# ncu --set full of the kernels of rows N2 (Neural SDE) and N4 (FFJORD): one launch each out of a training step
mkdir -p gpurun_out
cap() {  # name regex skip script
  (timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_r2zz_$1 python $4 2>&1 | tail -2) >> gpurun_out/r2zz_ncu.txt
  ncu -i gpurun_out/prof_r2zz_$1.ncu-rep --page details > gpurun_out/r2zz_$1_details.txt 2>&1
  ncu -i gpurun_out/prof_r2zz_$1.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
for k,x in zip(h,v):
    if any(t in k for t in ('pipe_','stall','dram__bytes','gpu__time_duration','smsp__inst_executed.sum','l1tex__data_bank','lsu_mem_shared')) and 'pct' not in k or 'inst_executed_pipe' in k: print(k,x)
" > gpurun_out/r2zz_$1_raw.txt
  rm -f gpurun_out/prof_r2zz_$1.ncu-rep     # 28 MB each: more than gpurun copies back
}
: > gpurun_out/r2zz_ncu.txt
cap sde_fwd 'sde_kernel' 1 tools/nsde_step.py
cap sde_bwd 'sde_bwd_kernel' 1 tools/nsde_step.py
cap ffjord_fwd 'fwd_kernel' 1 tools/ffjord_step.py
cap ffjord_bwd 'bwd_kernel' 1 tools/ffjord_step.py
cap ffjord_wgrad 'dense_wgrad_kernel' 1 tools/ffjord_step.py
rm -f gpurun_out/prof_r2zz_*.ncu-rep.tmp
cat gpurun_out/r2zz_ncu.txt
for f in sde_fwd sde_bwd ffjord_fwd ffjord_bwd ffjord_wgrad; do echo == $f; grep -E "Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Dynamic Shared|Achieved Occupancy|Grid Size|Block Size|DRAM Throughput|Compute \(SM\)" gpurun_out/r2zz_${f}_details.txt | head -12; done
