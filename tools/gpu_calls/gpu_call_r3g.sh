#!/bin/bash
# warp-stall samples of the two persistent kernels of the main path (is the instruction cache a limit? kernels are 170-220 KB of SASS)
mkdir -p gpurun_out; : > gpurun_out/r3g_stalls.txt
for k in fwd4s bwd4tc; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -f -o /tmp/p_$k python tools/bwd_ab.py > /dev/null 2>&1
  echo "== $k" >> gpurun_out/r3g_stalls.txt
  ncu -i /tmp/p_$k.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
d={k:x for k,x in zip(h,v)}
tot=float(d.get('smsp__pcsamp_sample_count','0').replace(',','') or 0)
print('samples',tot,'duration_ns',d.get('gpu__time_duration.sum'))
rows=[(float(x.replace(',','')),k) for k,x in d.items() if k.startswith('smsp__pcsamp_warps_issue_stalled_') and not k.endswith('_not_issued')]
for x,k in sorted(rows,reverse=True)[:12]: print('  %-60s %8.0f %.3f'%(k[len('smsp__pcsamp_warps_issue_stalled_'):],x,x/max(tot,1)))
" >> gpurun_out/r3g_stalls.txt
done
cat gpurun_out/r3g_stalls.txt
