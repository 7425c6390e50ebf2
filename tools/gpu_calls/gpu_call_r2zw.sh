#!/bin/bash
# ncu --set full of one dense_wgrad_kernel variant: where the time goes (stall reasons, hottest SASS lines)
# (record of what ran: build_variants/lib_*.so were whole-library builds of intermediate revisions of chain.cuh with -DRNDE_CW_BATCH / _MINB / _NT / _TPT;
#  the directory is not kept -- results in profiles/r2zz_next_rows_ncu.txt)
mkdir -p gpurun_out
V=${1:-b4t512}
cp build_variants/lib_$V.so regneuralde/jl_b200/libregnde.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_wgrad -s 1 -c 1 -f -o /tmp/dw python tools/ffjord_step.py 2>&1 | tail -2
ncu -i /tmp/dw.ncu-rep --page details > gpurun_out/r2zw_${V}_details.txt 2>&1
ncu -i /tmp/dw.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
r=list(csv.reader(sys.stdin)); h=r[0]; v=r[-1]
for k,x in zip(h,v):
    if any(t in k for t in ('pcsamp','pipe_fp64','pipe_lsu','pipe_xu','gpu__time_duration.sum','bank_conflicts','wavefronts_mem_shared','achieved_occupancy','issue_active')): print(k,x)
" > gpurun_out/r2zw_${V}_raw.txt
ncu -i /tmp/dw.ncu-rep --page source --csv --print-source sass 2>/dev/null > /tmp/src.csv
python - <<'PY' > gpurun_out/r2zw_${V}_hot.txt
import csv
rows=list(csv.reader(open('/tmp/src.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0] in ('#','Address') or (r and 'Source' in r)]
h=rows[hi[0]] if hi else rows[0]
print(h[:12])
si=None
for j,c in enumerate(h):
    if c.strip().startswith('# Samples') or c.strip()=='Sampling Data (All)' or 'Samples' in c: si=j; break
print('sample col',si)
body=[r for r in rows[(hi[0] if hi else 0)+1:] if len(r)>si]
def f(x):
    try: return float(x.replace(',',''))
    except: return 0.0
tot=sum(f(r[si]) for r in body)
body.sort(key=lambda r:-f(r[si]))
for r in body[:40]: print(f(r[si])/max(tot,1), r[:3], r[si])
PY
cat gpurun_out/r2zw_${V}_hot.txt | head -60
grep -E "Duration|Registers Per|Achieved Occ|Theoretical Occ|Block Limit" gpurun_out/r2zw_${V}_details.txt
grep pcsamp gpurun_out/r2zw_${V}_raw.txt | sort -k2 -n -r | head -12
