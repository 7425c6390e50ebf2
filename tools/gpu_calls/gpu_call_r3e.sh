#!/bin/bash
# launch lists (ncu gpu__time_duration) of one FFJORD and one Neural-SDE training step on the final library
mkdir -p gpurun_out
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_ffjord_launches.csv python tools/ffjord_step.py > /dev/null 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3e_nsde_launches.csv python tools/nsde_step.py > /dev/null 2>&1
python - <<'PY'
import csv, collections
for name in ("ffjord", "nsde"):
    rows = list(csv.reader(open(f"gpurun_out/r3e_{name}_launches.csv")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
    body = [r for r in rows[hdr + 1:] if len(r) > vi and r[0].isdigit()]
    half = body[len(body) // 2:]          # the second of the two identical steps
    agg = collections.OrderedDict()
    for r in half:
        a = agg.setdefault(r[ki][:80], [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"== {name}: second step, {len(half)} launches, {tot/1e6:.3f} ms of kernel time")
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1])[:10]:
        print(f"  {n:80s} x{a[0]:3d} {a[1]/1e6:8.3f} ms  {a[1]/tot:.3f}")
PY
