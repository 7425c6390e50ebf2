#!/bin/bash
# GPU call N: row A6 after moving the heuristic's adjoint into the cluster-4 sweeps
mkdir -p gpurun_out
python tools/a6_check.py 2>&1 | grep -v "blocks\|Warning\|per = " | tail -12 > gpurun_out/r2n_a6_check.txt
python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/r2n_suite.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2n_ncu_bench.log 2>&1
cat gpurun_out/r2n_a6_check.txt; tail -4 gpurun_out/r2n_suite.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2n_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['run']['nfe_mean'], d['run']['us_per_nfe'], d['fixed_work']['value'], d['fixed_work']['ms_per_step'], d['e2e']['value'], d.get('grad_check'))
PY
