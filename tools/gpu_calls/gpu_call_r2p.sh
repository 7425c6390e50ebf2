#!/bin/bash
# GPU call P: row A6 with the lean tensor-core sweep
mkdir -p gpurun_out
python tools/sweep_variants.py 2>&1 | tail -3 > gpurun_out/r2p_sweep.txt
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2p_suite.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2p_ncu_bench.log 2>&1
cat gpurun_out/r2p_sweep.txt; tail -6 gpurun_out/r2p_suite.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2p_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['run']['nfe_mean'], d['run']['us_per_nfe'], d['fixed_work']['value'], d['fixed_work']['ms_per_step'], d['e2e']['value'])
PY
