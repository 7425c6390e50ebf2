#!/bin/bash
# GPU call M: row A6 tests, full suite, smoke, bench line + launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -q -m gpu -k "first_dt" 2>&1 | tail -30 > gpurun_out/r2m_a6.txt
python -m pytest tests -q -m gpu 2>&1 | tail -40 > gpurun_out/r2m_suite.txt
python __graft_entry__.py smoke > gpurun_out/r2m_smoke.txt 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-secondary > gpurun_out/r2m_ncu_bench.log 2>&1
tail -12 gpurun_out/r2m_a6.txt; tail -8 gpurun_out/r2m_suite.txt; tail -3 gpurun_out/r2m_smoke.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2m_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['run']['nfe_mean'], d['run']['us_per_nfe'], d['fixed_work']['value'], d['fixed_work']['ms_per_step'], d['e2e']['value'], d.get('grad_check'))
PY
