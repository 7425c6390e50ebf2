#!/bin/bash
# final validation of round 2: whole GPU suite, smoke, sanitizer of the new reverse sweeps, both bench arms
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2y_suite.txt
python __graft_entry__.py smoke > gpurun_out/r2y_smoke.txt 2>&1
for c in sde ffjord; do
  echo "== memcheck $c" >> gpurun_out/r2y_sanitizer.txt
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | tail -3 >> gpurun_out/r2y_sanitizer.txt
  echo "== racecheck $c" >> gpurun_out/r2y_sanitizer.txt
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | tail -2 >> gpurun_out/r2y_sanitizer.txt
done
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_ref.json 2> gpurun_out/r2y_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
tail -4 gpurun_out/r2y_suite.txt; tail -3 gpurun_out/r2y_smoke.txt; cat gpurun_out/r2y_sanitizer.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2y_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'fixed', d['fixed_work']['value'], 'frac', d['roofline']['frac'], d.get('grad_check'))
print(d['secondary'])
r=json.loads(open('gpurun_out/r2y_ref.json').read().strip().splitlines()[-1]); print('reference arm', r['value'])
PY
tail -2 gpurun_out/r2y_bench.err
