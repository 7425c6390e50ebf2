#!/bin/bash
# validation after the dense_wgrad_kernel rewrite: whole GPU suite, smoke, sanitizer of the chain / GRU / FFJORD contractions, bench
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r3a_suite.txt
python __graft_entry__.py smoke > gpurun_out/r3a_smoke.txt 2>&1
for c in ffjord chain gru; do
  echo "== memcheck $c" >> gpurun_out/r3a_sanitizer.txt
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | tail -3 >> gpurun_out/r3a_sanitizer.txt
  echo "== racecheck $c" >> gpurun_out/r3a_sanitizer.txt
  timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py $c 2>&1 | tail -2 >> gpurun_out/r3a_sanitizer.txt
done
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
tail -4 gpurun_out/r3a_suite.txt; tail -3 gpurun_out/r3a_smoke.txt; cat gpurun_out/r3a_sanitizer.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3a_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['value'], 'fixed', d['fixed_work']['value'], 'frac', d['roofline']['frac'], d.get('grad_check'))
print(d['secondary'])
PY
tail -2 gpurun_out/r3a_bench.err
