#!/bin/bash
# row N4 complete: FFJORD tests (gradient, sample round trip, ADAM), the whole GPU suite, the bench line with its FFJORD row
mkdir -p gpurun_out
python -m pytest tests/test_gpu_ffjord.py -q -m gpu 2>&1 | tail -12 > gpurun_out/r2x_ffjord.txt
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2x_suite.txt
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
cat gpurun_out/r2x_ffjord.txt; tail -4 gpurun_out/r2x_suite.txt; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2x_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['secondary'])
PY
tail -3 gpurun_out/r2x_bench.err
