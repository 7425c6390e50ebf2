#!/bin/bash
# the default bench invocation (no flags) on the last tree of round 2
mkdir -p gpurun_out
timeout 280 python bench.py > gpurun_out/r3h_bench_default.json 2> gpurun_out/r3h_bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3h_bench_default.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','steps','warmup')}, 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
print(json.dumps(d['secondary']['neural_sde'])[:600]); print(json.dumps(d['secondary']['ffjord'])[:700])
PY
tail -2 gpurun_out/r3h_bench_default.err
