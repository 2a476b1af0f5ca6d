#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lag_gram.py -x -q > gpurun_out/r2u_pytest_lag.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2u_pytest_lag.log
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
print(r['per_entry_ms_per_step'])
print({k:(v.get('frac'), v.get('ms')) for k,v in r['other_kernels'].items()})
print(d['best_params'], d['best_score'])
PY
