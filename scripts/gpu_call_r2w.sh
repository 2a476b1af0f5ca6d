#!/bin/bash
mkdir -p gpurun_out
for ms in 200 1000 200 1000; do
  BENCH_SAMPLER_MS=$ms timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-probes --steps 3 --warmup 3 > gpurun_out/r2w_tmp.json 2> gpurun_out/r2w_tmp.err
  python - $ms <<'PY'
import json, sys
d=json.load(open('gpurun_out/r2w_tmp.json'))
r=d['roofline']
print('sampler ms', sys.argv[1], 'step', round(d['ms_per_step'],1), 'sum of entries', round(sum(r['per_entry_ms_per_step'].values()),1), d['clocks'])
PY
done
