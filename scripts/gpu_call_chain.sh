#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3a_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3a_pytest.log
DIAG_EVERY=1,8 DIAG_PLANS="4x2@0.3,0x0;4x4@0.3,0x0;2x4@0.3,0x0" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r3a_cd_timers.log | grep variant
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3a_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3a_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['best_params'], d['best_score'])
print(r['per_entry_ms_per_step'])
print(r['launch_ms'])
PY
