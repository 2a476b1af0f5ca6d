#!/bin/bash
# full GPU test suite + full-size bench with the three-part default CD plan
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['parts_ms_per_step'], d['roofline']['plan'], d['roofline']['chain'])
PY
