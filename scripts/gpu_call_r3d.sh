#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3d_pytest.log
BENCH_GAPS=1 timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; grep -E "gaps|gc\]" gpurun_out/r3d_bench.err | cut -c1-900
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3d_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['best_params'], d['best_score'])
print(r['per_entry_ms_per_step'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_|coldot|pb_gemm|quadform" -c 60 --csv --log-file gpurun_out/r3d_ncu_tc.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r3d_ncu_tc.csv')))
h=next(i for i,r in enumerate(rows) if 'Kernel Name' in r)
kn=rows[h].index('Kernel Name'); mv=rows[h].index('Metric Value'); mu=rows[h].index('Metric Unit')
for r in rows[h+1:][-26:]:
    if len(r)>mv: print(r[kn][:40], r[mv], r[mu])
PY
