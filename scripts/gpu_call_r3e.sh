#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r3e_pytest.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 900 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/r3e_bench.json 2> gpurun_out/r3e_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r3e_bench.json'))
r=d['roofline']
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['best_score'])
print({k: round(v,2) for k,v in r['per_entry_ms_per_step'].items()})
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active,lts__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"tc_gram_i8" -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes 2>&1 | grep -E "duration|bytes_read|hit_rate|imma|lts__throughput"
