#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_at_scale.py tests/test_gpu_lag_gram.py -x -q 2>&1 | tail -2
for env in "SGLM_TUNING=0" "SGLM_TUNING=1 SGLM_TC_LOCKSTEP=1"; do
  env $env timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-probes --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$env', 'step', round(d['ms_per_step'],1), 'gram', round(r['per_entry_ms_per_step']['sglm_gram_tc_f64'],2))"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"tc_gram_i8" -s 1 -c 1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes 2>&1 | grep -E "duration|bytes_read|hit_rate|imma"
