#!/bin/bash
for parts in 2 3 4 5 6 8 11; do
  SGLM_TUNING=1 SGLM_TC_PARTS=$parts timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-probes --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('parts', $parts, 'step', round(d['ms_per_step'],1), 'gram', round(r['per_entry_ms_per_step']['sglm_gram_tc_f64'],2))"
done
