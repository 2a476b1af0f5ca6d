#!/bin/bash
# full-size CD plan sweep through bench.py (SGLM_CD_PLAN), device-resident timing only
mkdir -p gpurun_out
: > gpurun_out/r2t_plans.txt
for plan in "4x4#24,4x2#426,0x0" "4x4#48,4x2#402,0x0" "4x4#72,4x2#378,0x0" "4x4#48,4x2#354,0x0" "4x4#48,4x2#450,0x0" "4x4#96,4x2#354,0x0" "2x4#12,4x4#36,4x2#402,0x0" "4x4#48,4x2#300,0x0"; do
  SGLM_CD_PLAN="$plan" timeout 600 python bench.py --no-cpu-baseline --no-e2e --no-probes --steps 2 --warmup 1 > gpurun_out/r2t_tmp.json 2> gpurun_out/r2t_tmp.err
  python - "$plan" <<'PY' | tee -a gpurun_out/r2t_plans.txt
import json, sys
d=json.load(open('gpurun_out/r2t_tmp.json'))
r=d['roofline']
print(sys.argv[1], 'step', round(d['ms_per_step'],1), 'cd', round(r['per_entry_ms_per_step']['sglm_enet_cd (cluster + per-model parts, concurrent)'],1), [(p['shape'],p['models'],p['max_blocks_one_model']) for p in r['parts'][:4]])
PY
done
