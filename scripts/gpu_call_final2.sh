#!/bin/bash
# round 2 final measurements (second pass, final code): smoke, bench lines, ncu launch list + full capture at full size
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench.err
timeout 900 python scripts/config_bench.py c1 c2 c4 > gpurun_out/r2g_configs.jsonl 2> gpurun_out/r2g_configs.err; echo "configs rc=$?"
timeout 900 python bench.py --mode sessions --sessions-per-gpu 8 --T 250000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2g_bench_sessions8.json 2> gpurun_out/r2g_bench_s8.err; echo "sessions rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_ncu_launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2g_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"enet_cd|tc_gram_i8|tc_slice|tc_expand|tc_cell_sum|tc_combine|tc_zero|pb_gemm_nn|coldot_kernel" -s 16 -c 16 -o gpurun_out/r2g_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2g_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2g_full.ncu-rep --page raw --csv > gpurun_out/r2g_ncu_full_raw.csv 2>/dev/null; echo "export rc=$?"
rm -f gpurun_out/r2g_full.ncu-rep
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['gpu_launches'])
PY
