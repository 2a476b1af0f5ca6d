#!/bin/bash
# round 2 final measurements, 1 GPU: tests, bench lines, ncu launch list + full capture at full size
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r2z_bench.err
timeout 900 python scripts/config_bench.py c1 c2 c4 > gpurun_out/r2z_configs.jsonl 2> gpurun_out/r2z_configs.err; echo "configs rc=$?"; tail -3 gpurun_out/r2z_configs.err
timeout 900 python bench.py --mode sessions --sessions-per-gpu 8 --T 250000 --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/r2z_bench_sessions8.json 2> gpurun_out/r2z_bench_s8.err; echo "sessions rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_ncu_launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 2400 ncu --set full --clock-control none --import-source on -k regex:"enet_cd|tc_gram_i8|tc_slice|tc_expand|tc_cell_sum|tc_combine" -s 9 -c 9 -o gpurun_out/r2z_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2z_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2z_full.ncu-rep --page raw --csv > gpurun_out/r2z_ncu_full_raw.csv 2>/dev/null; echo "export rc=$?"
SZ=$(stat -c %s gpurun_out/r2z_full.ncu-rep 2>/dev/null || echo 0); if [ "$SZ" -gt 30000000 ]; then rm -f gpurun_out/r2z_full.ncu-rep; echo "removed large rep ($SZ)"; fi
ls -la gpurun_out | tail -12
