cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 900 python scripts/config_bench.py c1 c2 c4 > gpurun_out/r2e_configs.jsonl 2> gpurun_out/r2e_configs.err; echo "configs rc=$?"; tail -8 gpurun_out/r2e_configs.err; cut -c1-900 gpurun_out/r2e_configs.jsonl
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2e_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_ncu_launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2e_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 1500 ncu --set full --clock-control none -k regex:"enet_cd|tc_gram_i8|tc_slice|tc_colmax|timeshift_staged" -s 6 -c 6 -o gpurun_out/r2e_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2e_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/r2e_full.ncu-rep --page raw --csv > gpurun_out/r2e_ncu_full_raw.csv 2>/dev/null; echo "export rc=$?"
ls -la gpurun_out/
SZ=$(stat -c %s gpurun_out/r2e_full.ncu-rep 2>/dev/null || echo 0); if [ "$SZ" -gt 40000000 ]; then rm -f gpurun_out/r2e_full.ncu-rep; echo "removed large rep"; fi
du -sh gpurun_out
