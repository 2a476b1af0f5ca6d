set -x
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -20 gpurun_out/r2c_pytest.log
timeout 900 python scripts/config_bench.py c1 c2 c4 > gpurun_out/r2c_configs.jsonl 2> gpurun_out/r2c_configs.err; echo "configs rc=$?"; tail -5 gpurun_out/r2c_configs.err; cut -c1-1500 gpurun_out/r2c_configs.jsonl
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2c_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_ncu_launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2c_ncu_launches.log 2>&1; echo "ncu list rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"enet_cd|tc_gram_i8|tc_slice|tc_colmax|timeshift_staged" -c 14 -o gpurun_out/r2c_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-probes > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -5 gpurun_out/r2c_ncu_full.log
ls -la gpurun_out/ | tail -12
