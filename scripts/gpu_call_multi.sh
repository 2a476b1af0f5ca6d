#!/bin/bash
# usage: gpu_call_multi.sh N  — strong-scaling bench (one grid over N GPUs); at N = 2 also the bit-identity check
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 900 $TR --master-port 29511 scripts/strong_check.py > gpurun_out/r2g_strong_check_2gpu.txt 2>&1; echo "strong_check rc=$?"; grep -E "STRONG CHECK|one GPU|Error|assert" gpurun_out/r2g_strong_check_2gpu.txt | tail -5
fi
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2g_bench$N.json 2> gpurun_out/r2g_bench$N.err; echo "bench rc=$?"; tail -2 gpurun_out/r2g_bench$N.err
python - $N <<'PY'
import json, sys
d=json.load(open(f'gpurun_out/r2g_bench{sys.argv[1]}.json'))
print('N', d['n_gpus'], 'value', round(d['value'],1), 'ms', round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1) if d.get('e2e') else None, d.get('strong_scaling'))
PY
