#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cd_cluster.py tests/test_gpu_parity.py -x -q 2>&1 | tail -2
DIAG_EVERY=1,8 DIAG_PLANS="4x2@0.3,0x0;4x4@0.3,0x0;2x4@0.3,0x0" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r3b_cd_timers.log | grep variant
