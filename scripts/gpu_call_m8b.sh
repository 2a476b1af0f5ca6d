#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cd_cluster.py -x -q 2>&1 | tail -2
DIAG_EVERY=1 DIAG_PLANS="8x4@0.3,0x0;8x2@0.3,0x0;4x2@0.3,0x0" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r3c_cd_timers.log | grep variant
DIAG_TAG=m8b DIAG_PLANS="4x4#24,4x2#426,0x0;4x4#24,8x2#426,0x0;4x4#24,8x4#426,0x0;4x4#24,8x4#96,8x2#330,0x0;4x4#24,8x4#192,8x2#384,0x0;8x4#48,8x2#402,0x0;4x4#24,8x2#576,0x0" timeout 600 python scripts/cd_dump.py 2>&1 | tee gpurun_out/r3c_cd_dump.log
