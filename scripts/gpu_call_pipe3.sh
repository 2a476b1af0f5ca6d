#!/bin/bash
mkdir -p gpurun_out
SGLM_CDC_VARIANT=9 timeout 600 python -m pytest tests/test_gpu_cd_cluster.py -x -q -k "every_shape or multi_part" 2>&1 | tail -2
DIAG_EVERY=1 DIAG_PLANS="4x2@0.3,0x0#0;4x2@0.3,0x0#8;4x2@0.3,0x0#9;4x2@0.3,0x0#10;4x2@0.3,0x0#11;4x4@0.3,0x0#0;4x4@0.3,0x0#10;4x4@0.3,0x0#11" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r2_cd_pipe3.log | grep variant
