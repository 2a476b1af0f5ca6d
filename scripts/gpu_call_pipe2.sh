#!/bin/bash
mkdir -p gpurun_out
DIAG_EVERY=1 DIAG_PLANS="4x2@0.3,0x0#0;4x2@0.3,0x0#12;4x2@0.3,0x0#13;4x2@0.3,0x0#14;4x2@0.3,0x0#15;4x4@0.032,4x2@0.268,0x0#12;4x4@0.032,4x2@0.268,0x0#14" timeout 900 python scripts/cd_timers.py 2>&1 | tee gpurun_out/r2p_cd_timers.log
