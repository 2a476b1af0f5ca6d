#!/bin/bash
mkdir -p gpurun_out
DIAG_EVERY=2 DIAG_PLANS="4x4@0.3,0x0;2x4#12,4x4#48,4x2#165,0x0;2x4#12,4x4#96,4x2#117,0x0;4x4#24,4x2#201,0x0;2x4#24,4x4#48,4x2#153,0x0;2x4#12,4x4#213,0x0;2x4#12,4x4#24,4x2#189,0x0" timeout 600 python scripts/cd_timers.py 2>&1 | grep variant | cut -c1-130
DIAG_EVERY=4 DIAG_PLANS="4x4@0.3,0x0;2x4#12,4x4#100,0x0;2x4#24,4x4#88,0x0;2x4#12,4x4#48,4x2#52,0x0;2x4#12,4x4#24,4x2#76,0x0" timeout 600 python scripts/cd_timers.py 2>&1 | grep variant | cut -c1-130
DIAG_EVERY=8 DIAG_PLANS="2x4@0.3,0x0;1x8#6,2x4#50,0x0;2x4#12,4x4#44,0x0" timeout 600 python scripts/cd_timers.py 2>&1 | grep variant | cut -c1-130
